#!/bin/bash
# First GPU validation: smoke, parity tests, sanitizer, reduced-size bench, ncu launch list + one full capture.
set +e
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/gpu_sanitize_case.py > gpurun_out/sanitize.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/sanitize.log
echo "== bench 200k"; timeout 900 python bench.py --n 200000 --graph oracle --steps 10 --warmup 3 > gpurun_out/bench_200k.json 2> gpurun_out/bench_200k.log; echo "rc=$?"; tail -4 gpurun_out/bench_200k.log; cat gpurun_out/bench_200k.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --n 200000 --graph oracle --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -o gpurun_out/prof_search python bench.py --n 200000 --graph oracle --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out
