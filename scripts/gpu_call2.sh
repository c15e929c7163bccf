#!/bin/bash
# Call 2: parity after the K1 restructure + first GPU build tests, knob sweep at 1M, full bench, ncu.
set +e
mkdir -p gpurun_out
echo "== pytest gpu (search parity)"; timeout 900 python -m pytest tests/test_gpu_search_parity.py -q -m gpu -x > gpurun_out/pytest_search.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_search.log
echo "== pytest gpu (build + python module)"; timeout 1200 python -m pytest tests/test_gpu_build.py tests/test_python_module.py -q -m gpu > gpurun_out/pytest_build.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_build.log
GRAPH=gpu
if [ -n "$(ls gpurun_cache/graph_*.npz 2>/dev/null)" ]; then GRAPH=oracle; fi
echo "== tune 1M (graph=$GRAPH)"; timeout 1200 python scripts/tune_search.py --n 1000000 --graph $GRAPH > gpurun_out/tune_1M.jsonl 2> gpurun_out/tune_1M.log; echo "rc=$?"; cat gpurun_out/tune_1M.jsonl; tail -3 gpurun_out/tune_1M.log
echo "== bench 1M"; timeout 1200 python bench.py --graph $GRAPH --steps 20 --warmup 3 > gpurun_out/bench_1M.json 2> gpurun_out/bench_1M.log; echo "rc=$?"; tail -4 gpurun_out/bench_1M.log; cat gpurun_out/bench_1M.json
echo "== ncu launches"; timeout 900 ncu --kernel-name-base mangled -k regex:idb --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --graph $GRAPH --steps 3 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -o gpurun_out/prof_search_1M python bench.py --graph $GRAPH --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out | head -30
