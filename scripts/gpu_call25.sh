#!/bin/bash
# Call 25: packed-math / FULL-batch K1 + bitmap visited tier: full GPU test-suite, A/B sweep, ncu --set full of K1 (traffic + source).
set +e
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== sweep 1M x 128"
timeout 600 python scripts/tune_search.py --n 1000000 --steps 6 --configs "IDB_OPT=0;IDB_VIS_BITMAP=0;IDB_CTAS_PER_SM=3;IDB_OPT=0" > gpurun_out/tune_call25.jsonl 2> gpurun_out/tune_call25.log; echo "rc=$?"; cat gpurun_out/tune_call25.jsonl; tail -2 gpurun_out/tune_call25.log
echo "== bench (graph cached by the sweep? no: own cache key) + ncu full of K1"
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.log; echo "rc=$?"; cat gpurun_out/bench_quick.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^search_kernel -s 8 -c 2 -f -o gpurun_out/k1_default python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_k1_default.log 2>&1; echo "rc=$?"; grep -c "Profiling" gpurun_out/ncu_k1_default.log
ls -la gpurun_out/*.ncu-rep
