#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== bench 1M"; timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1M.json 2> gpurun_out/bench_1M.log; echo "rc=$?"; tail -3 gpurun_out/bench_1M.log; cat gpurun_out/bench_1M.json
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -o gpurun_out/prof_search_1M_v3 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; echo "rc=$?"; cat gpurun_out/bench_ref.json
