#!/bin/bash
# Call 24: ncu --set full (with source) of the current K1 (bitmap visited tier) and of the PIPE variant, headline config.
set +e
mkdir -p gpurun_out
echo "== ncu full, default K1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 8 -c 2 -f -o gpurun_out/k1_bitmap python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_k1_bitmap.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_k1_bitmap.log
echo "== ncu full, PIPE variant"
IDB_VARIANT=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 8 -c 2 -f -o gpurun_out/k1_pipe python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_k1_pipe.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_k1_pipe.log
ls -la gpurun_out/*.ncu-rep
