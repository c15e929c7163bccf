#!/bin/bash
# Call 27: bucket-set visited tier (32-byte buckets, tables of all resident warps inside the persisting part of L2).
set +e
mkdir -p gpurun_out
echo "== parity (search + bf16, all flavours)"
timeout 900 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_bf16.py -x -q -m gpu > gpurun_out/pytest_buckets.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_buckets.log
echo "== sweep 1M x 128"
timeout 600 python scripts/tune_search.py --n 1000000 --steps 6 --configs "IDB_OPT=0;IDB_VIS_BUCKETS=0;IDB_CTAS_PER_SM=3;IDB_VIS_BUCKETS=0,IDB_VIS_BITMAP=0;IDB_OPT=0" > gpurun_out/tune_call27.jsonl 2> gpurun_out/tune_call27.log; echo "rc=$?"; cat gpurun_out/tune_call27.jsonl; tail -2 gpurun_out/tune_call27.log
