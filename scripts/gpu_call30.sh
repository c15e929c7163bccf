#!/bin/bash
# Call 30: bucket set, claims spread over the free slots of a bucket (vs first free slot = IDB_OPT=16): parity + A/B on one box.
set +e
mkdir -p gpurun_out
echo "== parity (search + bf16)"
timeout 600 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_bf16.py -x -q -m gpu > gpurun_out/pytest_call30.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_call30.log
echo "== sweep 1M x 128"
timeout 600 python scripts/tune_search.py --n 1000000 --steps 6 --configs "IDB_OPT=0;IDB_OPT=16;IDB_OPT=0;IDB_OPT=16" > gpurun_out/tune_call30.jsonl 2> gpurun_out/tune_call30.log; echo "rc=$?"; cat gpurun_out/tune_call30.jsonl; tail -2 gpurun_out/tune_call30.log
