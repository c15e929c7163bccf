#!/bin/bash
# Call 31: bucket set, slot choice = function of the id (repeated ids in a row claim the same slot): parity + A/B.
set +e
mkdir -p gpurun_out
echo "== parity (search + bf16)"
timeout 400 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_bf16.py -q -m gpu > gpurun_out/pytest_call31.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_call31.log
echo "== sweep 1M x 128"
timeout 200 python scripts/tune_search.py --n 1000000 --steps 6 --configs "IDB_OPT=0;IDB_OPT=16" > gpurun_out/tune_call31.jsonl 2> gpurun_out/tune_call31.log; echo "rc=$?"; cat gpurun_out/tune_call31.jsonl
