#!/bin/bash
# Call 23: K1 working ahead on the predicted next candidate (PIPE, default with the bitmap visited tier).  Parity, then A/B.
set +e
mkdir -p gpurun_out
echo "== parity (defaults: bitmap + PIPE)"
timeout 900 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_bf16.py -x -q -m gpu > gpurun_out/pytest_pipe.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_pipe.log
echo "== sweep 1M x 128"
timeout 900 python scripts/tune_search.py --n 1000000 --steps 6 --configs "IDB_OPT=0;IDB_VARIANT=7;IDB_VIS_BITMAP=0;IDB_OPT=0;IDB_VARIANT=7" > gpurun_out/tune_call23.jsonl 2> gpurun_out/tune_call23.log; echo "rc=$?"; cat gpurun_out/tune_call23.jsonl; tail -3 gpurun_out/tune_call23.log
