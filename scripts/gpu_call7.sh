#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu search parity (smem visited default)"; timeout 900 python -m pytest tests/test_gpu_search_parity.py -q -m gpu -x > gpurun_out/pytest_vm1.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_vm1.log
echo "== pytest gpu search parity (IDB_VIS_SMEM=0)"; IDB_VIS_SMEM=0 timeout 900 python -m pytest tests/test_gpu_search_parity.py -q -m gpu -x > gpurun_out/pytest_vm0.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_vm0.log
echo "== tune 1M"; timeout 1200 python scripts/tune_search.py --n 1000000 > gpurun_out/tune7_1M.jsonl 2> gpurun_out/tune7_1M.log; echo "rc=$?"; cat gpurun_out/tune7_1M.jsonl; tail -3 gpurun_out/tune7_1M.log
