#!/bin/bash
set +e
mkdir -p gpurun_out
for m in 2 1 0; do echo "== pytest search parity IDB_VIS_MODE=$m"; IDB_VIS_MODE=$m timeout 900 python -m pytest tests/test_gpu_search_parity.py -q -m gpu -x > gpurun_out/pytest_vm$m.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_vm$m.log; done
echo "== tune 1M"; timeout 1200 python scripts/tune_search.py --n 1000000 > gpurun_out/tune8_1M.jsonl 2> gpurun_out/tune8_1M.log; echo "rc=$?"; cat gpurun_out/tune8_1M.jsonl; tail -3 gpurun_out/tune8_1M.log
