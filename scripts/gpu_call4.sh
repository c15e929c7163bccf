#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (search parity + build)"; timeout 1500 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_build.py -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== tune 1M"; timeout 1200 python scripts/tune_search.py --n 1000000 > gpurun_out/tune4_1M.jsonl 2> gpurun_out/tune4_1M.log; echo "rc=$?"; cat gpurun_out/tune4_1M.jsonl; tail -3 gpurun_out/tune4_1M.log
