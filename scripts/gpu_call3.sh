#!/bin/bash
# Call 3: parity after the probe-overlap / load-path trims, occupancy-variant sweep at 1M.
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== tune 1M"; timeout 1200 python scripts/tune_search.py --n 1000000 > gpurun_out/tune3_1M.jsonl 2> gpurun_out/tune3_1M.log; echo "rc=$?"; cat gpurun_out/tune3_1M.jsonl; tail -3 gpurun_out/tune3_1M.log
