#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (bf16 + search parity)"; timeout 1200 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_search_parity.py -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 1200 python scripts/measure_configs.py --which config4s > gpurun_out/config4b.jsonl 2> gpurun_out/config4b.log; echo "rc=$?"; cat gpurun_out/config4b.jsonl; tail -3 gpurun_out/config4b.log
