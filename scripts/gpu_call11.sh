#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== tune build 1M"; timeout 1200 python scripts/tune_build.py 1000000 > gpurun_out/tune_build_1M.jsonl 2> gpurun_out/tune_build_1M.log; echo "rc=$?"; cat gpurun_out/tune_build_1M.jsonl; tail -3 gpurun_out/tune_build_1M.log
