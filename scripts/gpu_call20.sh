#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
