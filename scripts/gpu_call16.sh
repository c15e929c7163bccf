#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1500 python scripts/measure_configs.py --which config4s,config4 > gpurun_out/config4.jsonl 2> gpurun_out/config4.log; echo "rc=$?"; cat gpurun_out/config4.jsonl; tail -5 gpurun_out/config4.log
