#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1500 python scripts/measure_configs.py --which config0,uniform,config2s,config2 > gpurun_out/configs.jsonl 2> gpurun_out/configs.log; echo "rc=$?"; cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.log
