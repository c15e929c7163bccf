#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python scripts/sharded_1gpu_layout.py 10000000 > gpurun_out/sharded_1gpu_layout.json 2> gpurun_out/sharded_1gpu_layout.log; echo "rc=$?"; cat gpurun_out/sharded_1gpu_layout.json; tail -3 gpurun_out/sharded_1gpu_layout.log
