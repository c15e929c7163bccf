#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python scripts/gather_ceiling.py > gpurun_out/gather_ceiling.jsonl 2> gpurun_out/gather_ceiling.log; echo "rc=$?"; cat gpurun_out/gather_ceiling.jsonl; tail -3 gpurun_out/gather_ceiling.log
