#!/bin/bash
# Call 26: gather ceiling with K1's visited-style atomic stream riding along (what is the ceiling for K1's traffic MIX?)
set +e
mkdir -p gpurun_out
timeout 600 python scripts/gather_ceiling.py --mix --sizes > gpurun_out/gather_mix.jsonl 2> gpurun_out/gather_mix.log; echo "rc=$?"; cat gpurun_out/gather_mix.jsonl; tail -3 gpurun_out/gather_mix.log
