#!/bin/bash
set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== C: 2.5M points, 100k queries"; timeout 900 $TR --master-port 29511 scripts/sharded_check.py --points 2500000 --queries 100000 --no-oracle --bench 5 > gpurun_out/shC.json 2> gpurun_out/shC.log; echo "rc=$?"; cat gpurun_out/shC.json; grep DIAG gpurun_out/shC.log | cut -c1-1200
