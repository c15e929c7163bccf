#!/bin/bash
set +e
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== config 5: 10M x 128 sharded over 8 GPUs (1.25M per shard), 100k queries, one all-gather"
timeout 900 $TR --master-port 29521 scripts/sharded_check.py --points 10000000 --queries 100000 --no-oracle --bench 5 > gpurun_out/sharded_8gpu_10M.json 2> gpurun_out/sharded_8gpu_10M.log; echo "rc=$?"; cat gpurun_out/sharded_8gpu_10M.json; grep -E "DIAG|Error|error" gpurun_out/sharded_8gpu_10M.log | head -5
echo "== bench.py --gpus 8 (replicas, weak scaling)"
timeout 900 $TR --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.log; echo "rc=$?"; cat gpurun_out/bench_8gpu.json; tail -3 gpurun_out/bench_8gpu.log
