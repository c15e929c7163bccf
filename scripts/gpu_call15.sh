#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench 1M (f32, check no regression)"; timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_1M_b.json 2> gpurun_out/bench_1M_b.log; echo "rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_1M_b.json'));print(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['config']['graph'],d['recall_at_10'])"
