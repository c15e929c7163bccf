#!/bin/bash
# Call 29: bucket set with the row's slot claims overlapped and the step limit raised: parity, bench, ncu --set full, launch list.
set +e
mkdir -p gpurun_out
echo "== parity (search + bf16 + sharded + python module)"
timeout 900 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_bf16.py tests/test_gpu_sharded.py tests/test_python_module.py -x -q -m gpu > gpurun_out/pytest_call29.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_call29.log
echo "== bench (defaults)"; timeout 900 python bench.py > gpurun_out/bench_1M.json 2> gpurun_out/bench_1M.log; echo "rc=$?"; cat gpurun_out/bench_1M.json
echo "== ncu full of K1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^search_kernel -s 8 -c 2 -f -o gpurun_out/k1_buckets2 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_k1_buckets2.log 2>&1; echo "rc=$?"; grep -c "Profiling" gpurun_out/ncu_k1_buckets2.log
echo "== ncu launch list of the bench command"; timeout 900 ncu --kernel-name-base mangled -k regex:idb --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
agg=collections.defaultdict(list)
for r in rows[rows.index(hdr)+1:]:
    try: agg[(r[ki].split('(')[0][:60], r[gi])].append(float(r[vi].replace(',','')))
    except: pass
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])): print("%-62s grid=%-14s n=%4d total_ms=%9.3f avg_us=%9.1f"%(k[0],k[1],len(v),sum(v)/1e6,sum(v)/len(v)/1e3))
PY
