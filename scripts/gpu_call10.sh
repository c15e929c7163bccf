#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== plain build timing"; python scripts/build_profile.py 1000000 128
echo "== ncu launch list of a 300k build"; timeout 1200 ncu --kernel-name-base mangled -k regex:idb --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/build_launches.csv python scripts/build_profile.py 300000 128 > gpurun_out/build_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/build_ncu.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/build_launches.csv')) if len(r)>5]
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[rows.index(hdr)+1:]:
    try: agg[r[ki].split('(')[0][:70]].append(float(r[vi].replace(',','')))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1])): print("%-72s n=%5d total_ms=%9.2f share=%5.1f%% avg_us=%8.1f"%(k,len(v),sum(v)/1e6,100*sum(v)/tot,sum(v)/len(v)/1e3))
PY
echo "== ncu full of one late relink + select_new + insert_search"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:relink_kernel -s 150 -c 1 -o gpurun_out/prof_relink python scripts/build_profile.py 300000 128 > gpurun_out/ncu_relink.log 2>&1; echo "rc=$?"
