#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1200 python scripts/tune_search.py --n 1000000 --configs "IDB_OPT=0;IDB_OPT=8;IDB_OPT=10;IDB_OPT=8,IDB_VIS_MULT=2;IDB_OPT=0" > gpurun_out/tune19_1M.jsonl 2> gpurun_out/tune19_1M.log; echo "rc=$?"; cat gpurun_out/tune19_1M.jsonl; tail -2 gpurun_out/tune19_1M.log
