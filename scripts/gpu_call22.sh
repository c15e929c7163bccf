#!/bin/bash
# Call 22: K1 variants — shared-memory landing zone (IDB_VARIANT=5), + speculative rows under the visited probes (6),
# bitmap flavour of the wide visited tier (IDB_VIS_BITMAP=1).  Parity first, then the sweep on the headline workload.
set +e
mkdir -p gpurun_out
echo "== smoke (defaults)"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"; echo "rc=$?"
echo "== parity, IDB_VARIANT=5 IDB_VIS_BITMAP=1"
IDB_VARIANT=5 IDB_VIS_BITMAP=1 timeout 900 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_bf16.py -x -q -m gpu > gpurun_out/pytest_v5_bm.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_v5_bm.log
echo "== parity, IDB_VARIANT=6"
IDB_VARIANT=6 timeout 900 python -m pytest tests/test_gpu_search_parity.py tests/test_gpu_bf16.py -x -q -m gpu > gpurun_out/pytest_v6.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_v6.log
echo "== sweep 1M x 128"
timeout 900 python scripts/tune_search.py --n 1000000 --steps 6 --configs "IDB_OPT=0;IDB_VIS_BITMAP=1;IDB_VARIANT=5;IDB_VARIANT=5,IDB_VIS_BITMAP=1;IDB_VARIANT=6;IDB_VARIANT=6,IDB_VIS_BITMAP=1;IDB_VARIANT=5,IDB_VIS_BITMAP=1,IDB_OPT=8;IDB_OPT=0" > gpurun_out/tune_call22.jsonl 2> gpurun_out/tune_call22.log; echo "rc=$?"; cat gpurun_out/tune_call22.jsonl; tail -3 gpurun_out/tune_call22.log
