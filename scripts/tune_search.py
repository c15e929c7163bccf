"""Sweeps the K1 tuning knobs (env vars read at index creation) on one graph; prints kernel ms per config.
Usage (GPU box): python scripts/tune_search.py --n 1000000 [--configs "OPT,VIS,L2P,CTAS;..."]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--ef", type=int, default=100)
ap.add_argument("--efc", type=int, default=100)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--seed", type=int, default=20260923)
ap.add_argument("--data", default="sift")
ap.add_argument("--graph", default="gpu")
ap.add_argument("--batch", type=int, default=10000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--configs", default="IDB_OPT=0;IDB_VIS_TIER=1;IDB_VIS_TIER=0;IDB_B16_BYTES=16384;IDB_B16_BYTES=32768;IDB_L2_PERSIST=0")
args = ap.parse_args()
args.no_cache = False

import torch  # noqa: E402

from instant_distance_b200 import _abi  # noqa: E402

gen = bench.generator(args.data)
pts = gen(args.n, args.dim, 1)
p, zero, upper, _ = bench.obtain_graph(pts, args.n, args.dim, args.data, 1, args.M, args.efc, args.ef, args.seed, 0, use_abi=True)
print(json.dumps({"graph": bench.GRAPH_NOTE, "n": args.n, "data": args.data, "ef": args.ef}), flush=True)
del pts
qs = [torch.from_numpy(gen(args.batch, args.dim, 7000 + s)).cuda() for s in range(args.steps + 2)]
k = 10
d_ids = torch.empty((args.batch, k), dtype=torch.int32, device="cuda")
d_dist = torch.empty((args.batch, k), dtype=torch.float32, device="cuda")
d_len = torch.empty((args.batch,), dtype=torch.int32, device="cuda")
ref_ids = None
results = []
KNOBS = ["IDB_OPT", "IDB_VIS_MULT", "IDB_CTAS_PER_SM", "IDB_VARIANT", "IDB_VIS_TIER", "IDB_B16_BYTES", "IDB_L2_PERSIST"]
for cfg in args.configs.split(";"):
    for kname in KNOBS:
        os.environ.pop(kname, None)
    kv = dict(x.split("=") for x in cfg.split(",") if x)
    os.environ.update(kv)
    ix = _abi.Index.from_graph(p, zero, upper, args.M, args.ef)
    ix.set_profiling(True)
    ms = []
    for s in range(args.steps + 2):
        ix.search_device(qs[s].data_ptr(), args.batch, args.ef, k, d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr())
        t, _ = ix.last_kernel_ms()
        if s >= 2:
            ms.append(t)
    ix.sync()  # the retry pass runs after the timed kernel on the library's own stream
    ids = d_ids.cpu().numpy()
    if ref_ids is None:
        ref_ids = ids
    same = bool((ids == ref_ids).all())
    cnt = ix.last_counters(args.batch)
    byts = float(bench.algorithmic_bytes(cnt, args.dim, args.M, k).sum())
    r = {"config": cfg, "kernel_ms": float(np.mean(ms)), "min_ms": float(np.min(ms)), "GBps": byts / (np.mean(ms) / 1e3) / 1e9,
         "frac": byts / (np.mean(ms) / 1e3) / 1e9 / bench.measured_peaks()[0], "qps": args.batch / (np.mean(ms) / 1e3), "same_ids": same,
         "retried": ix.last_retried(0), "failed_after_retry": ix.last_failures(0), "n_dist_zero_mean": float(cnt[:, 3].mean()), "n_dist_zero_p999": float(np.quantile(cnt[:, 3], 0.999)),
         "n_dist_zero_max": int(cnt[:, 3].max()), "n_expand_zero_mean": float(cnt[:, 2].mean())}
    print(json.dumps(r), flush=True)
    results.append(r)
    ix.close()
    torch.cuda.empty_cache()
