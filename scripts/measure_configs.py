"""Measures the other BASELINE.json configs on one GPU (build + search through the C ABI); prints one JSON line each.
  config0: 10k x 32, M=16, ef=100, 1k queries           (GPU build + search, recall vs brute force)
  config2: 2M x 300, ef_construction=200, M=24, batch 10k (GPU Builder::build + search)
  uniform: 1M x 128 uniform-random (north_star's data) at ef = 100 and the ef that reaches recall@10 >= 0.95"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "instant-distance_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from instant_distance_b200 import _abi  # noqa: E402
from tests import datagen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--which", default="config0,uniform,config2")
args = ap.parse_args()


def run(name, gen, n, dim, M, efc, efs, nq, steps=5, storage="f32"):
    pts = gen(n, dim, 1)
    t0 = time.time()
    ix, ids = _abi.Index.build(pts, M=M, ef_construction=efc, ef_search=efs[0], seed=7, storage=storage)
    ix.sync()
    build_s = time.time() - t0
    p, zero, upper = ix.export_graph()
    rq = gen(1000, dim, 999)
    truth = bench.brute_force_topk_torch(torch.from_numpy(p).cuda(), rq, 10)[0]
    torch.cuda.empty_cache()
    ix.set_profiling(True)
    del pts
    out = {"config": name, "storage": storage, "n": n, "dim": dim, "M": M, "ef_construction": efc, "build_s": round(build_s, 2),
           "build_points_per_s": n / build_s, "searches": []}
    qs = [torch.from_numpy(gen(nq, dim, 7000 + s)).cuda() for s in range(steps + 2)]
    d_ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
    d_dist = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    d_len = torch.empty((nq,), dtype=torch.int32, device="cuda")
    for ef in efs:
        got, _, _ = ix.search(rq, ef_search=ef, k=10)
        rec = bench.recall_at_k(got, truth)
        ms = []
        for s in range(steps + 2):
            ix.search_device(qs[s].data_ptr(), nq, ef, 10, d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr())
            t, _ = ix.last_kernel_ms()
            if s >= 2:
                ms.append(t)
        byts = float(bench.algorithmic_bytes(ix.last_counters(nq), dim, M, 10, elem=4 if storage == "f32" else 2).sum())
        cnt = ix.last_counters(nq).mean(0).tolist()
        out["searches"].append({"ef_search": ef, "recall_at_10": rec, "kernel_ms": float(np.mean(ms)), "qps": nq / (np.mean(ms) / 1e3),
                                "GBps": byts / (np.mean(ms) / 1e3) / 1e9, "frac_of_6572": byts / (np.mean(ms) / 1e3) / 1e9 / 6572.5,
                                "counters_mean": cnt, "retried": ix.last_retried(0)})
        if rec >= 0.95 and (name == "uniform" or name.startswith("config4")) and ef > efs[0]:
            break
    print(json.dumps(out), flush=True)
    ix.close()
    torch.cuda.empty_cache()


for w in args.which.split(","):
    if w == "config0":
        run("config0 10k x 32 M=16", datagen.uniform, 10_000, 32, 16, 100, [100], 1000)
    elif w == "uniform":
        run("uniform", datagen.uniform, 1_000_000, 128, 32, 100, [100, 200, 300, 400, 512], 10_000)
    elif w == "config2":
        run("config2 2M x 300 M=24 efc=200", datagen.uniform, 2_000_000, 300, 24, 200, [100, 200], 10_000)
    elif w == "config4":
        run("config4 5M x 768 bf16 ef=128 batch 64k", datagen.sift_shaped, 5_000_000, 768, 32, 100, [128, 160, 200, 256], 65_536, steps=3, storage="bf16")
    elif w == "config4s":
        run("config4 (1M subset) 1M x 768 bf16 ef=128 batch 64k", datagen.sift_shaped, 1_000_000, 768, 32, 100, [128], 65_536, steps=3, storage="bf16")
    elif w == "config2s":
        run("config2 (sift-shaped) 2M x 300 M=24 efc=200", datagen.sift_shaped, 2_000_000, 300, 24, 200, [100, 200], 10_000)
