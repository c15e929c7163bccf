"""Sweeps the GPU build knobs (env) at n x dim sift-shaped; prints build seconds and recall@10 at ef=100 per setting."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
configs = (sys.argv[2] if len(sys.argv) > 2 else
           "IDB_BUILD_STAGE=1;IDB_BUILD_STAGE=0;IDB_BUILD_MAXBATCH=8192;IDB_BUILD_MAXBATCH=16384;IDB_BUILD_MAXBATCH=16384,IDB_BUILD_GROWTH=8;"
           "IDB_BUILD_MAXBATCH=2048")
pts = datagen.sift_shaped(n, 128, 1)
rq = datagen.sift_shaped(1000, 128, 999)
for cfg in configs.split(";"):
    for k in ("IDB_BUILD_STAGE", "IDB_BUILD_MAXBATCH", "IDB_BUILD_GROWTH"):
        os.environ.pop(k, None)
    os.environ.update(dict(x.split("=") for x in cfg.split(",") if x))
    t = time.time()
    ix, ids = _abi.Index.build(pts, seed=7)
    ix.sync()
    bs = time.time() - t
    p, _, _ = ix.export_graph()
    truth = bench.brute_force_topk_torch(torch.from_numpy(p).cuda(), rq, 10)[0]
    got, _, _ = ix.search(rq, ef_search=100, k=10)
    print(json.dumps({"config": cfg, "n": n, "build_s": round(bs, 2), "points_per_s": n / bs, "recall_at_10": bench.recall_at_k(got, truth)}), flush=True)
    ix.close()
    torch.cuda.empty_cache()
