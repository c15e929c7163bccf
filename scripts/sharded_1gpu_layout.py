"""Scaling denominator for the sharded config: ONE GPU holding the same 8-sub-index layout of 10M points and
searching every query on every sub-index.  Prints the time per 100k-query batch (sum over the 8 sub-indexes)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "instant-distance_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from instant_distance_b200 import _abi, sharded  # noqa: E402
from tests import datagen  # noqa: E402

n, world, nq, k, ef = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 8, 100_000, 10, 100
pts = datagen.sift_shaped(n, 128, 5)
q = torch.from_numpy(datagen.sift_shaped(nq, 128, 6)).cuda()
shards = []
t0 = time.time()
for r in range(world):
    lo, hi = sharded.shard_range(n, r, world)
    ix, ids = _abi.Index.build(pts[lo:hi], seed=100 + r)
    ix.set_id_map(sharded.global_id_map(ids, lo))
    ix.set_profiling(True)
    shards.append(ix)
build_s = time.time() - t0
d_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
d_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
d_l = torch.empty((nq,), dtype=torch.int32, device="cuda")
per = []
for rep in range(3):
    ms = 0.0
    for ix in shards:
        ix.search_device(q.data_ptr(), nq, ef, k, d_ids.data_ptr(), d_d.data_ptr(), d_l.data_ptr())
        ms += ix.last_kernel_ms()[0]
    per.append(ms)
print(json.dumps({"layout": f"{world} sub-indexes of {n // world} points on ONE GPU", "n": n, "queries": nq, "build_s_total": round(build_s, 1),
                  "search_ms_per_batch_sum_over_shards": float(np.mean(per[1:])), "qps": nq / (np.mean(per[1:]) / 1e3)}), flush=True)
