"""Tiny parity case meant to be run under compute-sanitizer (memcheck / racecheck) on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "instant-distance_b200", "python"))
import numpy as np  # noqa: E402

from instant_distance_b200 import _abi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import datagen  # noqa: E402

for (n, dim, M, ef, gen) in [(1500, 24, 32, 50, datagen.uniform), (800, 3, 32, 10, datagen.grid_ties), (600, 300, 24, 20, datagen.uniform)]:
    pts = gen(n, dim, 1)
    ix, _ = O.build(pts, seed=1, M=M)
    g = ix.export()
    gpu = _abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
    q = gen(64, dim, 2)
    ids, dist, lens = gpu.search(q, ef_search=ef, k=ef)
    o = ix.search(q, ef_search=ef, k=ef)
    assert (ids == o[0]).all() and dist.tobytes() == o[1].tobytes() and (lens == o[2]).all(), (n, dim, M, ef)
    gpu.close()
# long rows (query in shared memory), a small GPU build (KA / K2 / K2' + the shared table pool), and a 3-shard world-of-one sharded search
pts = datagen.uniform(400, 1100, 5)
ix, _ = O.build(pts, seed=2)
g = ix.export()
gpu = _abi.Index.from_graph(g.points, g.zero, g.upper, g.M)
q = datagen.uniform(16, 1100, 6)
a, b = gpu.search(q, ef_search=20, k=20), ix.search(q, ef_search=20, k=20)
assert (a[0] == b[0]).all() and a[1].tobytes() == b[1].tobytes()
gpu.close()
built, ids = _abi.Index.build(datagen.uniform(1500, 16, 7), seed=3, insert_batch=64)
ref, _ = O.build(datagen.uniform(1500, 16, 7), seed=3)
got = built.search(datagen.uniform(32, 16, 8), ef_search=50, k=10)
assert (got[2] == 50).all()
built.close()
print("sanitize case ok")
