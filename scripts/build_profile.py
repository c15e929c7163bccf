"""GPU build of n x dim sift-shaped points through the C ABI (used under ncu to get the per-kernel time split).

  python scripts/build_profile.py [n] [dim] [M] [ef_construction]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "instant-distance_b200", "python"))
from instant_distance_b200 import _abi  # noqa: E402
from tests import datagen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32
efc = int(sys.argv[4]) if len(sys.argv) > 4 else 100
pts = datagen.sift_shaped(n, dim, 1)
t = time.time()
ix, ids = _abi.Index.build(pts, seed=7, M=M, ef_construction=efc)
ix.sync()
print(f"build {n} x {dim}, M={M}, ef_construction={efc}: {time.time() - t:.2f}s", flush=True)
