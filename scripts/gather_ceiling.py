"""Gather ceiling next to K1 (measurement only): random 512-byte row gathers in K1's launch shape, with 0 / 3 / 1 independent
16-row batches between dependent steps.  Prints GB/s per setting for a 1M x 128 f32 index."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "instant-distance_b200", "python"))
import numpy as np  # noqa: E402

from instant_distance_b200 import _abi  # noqa: E402
from tests import datagen  # noqa: E402

MIX = "--mix" in sys.argv  # also measure with K1's visited-style atomic stream riding along (call 26)
n = 1_000_000
pts = datagen.sift_shaped(n, 128, 1)
zero = np.full((n, 64), 0xFFFFFFFF, dtype=np.uint32)
for ctas in (("4",) if MIX else ("4", "3", "2")):
    os.environ["IDB_CTAS_PER_SM"] = ctas
    ix = _abi.Index.from_graph(pts, zero, [], 32)
    for chain in (0, 6, 3, 1):
        ms, by = ix.gather_bench(n_items=10000, batches=288, chain=chain, reps=3)
        print(json.dumps({"ctas_per_sm": int(ctas), "warps_per_sm": int(ctas) * 4, "chain": chain, "ms": ms, "GBps": by / (ms / 1e3) / 1e9,
                          "frac_of_6572": by / (ms / 1e3) / 1e9 / 6572.5}), flush=True)
    if MIX and "--sizes" in sys.argv:  # table-size study: does a visited structure that fits L2 stop costing throughput?
        for words, persist in ((4096, 0), (4096, 1), (8192, 0), (8192, 1), (6144, 1), (31360, 1)):
            os.environ["IDB_DEBUG_BM_WORDS"] = str(words)
            os.environ["IDB_DEBUG_BM_PERSIST"] = str(persist)
            for mode in (1, 5):
                ms, by = ix.gather_mix_bench(n_items=10000, batches=288, chain=3, reps=3, atomics=21, mode=mode)
                print(json.dumps({"table_words_per_warp": words, "l2_persist": persist, "mix_mode": mode, "atomics_per_batch": 21, "ms": ms,
                                  "row_GBps": by / (ms / 1e3) / 1e9, "frac_of_6572": by / (ms / 1e3) / 1e9 / 6572.5}), flush=True)
    elif MIX:
        for mode, chain, atomics in ((1, 3, 21), (3, 3, 21), (4, 3, 21), (5, 3, 21), (1, 3, 0), (4, 3, 32), (1, 3, 8)):
            ms, by = ix.gather_mix_bench(n_items=10000, batches=288, chain=chain, reps=3, atomics=atomics, mode=mode)
            print(json.dumps({"ctas_per_sm": int(ctas), "mix_mode": mode, "atomics_per_batch": atomics, "chain": chain, "ms": ms,
                              "row_GBps": by / (ms / 1e3) / 1e9, "frac_of_6572": by / (ms / 1e3) / 1e9 / 6572.5}), flush=True)
    ix.close()
