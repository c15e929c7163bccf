"""Generates the committed golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference is Rust and cannot be built or imported in this image, and its own tests hold no fixed-seed expected-id
vectors; the exact answers it does pin (tests/all.rs `map`, recall bars, the layer schedule) are asserted against the oracle in
tests/test_oracle_reference_pins.py.  These fixtures freeze the ORACLE's outputs on small seeded inputs instead, so that
  * the CPU suite notices any drift of the oracle itself (tests/test_golden.py, not gpu), and
  * the GPU suite compares the CUDA path with committed numbers, not only with the oracle built from today's sources.
Inputs are regenerated from seeds (tests/datagen.py); only the expected outputs and the small graphs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import datagen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (name, data generator, n, dim, M, ef_construction, build seed, ef_search list)
SEARCH_CASES = [
    ("uniform_2000x16_M32", "uniform", 2000, 16, 32, 100, 7, [1, 10, 100]),
    ("grid_ties_1500x3_M32", "grid_ties", 1500, 3, 32, 100, 5, [10, 100]),
    ("sift_3000x128_M32", "sift_shaped", 3000, 128, 32, 100, 3, [100]),
    ("uniform_1200x40_M16", "uniform", 1200, 40, 16, 60, 9, [64]),
]
NQ = 48


def case_inputs(gen, n, dim):
    f = getattr(datagen, gen)
    return f(n, dim, 1234), f(NQ, dim, 4321)


def main():
    for name, gen, n, dim, M, efc, seed, efs in SEARCH_CASES:
        pts, q = case_inputs(gen, n, dim)
        ix, ids = O.build(pts, seed=seed, M=M, ef_construction=efc, threads=1)  # sequential = deterministic (core:313-318)
        g = ix.export()
        out = {"ids_map": ids, "zero": g.zero, "n_upper": np.int64(len(g.upper))}
        for i, u in enumerate(g.upper):
            out[f"upper{i}"] = u
        for ef in efs:
            r_ids, r_dist, r_len, r_cnt = ix.search(q, ef_search=ef, k=min(ef, 16), counters=True)
            out[f"ef{ef}_ids"], out[f"ef{ef}_dist"], out[f"ef{ef}_len"], out[f"ef{ef}_cnt"] = r_ids, r_dist, r_len, r_cnt.astype(np.uint32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("ef") or k == "zero"})


if __name__ == "__main__":
    main()
