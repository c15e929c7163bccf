#!/usr/bin/env python
"""How often does the product's canonical fp32 summation order change RESULTS relative to the reference's own order?

The product (CUDA kernels + oracle) computes squared L2 in one canonical order for every dim; the reference's only
f32-vector Point, FloatArray (instant-distance-py/src/lib.rs:378-421, dim fixed at 300), sums in an 8-lane AVX2 order.  Both are
correctly-rounded-per-operation fp32, so individual distances differ by a few ulp (<= 1e-4 relative, pinned in
tests/test_oracle_reference_pins.py) — but a traversal takes thousands of comparisons, and a flipped near-tie can change the ids.
This script measures that: ONE graph (built by the oracle with the reference's order, i.e. what the Rust crate would build up to
rayon's scheduling), searched twice — distances in the reference's order vs in the canonical order — same queries, same ef.
CPU only (oracle = test infrastructure); writes one JSON line per (data, n).

  python scripts/summation_order_gap.py [--n 100000] [--nq 10000] [--threads 8] [--out profiles/r02_summation_order_gap.jsonl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import datagen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=300)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--ef", type=int, default=100)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--data", default="sift,uniform")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    assert a.dim % 8 == 4, "the reference's order is defined for dim % 8 == 4 (py:388)"
    for data in a.data.split(","):
        gen = datagen.sift_shaped if data == "sift" else datagen.uniform
        pts, q = gen(a.n, a.dim, 1), gen(a.nq, a.dim, 2)
        t = time.time()
        ix, _ = O.build(pts, seed=7, threads=a.threads, metric=O.METRIC_REFERENCE_AVX2)
        build_s = time.time() - t
        g = ix.export()
        ref = O.from_graph(g, metric=O.METRIC_REFERENCE_AVX2)
        can = O.from_graph(g, metric=O.METRIC_CANONICAL)
        r_ids, r_dist, r_len, r_cnt = ref.search(q, ef_search=a.ef, k=a.ef, threads=a.threads, counters=True)
        c_ids, c_dist, c_len, c_cnt = can.search(q, ef_search=a.ef, k=a.ef, threads=a.threads, counters=True)
        top10_list = (r_ids[:, :10] != c_ids[:, :10]).any(1)
        top10_set = np.array([set(x[:10].tolist()) != set(y[:10].tolist()) for x, y in zip(r_ids, c_ids)])
        full_list = (r_ids != c_ids).any(1)
        trav = (r_cnt != c_cnt).any(1)
        same = r_ids == c_ids
        rel = np.abs(r_dist[same].astype(np.float64) - c_dist[same]) / np.maximum(r_dist[same], 1e-30)
        truth, _ = O.bruteforce(g.points, q[:1000], 10, metric=O.METRIC_CANONICAL, threads=a.threads)
        rec = lambda ids: float(np.mean([len(set(x[:10].tolist()) & set(t.tolist())) / 10 for x, t in zip(ids[:1000], truth)]))  # noqa: E731
        line = {
            "what": "reference AVX2 summation order (py:390-411) vs canonical order, same graph, same queries",
            "data": data, "n": a.n, "dim": a.dim, "nq": a.nq, "ef_search": a.ef, "graph": f"oracle build, reference order, {a.threads} threads, {build_s:.0f}s",
            "queries_top10_ids_differ_as_list": float(top10_list.mean()), "queries_top10_ids_differ_as_set": float(top10_set.mean()),
            "queries_full_ef_list_differs": float(full_list.mean()), "queries_traversal_counters_differ": float(trav.mean()),
            "max_rel_distance_gap_same_id": float(rel.max()), "recall10_reference_order": rec(r_ids), "recall10_canonical_order": rec(c_ids),
        }
        print(json.dumps(line), flush=True)
        if a.out:
            with open(os.path.join(ROOT, a.out) if not os.path.isabs(a.out) else a.out, "a") as f:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
