"""Builds the bench graph with the oracle (reference algorithm) on this box's cores and stores it under gpurun_cache/
so GPU calls need not spend box time on the CPU build.  Usage: python scripts/prebuild_graph.py [bench.py args]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--ef", type=int, default=100)
ap.add_argument("--efc", type=int, default=100)
ap.add_argument("--M", type=int, default=32)
ap.add_argument("--seed", type=int, default=20260923)
ap.add_argument("--data", default="sift")
args = ap.parse_args()
args.graph, args.no_cache = "oracle", False
pts, _ = bench.make_workload(args)
bench.obtain_graph(args, pts, 0)
print("cached at", bench.graph_cache_path(args))
