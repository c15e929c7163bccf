"""Host-side helpers of bench.py (no GPU): the algorithmic-bytes formula of SURVEY §8d, recall@k, and the committed ncu traffic figure."""
import json
import os

import numpy as np

import bench


def test_algorithmic_bytes_formula():
    # B(q) = vec + n_expand_upper * M*4 + n_dist_upper * vec + n_expand_zero * 2M*4 + n_dist_zero * vec + k*8
    c = np.array([[10, 300, 100, 5000], [0, 1, 1, 1]], dtype=np.uint64)
    got = bench.algorithmic_bytes(c, dim=128, M=32, k=10)
    vec = 128 * 4
    assert got[0] == vec + 10 * 128 + 300 * vec + 100 * 256 + 5000 * vec + 80
    assert got[1] == vec + 0 + vec + 256 + vec + 80


def test_recall_at_k():
    truth = np.array([[1, 2, 3], [4, 5, 6]])
    assert bench.recall_at_k(np.array([[3, 2, 1], [4, 9, 9]]), truth, k=3) == (3 + 1) / 6
    assert bench.recall_at_k(truth, truth, k=2) == 1.0


def test_committed_ncu_traffic_is_consistent():
    traffic, source = bench.ncu_traffic()
    d = json.load(open(os.path.join(bench.ROOT, "profiles", "k1_ncu_traffic.json")))
    assert traffic == d["dram_bytes_per_launch"] and abs(d["dram_bytes_read"] + d["dram_bytes_write"] - traffic) < 1e6
    assert os.path.exists(os.path.join(bench.ROOT, source.split(" ")[0]))  # the summary the figure was read from is committed


def test_workload_config_names_the_baseline_config():
    class A:
        n, dim, data, M, efc, ef, batch, gpus = 1_000_000, 128, "sift", 32, 100, 100, 10_000, 1

    cfg = bench.workload_config(A, "GPU Builder::build")
    assert "1000000 x 128" in cfg["workload"] and "ef_search=100" in cfg["workload"] and cfg["parallelism"] == "1 GPU"
