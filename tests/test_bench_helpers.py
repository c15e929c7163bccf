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


def test_configs_name_the_baseline_configs_and_match_across_arms():
    class A:
        n, dim, data, M, efc, ef, batch, gpus, shard_n, shard_batch = 1_000_000, 128, "sift", 32, 100, 100, 10_000, 1, 1_250_000, 100_000

    cfg = bench.search_config(A, 1, "headline")
    assert "1000000 x 128" in cfg["workload"] and "ef_search=100" in cfg["workload"] and cfg["parallelism"] == "1 GPU"
    # no timings / cache markers inside config: both arms of a run print the identical object
    assert cfg == bench.search_config(A, 1, "headline") and "cached" not in json.dumps(cfg) and "(s)" not in cfg["graph"]
    sh = bench.sharded_config(A, 8)
    assert "8 x 1250000 = 10000000 x 128" in sh["workload"] and "batch=100000" in sh["workload"] and "ONE ncclAllGather" in sh["parallelism"]


def test_cpu_baseline_protocol_counts_only_the_timed_passes():
    calls = []
    qps, dt = bench.cpu_qps(lambda b: calls.append(len(b)), [np.zeros((7, 2))] * 6, 2, 4)
    assert calls == [7] * 6 and dt > 0 and abs(qps - 28 / dt) < 1e-6 * qps
