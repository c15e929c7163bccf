"""bench.py's N=1 line assembled on a CPU box: the measuring legs are replaced by canned results (they need a GPU), everything
else — argument handling, secondary legs, failure isolation of the secondary legs, the one JSON line the driver parses — runs
for real."""
import json
import sys

import pytest

import bench

ROOF = {"bound": "hbm", "achieved": 6000.0, "peak": 6566.7, "unit": "GB/s", "frac": 0.91, "traffic": None, "kernel": "K1", "kernel_ms": 4.6}
CLOCKS = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 9}


def fake_search(a, rank, local_rank, world, full=True):
    res = {"value": 2.0e6, "ms_per_step": 5.0, "recall_at_10": 0.97 if a.data == "sift" else 0.32, "ef_search": a.ef, "gpu_launches": 2 * a.steps,
           "clocks": CLOCKS, "retried_per_launch": 0, "roofline": dict(ROOF)}
    if full:
        res["e2e"] = {"value": 1.9e6, "unit": "queries/s", "h2d_bytes_per_step": a.batch * a.dim * 4, "d2h_bytes_per_step": a.batch * 84,
                      "callers": a.callers}
        res["cpu_baseline"] = {"value": 1.0e4, "unit": "queries/s", "cores": 16, "kind": "port", "sample": "canned"}
    return res


def fake_sharded(a, rank, local_rank, world, full=True):
    return {"value": 2.5e5, "unit": "queries/s", "n_gpus": world, "ms_per_step": 400.0, "recall_at_10": 0.989, "merged_eq_protocol": True,
            "gpu_launches": 380, "clocks": CLOCKS, "roofline": dict(ROOF), "build_s_per_rank": 20.0}


def fake_build(a, local_rank):
    assert (a.n, a.dim, a.M, a.efc) == (2_000_000, 300, 24, 200)  # BASELINE configs[2]
    return {"metric": "GPU Builder::build throughput", "value": 1.7e5, "unit": "points/s", "seconds": [11.7], "recall_at_10_of_built_graph": 0.89,
            "ef_search": 100, "config": {"workload": "canned"}, "search": {"value": 5.0e5}, "cpu_baseline": {"value": 8.9e3}}


@pytest.fixture
def cpu_bench(monkeypatch):
    import torch
    from instant_distance_b200 import _abi

    class FakeLib:
        def idb_device_count(self):
            return 1

    monkeypatch.setattr(_abi, "lib", lambda: FakeLib())
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_: None)
    monkeypatch.setattr(bench, "leg_search", fake_search)
    monkeypatch.setattr(bench, "leg_sharded", fake_sharded)
    monkeypatch.setattr(bench, "leg_build", fake_build)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)

    def run(argv, capsys):
        monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
        bench.main()
        out = capsys.readouterr().out.strip().splitlines()
        assert len(out) == 1, "exactly ONE line on stdout"
        return json.loads(out[0])

    return run


def test_default_line_has_every_contract_key(cpu_bench, capsys):
    line = cpu_bench(["--steps", "7", "--warmup", "1"], capsys)
    assert line["metric"].startswith("batched QPS at recall@10") and line["unit"] == "queries/s" and line["n_gpus"] == 1
    assert line["steps"] == 7 and line["warmup"] == 3  # timing rule: at least 3 warm-up steps
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f32" and line["data"] == "synthetic"
    for key in ("value", "ms_per_step", "scaling", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "recall_at_10"):
        assert key in line, key
    assert {"h2d_bytes_per_step", "d2h_bytes_per_step", "value", "unit"} <= set(line["e2e"])
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["sharded"]["merged_eq_protocol"] is True and line["uniform"]["recall_at_10"] == 0.32
    assert line["build"]["unit"] == "points/s" and "search" in line["build"] and "cpu_baseline" in line["build"]


def test_a_failing_secondary_leg_does_not_take_the_headline_down(cpu_bench, capsys, monkeypatch):
    def boom(a, local_rank):
        raise RuntimeError("out of memory (canned)")

    monkeypatch.setattr(bench, "leg_build", boom)
    line = cpu_bench([], capsys)
    assert line["value"] == 2.0e6 and "error" in line["build"] and "canned" in line["build"]["error"]
    assert "error" not in line["sharded"] and "error" not in line["uniform"]


def test_skip_secondary_prints_the_headline_only(cpu_bench, capsys):
    line = cpu_bench(["--skip-secondary"], capsys)
    assert "sharded" not in line and "uniform" not in line and "build" not in line and line["roofline"]["bound"] == "hbm"
