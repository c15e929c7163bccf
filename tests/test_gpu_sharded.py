"""Sharded search on >= 2 GPUs: runs scripts/sharded_check.py under torchrun, one rank per GPU.
Checks (inside the script): every shard's GPU search == the oracle on that shard's graph; the fused path (K1 epilogue
pack -> ONE ncclAllGather -> merge kernel) == the host statement of the protocol; identical results on all ranks."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpu_count():
    from instant_distance_b200 import _abi

    return _abi.lib().idb_device_count()


@pytest.mark.skipif("_gpu_count() < 2", reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_sharded_search_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "sharded_check.py"), "--points", "60000", "--queries", "3000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["local_eq_oracle"] and res["fused_eq_protocol"] and res["world"] == 2
