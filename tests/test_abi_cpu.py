"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/instant_distance_b200.h declares, and FAILS LOUDLY (no CPU fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, _has_gpu


def _abi():
    from instant_distance_b200 import _abi

    return _abi


def test_library_is_built_in_tree():
    assert os.path.exists(_abi().LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported():
    abi = _abi()
    header = open(os.path.join(ROOT, "include", "instant_distance_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(idb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "header declares no functions?"
    L = abi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported by the .so"
    assert sorted(abi.SYMBOLS) == declared, "python binding and header disagree on the symbol list"


def test_struct_layouts_match_header():
    import ctypes as C

    abi = _abi()
    assert C.sizeof(abi.Params) == 64  # 4*4 + 8 + 5*4 + 4 (storage) + 2 pointers
    assert C.sizeof(abi.Info) == 8 + 16 + 32 * 8 + 8
    p = abi.default_params()
    assert (p.M, p.ef_construction, p.ef_search, p.heuristic, p.extend_candidates, p.keep_pruned) == (32, 100, 100, 1, 0, 1)
    assert abs(p.ml - 1.0 / np.log(32.0)) < 1e-7  # lib.rs:107


def test_product_does_not_touch_the_oracle():
    """The shipped package must not import/link anything under oracle/ (the oracle is the checker only)."""
    pkg = os.path.join(ROOT, "instant-distance_b200")
    for base, _, files in os.walk(pkg):
        if os.sep + "lib" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in text and "hnsw_oracle" not in text and "from oracle" not in text, os.path.join(base, f)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_fails_loudly_without_a_device():
    abi = _abi()
    assert abi.lib().idb_device_count() == 0
    with pytest.raises(abi.IdbError) as e:
        abi.distance(np.ones(4), np.zeros(4))
    assert e.value.status == abi.ERR_CUDA and "no CPU fallback" in str(e.value)
    pts = np.zeros((4, 4), dtype=np.float32)
    zero = np.full((4, 64), 0xFFFFFFFF, dtype=np.uint32)
    with pytest.raises(abi.IdbError) as e:
        abi.Index.from_graph(pts, zero, [], 32)
    assert e.value.status == abi.ERR_CUDA


def test_argument_validation_needs_no_device():
    abi = _abi()
    import ctypes as C

    h = C.c_void_p()
    st = abi.lib().idb_index_from_graph_f32(None, 0, 0, 32, 100, None, 0, None, None, 0, C.byref(h))
    assert st == abi.ERR_INVALID_ARG and b"dim" in abi.lib().idb_last_error()
    st = abi.lib().idb_index_from_graph_f32(None, 0, 8, 1000, 100, None, 0, None, None, 0, C.byref(h))
    assert st == abi.ERR_INVALID_ARG
