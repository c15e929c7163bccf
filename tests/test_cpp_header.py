"""The C++17 host mirror (instant-distance_b200/cpp/instant_distance.hpp) compiles, links against the C ABI library and
behaves like the reference API: the `map` test (tests/all.rs:9-39) on GPU, loud failure without a device."""
import os
import subprocess

import pytest

from tests.conftest import ROOT, _has_gpu

SRC = os.path.join(ROOT, "tests", "cpp", "test_hpp.cpp")
LIBDIR = os.path.join(ROOT, "instant-distance_b200", "lib")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "test_hpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", out, SRC, "-L" + LIBDIR, "-linstant_distance_b200",
                           "-Wl,-rpath," + LIBDIR])
    return out


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_header_compiles_links_and_fails_loudly_without_device(exe):
    r = subprocess.run([exe, "nodevice"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_map_reference_test_in_cpp(exe):
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
