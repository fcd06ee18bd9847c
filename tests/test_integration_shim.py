"""The reference-side C++ shim of INTEGRATION.md (integration/TensorB200.hpp: Xped::Tensor<double, Rank, CoRank, Symmetry> forwarding to the C ABI) is compiled
with g++ against a minimal mock of the reference's Qbasis / symmetry tags and linked with libxped_b200.so -- on the CPU that proves the sketch binds (every
symbol resolves); with a GPU the program also runs the permute round trip, product, contract, tSVD and tQR checks through the shim."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim():
    import xped_b200
    xped_b200.load()                      # the library must exist (built by __graft_entry__.build() / make)
    out = os.path.join(ROOT, "build", "test_shim")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "integration"),
           "-I" + os.path.join(ROOT, "integration", "mock"), os.path.join(ROOT, "integration", "test_shim.cpp"), "-o", out,
           "-L" + os.path.join(ROOT, "xped_b200", "lib"), "-lxped_b200", "-Wl,-rpath," + os.path.join(ROOT, "xped_b200", "lib")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_shim_compiles_and_links():
    exe = build_shim()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode in (0, 77), r.stdout + r.stderr      # 77: no CUDA device (nothing to run)


@pytest.mark.gpu
def test_shim_runs_on_the_device():
    exe = build_shim()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
