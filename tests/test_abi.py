"""CPU test: libhercules_gpu.so loads and exports every symbol include/hercules_gpu.h declares
(no compute calls without a GPU), and refuses to run without a CUDA device instead of falling
back to anything."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def hb():
    import hercules_b200 as hb
    if not hb.SO.exists():
        hb.build()
    return hb


def test_header_symbols_exported(hb):
    hdr = (ROOT / "include" / "hercules_gpu.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hgpu_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    import ctypes
    L = ctypes.CDLL(str(hb.SO))
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    from hercules_b200 import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert hb.lib().hgpu_abi_version() == 1


def test_mesh_header_symbols_exported(hb):
    """libhercules_mesh.so (host-side octree primitives of the per-rank mesher) exports what
    include/hercules_mesh.h declares, and the Python side binds each of them."""
    import ctypes
    from hercules_b200 import octree_local
    hdr = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "hercules_mesh.h").read_text(), flags=re.S)
    declared = set(re.findall(r"\b(hmesh_[a-z_0-9]+)\s*\(", hdr))
    assert declared == {"hmesh_abi_version", "hmesh_free", "hmesh_chunk_leaves", "hmesh_chunk_nodes", "hmesh_lnid",
                        "hmesh_corner_sums", "hmesh_discovery"}
    L = ctypes.CDLL(str(ROOT / "hercules_b200" / "libhercules_mesh.so"))
    assert not [n for n in declared if not hasattr(L, n)]
    assert octree_local.mesh_lib().hmesh_abi_version() == 1


def test_no_cpu_fallback(hb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from conftest import load_golden
    g = load_golden("graded2_rayleigh_eff")
    with pytest.raises(hb.HerculesGpuError, match="no CUDA device|CUDA"):
        hb.Solver(hb.HostMesh.from_dump(g), dt=1e-3)


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under hercules_b200/ or include/ may name it."""
    for p in list((ROOT / "hercules_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ""} and p.name != "Makefile":
            txt = p.read_text(errors="ignore")
            assert "hercules_oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, p


@pytest.mark.parametrize("argv", [["--edge", "32"], ["--edge", "64", "--workload", "adaptive"],
                                  ["--edge", "32", "--damping", "bkt"], ["--edge", "64", "--workload", "adaptive", "--strong"],
                                  ["--edge", "128", "--workload", "basin", "--damping", "bkt"]])
def test_bench_builds_its_workload_and_needs_a_gpu(hb, argv, monkeypatch):
    """bench.py up to the creation of the solver (argument handling, mesh tables, source and station
    indices) runs on CPU; the solver itself refuses to exist without a CUDA device."""
    import runpy
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "5", "--warmup", "3"] + argv)
    fd = __import__("os").dup(1)
    try:
        with pytest.raises(hb.HerculesGpuError, match="CUDA"):
            runpy.run_path(str(ROOT / "bench.py"), run_name="__main__")
    finally:
        __import__("os").dup2(fd, 1)          # bench.py points fd 1 at stderr until its JSON line is ready
        __import__("os").close(fd)
