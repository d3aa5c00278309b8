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
