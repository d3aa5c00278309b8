"""CPU check of the structured-tile path of the step kernel (no GPU): tests/host/struct_check.cu is
compiled by nvcc for the HOST against the kernel header itself (its index and butterfly helpers are
__host__ __device__) and emulates the kernel's thread mapping and pass order on one 8x8x8 cell: padded
layout is a bijection, every warp access is bank-conflict-free, no two threads update one accumulator
without an ordering barrier between them, and the forces equal 512 single-element evaluations."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host compilation of the kernel header)")
def test_structured_tile_schedule_on_host(tmp_path):
    exe = tmp_path / "struct_check"
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-w", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I", str(ROOT / "include"), "-I", str(ROOT / "hercules_b200" / "csrc"),
                    "-o", str(exe), str(ROOT / "tests" / "host" / "struct_check.cu")], check=True)
    p = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0 and "struct_check: ok" in p.stdout, p.stdout[-3000:]


def test_structured_tiles_recognised(monkeypatch):
    """The planner marks exactly the aligned uniform cells that do not touch a far face of the domain, gives
    them the canonical slot order (re-checked inside hgpu_plan_build), and leaves everything else alone."""
    from hercules_b200 import meshgen, solver
    monkeypatch.setenv("HGPU_STRUCT", "1")
    mesh, info = meshgen.uniform_halfspace(32, 32, 24, h=25.0, dt=0.002)
    r = solver.plan_build(mesh.elem_lnid, info["N"])
    assert r["ntiles"] == 48 and r["struct_tiles"] == 3 * 3 * 2
    assert r["smem_bytes"] <= 115712
    import bench
    n = 64
    mesh, info = meshgen.graded_halfspace(n, n, bench.adaptive_bands(n), h=bench.H_M, dt=bench.DT, freq=bench.FREQ,
                                          layers=bench.adaptive_layers(n))
    r = solver.plan_build(mesh.elem_lnid, info["N"])
    assert r["tile_elems_total"] == info["E"] and 0 < r["struct_tiles"] < r["ntiles"]
    assert r["smem_bytes"] <= 115712
