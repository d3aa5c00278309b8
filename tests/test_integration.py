"""Drop-in test at the reference's own boundary: integration/_bin/psolve_gpu is the UNMODIFIED
reference psolve.c (compiled where it lies) whose time loop is executed by libhercules_gpu.so
(integration/psolve_gpu.c).  The same case is run by the reference's CPU binary
(oracle/_ref/psolve_ref_O2) and by psolve_gpu; the station files -- written in both runs by the
reference's own interpolate_station_displacements (psolve.c:6680-6795) -- must agree to the
precision they are printed with (`% 8e`: 7 significant digits), the way the reference's shipped
goldens pin examples/simple.

Both binaries are built in the development container (they need /root/reference) and travel to
the GPU box with the repository snapshot; the tests skip when they are absent.
"""
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import refcase  # noqa: E402

pytestmark = pytest.mark.gpu

GPU_BIN = ROOT / "integration" / "_bin" / "psolve_gpu"

TWO_LAYER = dict(cvm_level=3, cvm_n=(8, 8, 4), vs_min=800, freq_hz=2.5,
                 layers=[(0, 3000, 1732, 2000), (125, 6000, 3464, 2700)])
THREE_LAYER = dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5,
                   layers=[(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)])
SRC = dict(src_xyz=(437.5, 562.5, 140.0), src_strike_dip_rake=(30.0, 70.0, 20.0),
           stations=[(500.0, 500.0, 0.0), (700.0, 300.0, 50.0), (120.0, 880.0, 300.0)])


def read_station(path: Path) -> np.ndarray:
    rows = [list(map(float, ln.split())) for ln in path.read_text().splitlines()
            if ln.strip() and not ln.lstrip().startswith("#")]
    return np.array(rows)


def run_both(case: refcase.Case, nranks: int):
    import subprocess, os
    if not (refcase.have_ref("psolve_ref_O2") and refcase.have_ref("mkcvm") and GPU_BIN.exists()):
        pytest.skip("reference binaries / integration/_bin/psolve_gpu not built (need /root/reference at build time)")
    out = {}
    for which in ("ref", "gpu"):
        with tempfile.TemporaryDirectory() as td:
            d = refcase.write_case(case, td)
            if which == "ref":
                log = refcase.run("psolve_ref_O2", d, nranks=nranks, timeout=300)
            else:
                env = dict(os.environ, HMPI_NP=str(nranks))
                p = subprocess.run([str(GPU_BIN), "parameters.in"], cwd=d, env=env, stdout=subprocess.PIPE,
                                   stderr=subprocess.STDOUT, text=True, timeout=300)
                assert p.returncode == 0, p.stdout[-4000:]
                log = p.stdout
                mon = (d / "out" / "monitor.txt").read_text()
                assert "gpu_solver_run() done" in mon, mon[-2000:]
            out[which] = ([read_station(d / "out" / "stations" / f"station.{i}") for i in range(len(case.stations))],
                          refcase.parse_timing(log))
    return out


def check(out, steps):
    (ref, tref), (gpu, tgpu) = out["ref"], out["gpu"]
    assert tref.get("elements") == tgpu.get("elements") and tref.get("steps") == tgpu.get("steps") == steps
    moved = False
    for a, b in zip(ref, gpu):
        assert a.shape == b.shape and a.shape[0] == steps
        assert np.array_equal(a[:, 0], b[:, 0])                    # time column
        scale = np.abs(a[:, 1:4]).max()
        moved |= scale > 0
        # 7 printed digits: one unit in the last place of either print, plus values that round to ~0
        # (a station the wave has not reached holds exact zeros in one run and 1e-45-size noise of
        # flushed products in the other: absolute floor far below anything physical)
        assert np.all(np.abs(a[:, 1:] - b[:, 1:]) <= 2.5e-6 * np.abs(a[:, 1:]) + 1e-9 * scale + 1e-30)
    assert moved


@pytest.mark.parametrize("damping,stiffness", [("rayleigh", "effective"), ("bkt", "effective"), ("none", "effective"),
                                               ("rayleigh", "conventional")])
def test_reference_main_with_gpu_time_loop(damping, stiffness):
    """Two-layer model: octor produces two refinement levels with hanging nodes on the interface."""
    c = refcase.Case(**TWO_LAYER, **SRC, damping=damping, stiffness=stiffness, end_t=0.06)
    check(run_both(c, 1), c.steps)


def test_reference_main_with_gpu_time_loop_accelerations():
    """print_station_accelerations = yes: the station writer also reads tm2 and tm3 (psolve.c:6738-6775)."""
    c = refcase.Case(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.03, print_accel="yes")
    check(run_both(c, 1), c.steps)


@pytest.mark.parametrize("nranks", [2, 4])
def test_reference_main_with_gpu_time_loop_multirank(nranks):
    """octor's partition across mini-MPI ranks, one GPU per rank, halo exchange over peer memory."""
    import hercules_b200 as hb
    if hb.lib().hgpu_device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    c = refcase.Case(**THREE_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.1)
    check(run_both(c, nranks), c.steps)


def test_reference_main_with_gpu_time_loop_1p5M_elements():
    """The same three-layer model meshed for 20 Hz: 1.5 M elements on three octree levels, 62 k
    hanging nodes.  CPU reference on up to 8 mini-MPI ranks, GPU run on one rank: the station
    series are partition-independent and must agree to their printed precision."""
    import os
    # stations next to the hypocentre: 60 steps of 0.125 ms do not carry the wave far
    src = dict(SRC, stations=[(437.5, 562.5, 140.0), (441.0, 560.0, 137.0), (430.0, 570.0, 150.0)])
    c = refcase.Case(**{**THREE_LAYER, "freq_hz": 20.0}, **src, damping="rayleigh", stiffness="effective",
                     dt=0.000125, end_t=0.000125 * 60.5, src_risetime=0.004)
    if not (refcase.have_ref("psolve_ref_O2") and refcase.have_ref("mkcvm") and GPU_BIN.exists()):
        pytest.skip("reference binaries / integration/_bin/psolve_gpu not built")
    import subprocess
    nref = 1
    while nref * 2 <= min(8, os.cpu_count() or 1):
        nref *= 2
    out = {}
    for which in ("ref", "gpu"):
        with tempfile.TemporaryDirectory() as td:
            d = refcase.write_case(c, td)
            if which == "ref":
                log = refcase.run("psolve_ref_O2", d, nranks=nref, timeout=1500)
            else:
                p = subprocess.run([str(GPU_BIN), "parameters.in"], cwd=d, env=dict(os.environ, HMPI_NP="1"),
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
                assert p.returncode == 0, p.stdout[-4000:]
                log = p.stdout
            out[which] = ([read_station(d / "out" / "stations" / f"station.{i}") for i in range(len(c.stations))],
                          refcase.parse_timing(log))
    assert out["gpu"][1]["elements"] == 1507328
    check(out, c.steps)


def test_device_stations_write_the_same_files():
    """psolve_gpu with the station interpolation on the device (default) and with the reference's
    interpolate_station_displacements on fetched displacements (PSOLVE_GPU_HOST_STATIONS=1) write
    byte-identical station files, velocities and accelerations included."""
    import os, subprocess
    if not (refcase.have_ref("mkcvm") and GPU_BIN.exists()):
        pytest.skip("integration/_bin/psolve_gpu not built")
    c = refcase.Case(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.05, print_accel="yes",
                     station_rate=3)
    files = {}
    for host in ("0", "1"):
        with tempfile.TemporaryDirectory() as td:
            d = refcase.write_case(c, td)
            p = subprocess.run([str(GPU_BIN), "parameters.in"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                               text=True, timeout=300, env=dict(os.environ, HMPI_NP="1", PSOLVE_GPU_HOST_STATIONS=host))
            assert p.returncode == 0, p.stdout[-4000:]
            files[host] = [(d / "out" / "stations" / f"station.{i}").read_bytes() for i in range(len(c.stations))]
    assert files["0"] == files["1"]
    assert all(len(f.splitlines()) > 10 for f in files["0"])


PLANES = [(100.0, 150.0, 0.0, 50.0, 15, 50.0, 12, 0.0, 0.0), (200.0, 300.0, 20.0, 40.0, 10, 30.0, 8, 30.0, 60.0)]


def _run_gpu(case, env_extra=None, nranks=1, keep=()):
    import os, subprocess
    with tempfile.TemporaryDirectory() as td:
        d = refcase.write_case(case, td)
        env = dict(os.environ, HMPI_NP=str(nranks), **(env_extra or {}))
        p = subprocess.run([str(GPU_BIN), "parameters.in"], cwd=d, env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0, p.stdout[-4000:]
        return {k: (d / k).read_bytes() for k in keep}, p.stdout


def test_planes_sparse_fetch_writes_the_same_files():
    """Output planes (io_planes.c:151-250): psolve_gpu fetches only the rows of tm1 the reference's own
    planes_print interpolates from (node list from the reference's plane tables, integration/io_planes_gpu.c)
    instead of the whole field.  planedisplacements.N must be byte-identical to the whole-field variant
    (PSOLVE_GPU_PLANES_FULL=1) and agree with the CPU reference's files to 1e-10 relative L2 (north_star)."""
    if not (refcase.have_ref("psolve_ref_O2") and refcase.have_ref("mkcvm") and GPU_BIN.exists()):
        pytest.skip("reference binaries / integration/_bin/psolve_gpu not built")
    c = refcase.Case(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.06, planes=PLANES, plane_rate=5)
    keep = [f"out/planes/planedisplacements.{i}" for i in range(len(PLANES))]
    sparse, _ = _run_gpu(c, {"PSOLVE_GPU_DEVICE_PLANES": "0"}, keep=keep)
    full, _ = _run_gpu(c, {"PSOLVE_GPU_DEVICE_PLANES": "0", "PSOLVE_GPU_PLANES_FULL": "1"}, keep=keep)
    assert sparse == full
    with tempfile.TemporaryDirectory() as td:
        d = refcase.write_case(c, td)
        refcase.run("psolve_ref_O2", d, nranks=1, timeout=300)
        for k in keep:
            ref = np.frombuffer((d / k).read_bytes(), np.float64)
            got = np.frombuffer(sparse[k], np.float64)
            assert ref.shape == got.shape and ref.size == 3 * (c.steps // 5 + (1 if c.steps % 5 else 0)) * (15 * 12 if k.endswith("0") else 10 * 8)
            assert np.abs(ref).max() > 0
            assert np.linalg.norm(got - ref) <= 1e-10 * np.linalg.norm(ref)


def test_async_checkpoint_same_bytes_and_restart():
    """Checkpoints (io_checkpoint.c:29-117): psolve_gpu snapshots tm1 / tm2 on the device, lets the time loop
    run on and writes the reference's file layout from a writer thread (hgpu_fetch_all_async).  The files
    must be byte-identical to those of the reference's own checkpoint_write on synchronously fetched fields
    (PSOLVE_GPU_ASYNC_CKPT=0), and a run restarted from one (checkpoint_read, use_checkpoint = 1) must
    continue exactly as the uninterrupted run."""
    import os, shutil, subprocess
    if not (refcase.have_ref("mkcvm") and GPU_BIN.exists()):
        pytest.skip("integration/_bin/psolve_gpu not built")
    c = refcase.Case(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.06, checkpoint_rate=20)
    keep = ["out/checkpoints/checkpoint.out0", "out/checkpoints/checkpoint.out1"] + [f"out/stations/station.{i}" for i in range(3)]
    a, _ = _run_gpu(c, keep=keep)
    b, _ = _run_gpu(c, {"PSOLVE_GPU_ASYNC_CKPT": "0"}, keep=keep)
    for k in keep:
        assert a[k] == b[k], k
    hdr = np.frombuffer(a[keep[1]][:12], np.int32)
    assert hdr[0] == 1 and hdr[1] == 40 and len(a[keep[1]]) == 12 + 2 * 24 * hdr[2]
    assert np.abs(np.frombuffer(a[keep[1]][12:], np.float64)).max() > 0
    # restart from the step-40 checkpoint
    c2 = refcase.Case(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.06, use_checkpoint=1)
    with tempfile.TemporaryDirectory() as td:
        d = refcase.write_case(c2, td)
        (d / "out" / "checkpoints" / "checkpoint.in").write_bytes(a[keep[1]])
        p = subprocess.run([str(GPU_BIN), "parameters.in"], cwd=d, env=dict(os.environ, HMPI_NP="1"), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0, p.stdout[-4000:]
        for i in range(3):
            rows_full = [ln for ln in a[f"out/stations/station.{i}"].decode().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
            rows_re = [ln for ln in (d / "out" / "stations" / f"station.{i}").read_text().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
            assert len(rows_full) == c.steps and rows_re == rows_full[40:], i
