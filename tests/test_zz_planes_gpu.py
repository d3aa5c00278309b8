"""The GPU cases of what was written last in round 2 (the file sorts last on purpose): planes interpolated on the
device (hgpu_planes_*, plane_kernel) -- run on a B200 --, the up-front station-ring check of hgpu_run
and conventional stiffness on its new default path -- not yet run on hardware.

Planes:

* through the C ABI: every recorded row equals the reference's interpolation arithmetic
  (Old_planes_print, io_planes.c:168-191) on the same field BIT FOR BIT, for arbitrary points;
* through the reference's own main (integration/_bin/psolve_gpu): planedisplacements.N written from device
  rows (the default on one rank, PSOLVE_GPU_DEVICE_PLANES=1 otherwise) are byte-identical to the files the
  reference's planes_print writes from fetched displacements, on one rank and on two.

The host side -- point tables, strip transport on 1/2/4 ranks, arithmetic order -- is pinned on CPU in
tests/test_planes_host.py.  Hardware run of this file: profiles/r02j_pytest_planes_device_1gpu.log.
"""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

from conftest import load_golden, params_of

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import refcase  # noqa: E402

pytestmark = pytest.mark.gpu
GPU_BIN = ROOT / "integration" / "_bin" / "psolve_gpu"


@pytest.fixture(scope="module")
def hb():
    import hercules_b200 as hb
    if not hb.SO.exists():
        hb.build()
    assert hb.lib().hgpu_device_count() > 0, "no CUDA device: GPU tests cannot run"
    return hb


def _plane_rows_numpy(loc, nodes, t1):
    """Old_planes_print's arithmetic (io_planes.c:168-191) in numpy: one rounded operation at a time."""
    xi = np.array([[-1, 1, -1, 1, -1, 1, -1, 1], [-1, -1, 1, 1, -1, -1, 1, 1], [-1, -1, -1, -1, 1, 1, 1, 1]], float)
    d = np.zeros((nodes.shape[0], 3))
    for i in range(8):
        phi = (1 + xi[0, i] * loc[:, 0]) * (1 + xi[1, i] * loc[:, 1]) * (1 + xi[2, i] * loc[:, 2]) / 8
        d = d + phi[:, None] * t1[nodes[:, i]]
    return d


def test_device_planes_bit_exact(hb):
    g = load_golden("graded2_rayleigh_eff")
    P = params_of(g)
    s = hb.Solver(hb.HostMesh.from_dump(g), dt=P["dt"], dt2=P["dt2"], damping=P["damping"], stiffness=P["stiffness"],
                  freq=P["freq"], loaded_lnid=g["loaded_lnid"])
    rng = np.random.default_rng(11)
    npts = 5000
    nodes = g["elem_lnid"][rng.integers(0, g["elem_lnid"].shape[0], npts)].astype(np.int32)
    loc = rng.uniform(-1, 1, (npts, 3))
    loc[:8] = np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], float)   # the corners
    s.planes_attach(nodes, loc)
    bufs = [np.empty((npts, 3)), np.empty((npts, 3))]
    moved = False
    for k in range(P["steps"]):
        s.step_begin(k)
        if k % 3 == 0:
            # two records in flight (tm1 does not change between them): both buffers must hold the rows
            a = s.planes_record(bufs[0])
            b = s.planes_record(bufs[1])
            want = _plane_rows_numpy(loc, nodes, s.fetch_all(hb.TM1))
            s.planes_wait()
            assert np.array_equal(a, want) and np.array_equal(b, want), k
            assert np.array_equal(want[:8], s.fetch_all(hb.TM1)[nodes[np.arange(8), np.arange(8)]])    # phi = 1 at a corner
            moved |= bool(np.abs(want).max() > 0)
        s.compute_force_source(g["forces"][k]); s.compute_force_stiffness(); s.compute_force_damping()
        s.send_force_and_adjust(); s.compute_displacement(); s.send_displacement_and_adjust()
    assert moved
    # re-attach with no points: record / wait are no-ops
    s.planes_attach(np.zeros((0, 8), np.int32), np.zeros((0, 3)))
    s.planes_record(np.empty((0, 3))); s.planes_wait()
    with pytest.raises(hb.HerculesGpuError, match="out of range"):
        s.planes_attach(np.full((1, 8), 10 ** 9, np.int32), np.zeros((1, 3)))
    s.close()


TWO_LAYER = dict(cvm_level=3, cvm_n=(8, 8, 4), vs_min=800, freq_hz=2.5,
                 layers=[(0, 3000, 1732, 2000), (125, 6000, 3464, 2700)])
THREE_LAYER = dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5,
                   layers=[(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)])
SRC = dict(src_xyz=(437.5, 562.5, 140.0), src_strike_dip_rake=(30.0, 70.0, 20.0), stations=[(500.0, 500.0, 0.0)])
PLANES = [(100.0, 150.0, 0.0, 50.0, 15, 50.0, 12, 0.0, 0.0), (200.0, 300.0, 20.0, 40.0, 10, 30.0, 8, 30.0, 60.0)]


def _run_gpu(case, env_extra, nranks, keep):
    with tempfile.TemporaryDirectory() as td:
        d = refcase.write_case(case, td)
        p = subprocess.run([str(GPU_BIN), "parameters.in"], cwd=d, env=dict(os.environ, HMPI_NP=str(nranks), **env_extra),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert p.returncode == 0, p.stdout[-4000:]
        return [(d / k).read_bytes() for k in keep], (d / "out" / "monitor.txt").read_text()


@pytest.mark.parametrize("nranks", [1, 2])
def test_device_planes_write_the_same_files(hb, nranks):
    if not (refcase.have_ref("mkcvm") and GPU_BIN.exists()):
        pytest.skip("integration/_bin/psolve_gpu not built")
    if hb.lib().hgpu_device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    model = TWO_LAYER if nranks == 1 else THREE_LAYER
    c = refcase.Case(**model, **SRC, damping="rayleigh", stiffness="effective", end_t=0.06, planes=PLANES, plane_rate=5)
    keep = [f"out/planes/planedisplacements.{i}" for i in range(len(PLANES))]
    # one rank: device planes are the default; several ranks: opt-in
    dev, mon = _run_gpu(c, {} if nranks == 1 else {"PSOLVE_GPU_DEVICE_PLANES": "1"}, nranks, keep)
    assert "planes: interpolated on the device" in mon
    host, mon = _run_gpu(c, {"PSOLVE_GPU_DEVICE_PLANES": "0"}, nranks, keep)
    assert "planes: reference planes_print" in mon
    assert dev == host
    assert all(np.abs(np.frombuffer(f, np.float64)).max() > 0 for f in dev)


def test_run_checks_the_station_ring_before_the_first_step(hb):
    """hgpu_run with stations recorded at a cadence: a run whose rows do not fit the device ring is refused up
    front (it used to fail in the middle of a step, after the swap), and leaves the solver untouched."""
    g = load_golden("graded2_rayleigh_eff")
    P = params_of(g)
    s = hb.Solver(hb.HostMesh.from_dump(g), dt=P["dt"], dt2=P["dt2"], damping=P["damping"], stiffness=P["stiffness"],
                  freq=P["freq"], loaded_lnid=g["loaded_lnid"])
    s.stations_attach(g["station_nodes"][:, 1:].astype(np.int32), g["station_local"], rate=2, capacity=3)
    with pytest.raises(hb.HerculesGpuError, match="station rows"):
        s.run(0, 10, g["forces"][:10])                           # rows at steps 0, 2, 4, 6, 8: five > three
    assert s.stations_pending() == 0 and not s.fetch_all(hb.TM1).any()
    s.run(0, 6, g["forces"][:6])                                 # rows at 0, 2, 4: fits
    steps, rows = s.stations_drain()
    assert list(steps) == [0, 2, 4] and rows.shape[0] == 3
    s.close()


# ---- conventional stiffness on its default path -------------------------------------------------------------
# A solver created with HGPU_STIFFNESS_CONVENTIONAL applies the operator in the factored form of the effective
# method since the end of round 2 (one host-side boolean; the literal dense products
# keep test_dense_conventional_kernel_matches_reference below).  These are the checks the
# conventional golden always ran -- per call against the oracle's compute_addforce_conventional, whole run against
# the reference's conventional-stiffness snapshots -- moved here because that switch has not run on hardware yet.

@pytest.mark.parametrize("tile_nodes", [0, 64, 200])
def test_conventional_force_calls_match_oracle(hb, oracle, tile_nodes):
    from test_gpu_parity import force_calls_case
    force_calls_case(hb, oracle, "graded2_rayleigh_conv", tile_nodes)


@pytest.mark.parametrize("flags", [0, 1])
def test_conventional_whole_run_matches_reference(hb, flags):
    from test_gpu_parity import whole_run_case
    whole_run_case(hb, "graded2_rayleigh_conv", flags)


@pytest.mark.parametrize("flags", [0, 1])
def test_dense_conventional_kernel_matches_reference(hb, flags):
    """HGPU_FLAG_DENSE_K: compute_addforce_conventional as the literal dense 24 x 24 K1 / K2 products (the DENSE
    variant of the step kernel) -- whole run against the reference's conventional-stiffness snapshots, in the
    loop of test_whole_run_matches_reference.  Without the flag a conventional solver applies the same operator
    in factored form (test_conventional_whole_run_matches_reference above)."""
    from test_gpu_parity import make_solver, snapshots, rel_l2, REL_TOL_RUN
    g = load_golden("graded2_rayleigh_conv")
    s, P = make_solver(hb, g, flags=flags | hb.FLAG_DENSE_K)
    assert P["stiffness"] == hb.CONVENTIONAL
    snaps = snapshots(g)
    F = g["forces"]
    for k in range(P["steps"]):
        s.step_begin(k)
        if k in snaps and np.abs(snaps[k]).max() > 0:
            assert rel_l2(s.fetch_all(hb.TM1), snaps[k]) < REL_TOL_RUN, k
        s.compute_force_source(F[k]); s.compute_force_stiffness(); s.compute_force_damping()
        s.send_force_and_adjust(); s.compute_displacement(); s.send_displacement_and_adjust()
    assert np.abs(snaps[max(snaps)]).max() > 0
    s.close()
