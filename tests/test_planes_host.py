"""CPU check of the device-plane boundary: integration/io_planes_gpu.c hands the reference's
plane point tables to hgpu_planes_attach and moves / prints device-interpolated rows the way
Old_planes_print does after its own interpolation (io_planes.c:193-247).

integration/_bin/psolve_planes_hostcheck is the UNMODIFIED reference (oracle/_ref/O2/psolve.o, CPU time loop)
whose planes_print goes through exactly those two functions, with the arithmetic of plane_kernel
(hercules_b200/csrc/hgpu_kernels.cuh) evaluated on the host.  Its planedisplacements.N must be byte-identical
to the reference's own on 1, 2 and 4 ranks (strips produced on every rank, received by rank 0).
The device kernel itself is checked in tests/test_zz_planes_gpu.py.
"""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import refcase  # noqa: E402

CHECK_BIN = ROOT / "integration" / "_bin" / "psolve_planes_hostcheck"
THREE_LAYER = dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5,
                   layers=[(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)])
SRC = dict(src_xyz=(437.5, 562.5, 140.0), src_strike_dip_rake=(30.0, 70.0, 20.0),
           stations=[(500.0, 500.0, 0.0)])
# a horizontal plane over most of the surface and a dipping one through the source region
PLANES = [(50.0, 50.0, 0.0, 60.0, 15, 60.0, 15, 0.0, 0.0), (200.0, 300.0, 20.0, 40.0, 12, 30.0, 10, 30.0, 60.0)]


@pytest.mark.parametrize("nranks", [1, 2, 4])
def test_plane_tables_and_strip_transport_reproduce_the_reference_files(nranks):
    if not (refcase.have_ref("psolve_ref_O2") and refcase.have_ref("mkcvm") and CHECK_BIN.exists()):
        pytest.skip("reference binaries / integration/_bin/psolve_planes_hostcheck not built (need /root/reference)")
    c = refcase.Case(**THREE_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.1, planes=PLANES,
                     plane_rate=4)
    keep = [f"out/planes/planedisplacements.{i}" for i in range(len(PLANES))]
    files = {}
    for which in ("ref", "check"):
        with tempfile.TemporaryDirectory() as td:
            d = refcase.write_case(c, td)
            if which == "ref":
                refcase.run("psolve_ref_O2", d, nranks=nranks, timeout=300)
            else:
                p = subprocess.run([str(CHECK_BIN), "parameters.in"], cwd=d, env=dict(os.environ, HMPI_NP=str(nranks)),
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
                assert p.returncode == 0, p.stdout[-4000:]
            files[which] = [(d / k).read_bytes() for k in keep]
    for a, b, (_, _, _, _, ns, _, nd, _, _) in zip(files["ref"], files["check"], PLANES):
        v = np.frombuffer(a, np.float64)
        assert v.size == 3 * ns * nd * (c.steps // 4 + (1 if c.steps % 4 else 0)) and np.abs(v).max() > 0
        assert a == b
