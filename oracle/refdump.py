"""refdump.py -- reader for oracle/_ref/ref_dump output and force_process.<rank> files.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_DT = {0: np.int32, 1: np.int64, 2: np.float32, 3: np.float64, 4: np.int8, 5: np.uint32}


def read_dump(path) -> dict:
    """dump.<rank>.bin -> {name: ndarray}; layout documented in oracle/ref_dump.c."""
    out = {}
    buf = Path(path).read_bytes()
    off = 0
    while off < len(buf):
        name = buf[off:off + 48].split(b"\0", 1)[0].decode()
        dtype, ndim = struct.unpack_from("<ii", buf, off + 48)
        dims = struct.unpack_from("<4q", buf, off + 56)
        off += 88
        shape = tuple(int(d) for d in dims[:ndim])
        n = int(np.prod(shape)) if shape else 1
        dt = np.dtype(_DT[dtype])
        out[name] = np.frombuffer(buf, dtype=dt, count=n, offset=off).reshape(shape).copy()
        off += n * dt.itemsize
    return out


def read_force_process(path, steps: int | None = None):
    """force_process.<rank>: int32 n; int32 lnid[n]; double F[steps][n][3]
    (quake/forward/psolve.c:3651-3667, written by quakesource.c:compute_print_source)."""
    raw = Path(path).read_bytes()
    n = struct.unpack_from("<i", raw, 0)[0]
    lnid = np.frombuffer(raw, dtype=np.int32, count=n, offset=4).copy()
    body = np.frombuffer(raw, dtype=np.float64, offset=4 + 4 * n)
    if n == 0:
        return lnid, np.zeros((steps or 0, 0, 3))
    nsteps = body.size // (3 * n)
    f = body[: nsteps * 3 * n].reshape(nsteps, n, 3).copy()
    if steps is not None:
        f = f[:steps]
    return lnid, f


def params(d: dict) -> dict:
    p = d["params"]
    return dict(dt=p[0], dt2=p[1], freq=p[2], damping=int(p[3]), stiffness=int(p[4]),
                steps=int(p[5]), rank=int(p[6]), nranks=int(p[7]), print_accel=int(p[8]),
                abase=p[9], bbase=p[10], ticksize=p[11], thr_damping=p[12], thr_vpvs=p[13],
                etotal=int(p[14]), ntotal=int(p[15]))


def snapshots(d: dict, which: str = "tm1") -> dict:
    pre = which + "_step"
    return {int(k[len(pre):]): v for k, v in d.items() if k.startswith(pre)}
