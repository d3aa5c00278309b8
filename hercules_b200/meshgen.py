"""Synthetic uniform meshes in the reference's own layout, for benchmarks and at-scale tests.

The reference meshes with octor from a material etree; at 256^3 elements that takes minutes of
single-threaded host time, so the bench builds its workload here instead.
For a uniform grid octor's output is fully determined:

* elements are the leaves in Morton (Z) order with x the fastest bit (octor.c:5373-5507),
* nodes are sorted by octor_zcompare on their tick coordinates (octor.c:3034, 6166) with the
  nodes on the far domain faces pulled in by one tick (octor.c:5466-5475), i.e. Morton order of
  the keys 2*ix (ix < nx) and 2*nx - 1 (ix = nx) on the half-step grid,
* elem_t.lnid lists the 8 corners x fastest, then y, then z (octor.c:6449-6470),

and the solver tables follow solver_init (psolve.c:3360-3473), mu_and_lambda (psolve.c:3236-3272),
compute_setab (psolve.c:5813-5876), compute_setflag / compute_setboundary (psolve.c:5629-5804,
built with -DBOUNDARY -DHALFSPACE) with the reference's float/double evaluation order.
tests/test_meshgen.py checks all of it bit-for-bit against a mesh the unmodified reference
produced (tests/golden/uniform_rayleigh_eff.npz).  This module only prepares HOST inputs; it
computes nothing the GPU path is responsible for.
"""
from __future__ import annotations

import math

import numpy as np

from .solver import HostMesh, RAYLEIGH, MASS

_XI = np.array([[-1, 1, -1, 1, -1, 1, -1, 1],
                [-1, -1, 1, 1, -1, -1, 1, 1],
                [-1, -1, -1, -1, 1, 1, 1, 1]], np.float64)      # psolve.c:5451-5453


def _part1by2(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def morton3(ix, iy, iz) -> np.ndarray:
    """x is the least significant interleaved bit (octor_zcompare, octor.c:3034-3110)."""
    return _part1by2(ix) | (_part1by2(iy) << np.uint64(1)) | (_part1by2(iz) << np.uint64(2))


def compute_K():
    """theK1 (= K1 + K3) and theK2, compute_K (psolve.c:5446-5573)."""
    def I1(xki, xkj, xli, xlj, xmi, xmj):
        return 4.5 * xki * xkj * (1 + xli * xlj / 3) * (1 + xmi * xmj / 3) / 8

    def I2(xki, xlj, xmi, xmj):
        return 4.5 * xki * xlj * (1 + xmi * xmj / 3) / 8
    x = _XI
    K1, K2, K3 = np.zeros((8, 8, 3, 3)), np.zeros((8, 8, 3, 3)), np.zeros((8, 8, 3, 3))
    for i in range(8):
        for j in range(8):
            for k in range(3):
                k0, k1, k2 = k % 3, (k + 1) % 3, (k + 2) % 3
                K3[i, j, k, k] = (I1(x[k0][i], x[k0][j], x[k1][i], x[k1][j], x[k2][i], x[k2][j])
                                  + I1(x[k1][i], x[k1][j], x[k2][i], x[k2][j], x[k0][i], x[k0][j])
                                  + I1(x[k2][i], x[k2][j], x[k0][i], x[k0][j], x[k1][i], x[k1][j]))
                for l in range(3):
                    if k == l:
                        K1[i, j, k, k] = I1(x[k][i], x[k][j], x[k1][i], x[k1][j], x[k2][i], x[k2][j])
                        K2[i, j, k, k] = I1(x[k][j], x[k][i], x[k1][j], x[k1][i], x[k2][j], x[k2][i])
                    else:
                        m = 3 - (k + l)
                        K1[i, j, k, l] = I2(x[k][j], x[l][i], x[m][j], x[m][i])
                        K2[i, j, k, l] = I2(x[k][i], x[l][j], x[m][i], x[m][j])
    return K1 + K3, K2


def compute_setab(damping: int, freq: float):
    """compute_setab (psolve.c:5813-5876)."""
    PI = 3.14159265358979323846264338327
    if damping == RAYLEIGH:
        w1, w2 = 2 * PI * freq * .2, 2 * PI * freq * 1
        lw1, lw2 = math.log(w1), math.log(w2)
        sw1, sw2 = w1 * w1, w2 * w2
        cw1, cw2 = w1 * w1 * w1, w2 * w2 * w2
        numer = w1 * w2 * (-2 * sw1 * lw2 + 2 * sw1 * lw1 - 2 * w1 * w2 * lw2
                           + 2 * w1 * w2 * lw1 + 3 * sw2 - 3 * sw1
                           - 2 * sw2 * lw2 + 2 * sw2 * lw1)
        denom = (cw1 - cw2 + 3 * sw2 * w1 - 3 * sw1 * w2)
        a = numer / denom
        numer = 3 * (2 * w1 * w2 * lw2 - 2 * w1 * w2 * lw1 + sw1 - sw2)
        return a, numer / denom
    if damping == MASS:
        w1, w2 = 2 * PI * freq * .1, 2 * PI * freq * 8
        return 1.3 * (2 * w2 * w1 * math.log(w2 / w1)) / (w2 - w1), 0.0
    return 0.0, 0.0


def uniform_halfspace(nx: int, ny: int, nz: int, h: float, dt: float, freq: float = 1.0,
                      damping: int = RAYLEIGH, layers=((0.0, 6000.0, 3464.0, 2700.0),),
                      thr_damping: float = 0.05, thr_vpvs: float = 3.0, exact: bool = False):
    """nx x ny x nz elements of edge h (x = north, y = east, z = depth).  layers = (ztop, Vp, Vs,
    rho) by depth of the element centre.  exact=True accumulates nTable in the reference's exact
    operation order (slow, for the bit-for-bit test); otherwise sums are grouped per element.
    Returns (HostMesh, info)."""
    f32 = np.float32
    # ---- elements in Morton order ------------------------------------------------------------
    ex, ey, ez = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
    order = np.argsort(morton3(ex, ey, ez), kind="stable")
    ex, ey, ez = ex[order], ey[order], ez[order]
    E = ex.size
    # ---- nodes in Morton order -------------------------------------------------------------------
    ix, iy, iz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    def key(i, n):
        i = i.ravel()
        return np.where(i == n, 2 * n - 1, 2 * i)
    norder = np.argsort(morton3(key(ix, nx), key(iy, ny), key(iz, nz)), kind="stable")
    N = norder.size
    rank = np.empty(N, np.int32)
    rank[norder] = np.arange(N, dtype=np.int32)
    rank = rank.reshape(nx + 1, ny + 1, nz + 1)
    lnid = np.empty((E, 8), np.int32)
    for j in range(8):
        lnid[:, j] = rank[ex + (j & 1), ey + ((j >> 1) & 1), ez + ((j >> 2) & 1)]
    del rank
    # ---- edata_t (floats) --------------------------------------------------------------------------
    zc = (ez + 0.5) * h
    Vp, Vs, rho = np.empty(E, f32), np.empty(E, f32), np.empty(E, f32)
    for (zt, vp, vs, r) in layers:
        sel = zc >= zt
        Vp[sel], Vs[sel], rho[sel] = vp, vs, r
    edge = np.full(E, h, f32)
    # ---- mu_and_lambda (psolve.c:3236-3272): float products, then double ------------------------
    mu = (rho * Vs * Vs).astype(np.float64)
    big = Vp > (Vs.astype(np.float64) * thr_vpvs)
    lam = np.where(big, (rho * Vs * Vs).astype(np.float64) * thr_vpvs * thr_vpvs - 2 * mu,
                   (rho * Vp * Vp).astype(np.float64) - 2 * mu)
    neg = lam < 0
    if neg.any():
        Vp = Vp.copy()
        f = np.where(Vs < 500, 2.45, np.where(Vs < 1200, 2.0, 1.87))
        Vp[neg] = (f[neg] * Vs[neg].astype(np.float64)).astype(f32)
        lam[neg] = (rho[neg] * Vp[neg] * Vp[neg]).astype(np.float64)
    dt2 = dt * dt
    abase, bbase = compute_setab(damping, freq)
    eT = np.empty((E, 4))
    eT[:, 0] = dt2 * edge * mu / 9
    eT[:, 1] = dt2 * edge * lam / 9
    zeta = (f32(10) / Vs).astype(np.float64)            # 10 / edata->Vs is a float division
    zeta = np.minimum(zeta, thr_damping)
    a, b = zeta * abase, zeta * bbase
    eT[:, 2] = b * dt * edge * mu / 9
    eT[:, 3] = b * dt * edge * lam / 9
    # ---- lumped mass and dashpots (psolve.c:3411-3473, 5752-5804) -------------------------------
    M = (rho * edge * edge * edge).astype(np.float64) / 8
    scale = (rho * (edge / f32(2)) * (edge / f32(2))).astype(np.float64)
    # absorbing faces: x near/far, y near/far, z far; the top (z near) is free under HALFSPACE
    touch = np.stack([np.stack([ex == 0, ex == nx - 1]), np.stack([ey == 0, ey == ny - 1]),
                      np.stack([np.zeros(E, bool), ez == nz - 1])])          # [axis][near/far][E]
    dash = np.zeros((E, 8, 3))
    bits = np.zeros((E, 8), np.int64)
    for j in range(8):
        for ax in range(3):
            far = (j >> ax) & 1
            bits[:, j] |= (touch[ax, far].astype(np.int64) << ax)
    nb = (bits & 1) + ((bits >> 1) & 1) + ((bits >> 2) & 1)
    vp_plus_2vs = (Vp + f32(2) * Vs).astype(np.float64)
    for c in range(3):
        on = ((bits >> c) & 1).astype(bool)
        vsel = np.where(on, Vp[:, None], Vs[:, None])                     # float
        two = (Vs[:, None] + vsel).astype(np.float64) * scale[:, None]    # (Vs + Vp|Vs) float, * double
        one = vsel.astype(np.float64) * scale[:, None]
        three = np.broadcast_to((vp_plus_2vs * scale)[:, None], (E, 8))
        dash[:, :, c] = np.where(nb == 3, three, np.where(nb == 2, two, np.where(nb == 1, one, 0.0)))
    boundary = nb.max(axis=1) > 0                                           # flag != 13
    nT = np.zeros((N, 7))
    flat = lnid.reshape(-1)
    if exact:
        # the reference's exact sequence per (element, corner, axis): -= dt a M; -= dt dashpot
        # (boundary elements only); += M  |  mass2: -= dt a M; -= dt dashpot; += 2 M
        np.add.at(nT[:, 0], flat, np.repeat(M, 8))
        daM = np.repeat(dt * a * M, 8)
        dd = dt * dash.reshape(-1, 3)
        bnd = np.repeat(boundary, 8)
        for ax in range(3):
            for col, mult in ((4 + ax, 1.0), (1 + ax, 2.0)):
                idx = np.stack([flat, flat, flat], 1)
                val = np.stack([-daM, np.where(bnd, -dd[:, ax], 0.0), np.repeat(M * mult, 8)], 1)
                keep = np.stack([np.ones_like(bnd), bnd, np.ones_like(bnd)], 1)
                np.add.at(nT[:, col], idx[keep], val[keep])
    else:
        nT[:, 0] = np.bincount(flat, np.repeat(M, 8), N)
        for ax in range(3):
            d = dt * dash[:, :, ax].reshape(-1)
            base = np.repeat(M - dt * a * M, 8)
            nT[:, 4 + ax] = np.bincount(flat, base - d, N)
            nT[:, 1 + ax] = np.bincount(flat, np.repeat(2 * M - dt * a * M, 8) - d, N)
    edata = np.zeros((E, 14), f32)
    edata[:, 0], edata[:, 1], edata[:, 2], edata[:, 3] = edge, Vp, Vs, rho
    K1, K2 = compute_K()
    mesh = HostMesh(lnid, eT, nT, np.zeros((0, 6), np.int32), edata, K1, K2)
    info = dict(E=E, N=N, abase=abase, bbase=bbase, node_order=norder, elem_xyz=(ex, ey, ez),
                dims=(nx, ny, nz), h=h)
    return mesh, info


def node_index(info: dict, ix: int, iy: int, iz: int) -> int:
    """Local node id of grid point (ix, iy, iz)."""
    nx, ny, nz = info["dims"]
    lin = (ix * (ny + 1) + iy) * (nz + 1) + iz
    return int(np.nonzero(info["node_order"] == lin)[0][0])


def element_index(info: dict, ex: int, ey: int, ez: int) -> int:
    """Local element id of the element whose lowest corner is grid point (ex, ey, ez) (local grid)."""
    X, Y, Z = info["elem_xyz"]
    hit = np.nonzero((X == ex) & (Y == ey) & (Z == ez))[0]
    if hit.size != 1:
        raise ValueError(f"element ({ex},{ey},{ez}) is not on this rank")
    return int(hit[0])
