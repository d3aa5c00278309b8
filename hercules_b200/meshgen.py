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

from .solver import HostMesh, MsgList, RAYLEIGH, MASS, BKT

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


_SPREAD8 = None
_PACK12 = None


def _luts():
    global _SPREAD8, _PACK12
    if _SPREAD8 is None:
        b = np.arange(256, dtype=np.uint64)
        _SPREAD8 = _part1by2(b)                                  # 8 bits -> every third bit of 24
        c = np.arange(4096, dtype=np.uint64)
        x, y, z = (_compact1by2(c >> np.uint64(k)).astype(np.uint32) for k in range(3))
        _PACK12 = x | (y << np.uint32(8)) | (z << np.uint32(16))   # 12 interleaved bits -> 4 bits of x, y, z
    return _SPREAD8, _PACK12


def morton3_fast(ix, iy, iz) -> np.ndarray:
    """morton3 for coordinates below 2^16, two table look-ups per coordinate."""
    S, _ = _luts()
    out = None
    for sh, v in ((0, ix), (1, iy), (2, iz)):
        v = np.asarray(v).astype(np.uint32)
        c = S[v & np.uint32(255)] | (S[v >> np.uint32(8)] << np.uint64(24))
        c = c << np.uint64(sh) if sh else c
        out = c if out is None else (out | c)
    return out


def demorton3_fast(codes: np.ndarray, chunks: int = 3):
    """Inverse of morton3 for codes below 2^(12 chunks) (4 chunks coordinate bits): x, y, z as int32."""
    _, P = _luts()
    x = y = z = None
    for k in range(chunks):
        t = P[(codes >> np.uint64(12 * k)) & np.uint64(4095)]
        px, py, pz = (t & np.uint32(15)), ((t >> np.uint32(8)) & np.uint32(15)), (t >> np.uint32(16))
        if k:
            px, py, pz = px << np.uint32(4 * k), py << np.uint32(4 * k), pz << np.uint32(4 * k)
        x, y, z = (px, py, pz) if x is None else (x | px, y | py, z | pz)
    return x.astype(np.int32), y.astype(np.int32), z.astype(np.int32)


def _compact1by2(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0x1249249249249249)
    v = (v | (v >> np.uint64(2))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v >> np.uint64(4))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v >> np.uint64(8))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v >> np.uint64(16))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v >> np.uint64(32))) & np.uint64(0x1FFFFF)
    return v.astype(np.int64)


def block_low(task: int, group: int, n: int) -> int:
    """octor.c:685-690."""
    return task * n // group


def block_owner(idx, group: int, n: int):
    """octor.c:738-742."""
    return ((np.asarray(idx, np.int64) + 1) * group - 1) // n


class _MortonIndex:
    """Global preorder (Morton) rank of the leaves of a uniform nx x ny x nz grid, i.e. octor's
    geid (octor.c:5362-5374, 5505), without sorting the whole grid: the grid is cut into cubes of
    n^3 elements, n the largest power of two dividing all three dimensions; cubes are ordered by
    the Morton code of their coordinates and elements inside a cube by theirs."""

    def __init__(self, nx: int, ny: int, nz: int):
        g = math.gcd(math.gcd(nx, ny), nz)
        n = g & -g
        self.n, self.n3 = n, n ** 3
        self.dims = (nx, ny, nz)
        B = (nx // n, ny // n, nz // n)
        bx, by, bz = np.meshgrid(np.arange(B[0]), np.arange(B[1]), np.arange(B[2]), indexing="ij")
        bx, by, bz = bx.ravel(), by.ravel(), bz.ravel()
        o = np.argsort(morton3(bx, by, bz), kind="stable")
        self.bx, self.by, self.bz = bx[o].astype(np.int64), by[o].astype(np.int64), bz[o].astype(np.int64)
        self.brank = np.empty(B, np.int64)
        self.brank[self.bx, self.by, self.bz] = np.arange(o.size)
        self.total = nx * ny * nz

    def index(self, ex, ey, ez) -> np.ndarray:
        n = self.n
        ex, ey, ez = np.asarray(ex, np.int64), np.asarray(ey, np.int64), np.asarray(ez, np.int64)
        return self.brank[ex // n, ey // n, ez // n] * self.n3 + \
            morton3(ex % n, ey % n, ez % n).astype(np.int64)

    def coords(self, idx):
        idx = np.asarray(idx, np.int64)
        b, w = idx // self.n3, (idx % self.n3).astype(np.uint64)
        n = self.n
        return (self.bx[b] * n + _compact1by2(w), self.by[b] * n + _compact1by2(w >> np.uint64(1)),
                self.bz[b] * n + _compact1by2(w >> np.uint64(2)))


# constract_Quality_Factor_Table (psolve.c:5575-5616): 26 rows are declared but only the first 18
# are copied into Global.theQTABLE; the rest stay zero (SURVEY.md 8a notes): kept as is.
_QTABLE = np.zeros((26, 6))
_QTABLE[:18] = np.array([
    [5., 0.211111102, 0.236842104, 0.032142857, 0.271428571, 0.14],
    [6.25, 0.188888889, 0.184210526, 0.039893617, 0.336879433, 0.10152],
    [8.33, 0.157777778, 0.139473684, 0.045, 0.38, 0.07],
    [10., 0.137777765, 0.12105263, 0.032942899, 0.27818448, 0.0683],
    [15., 0.097777765, 0.08105263, 0.032942899, 0.27818448, 0.045],
    [20., 0.078139527, 0.060526314, 0.031409788, 0.277574872, 0.034225],
    [25., 0.064285708, 0.049999999, 0.031578947, 0.285714286, 0.0266],
    [30., 0.053658537, 0.044736842, 0.026640676, 0.24691358, 0.023085],
    [35., 0.046341463, 0.038157895, 0.02709848, 0.251156642, 0.019669],
    [40., 0.040487805, 0.034210526, 0.025949367, 0.240506329, 0.01738],
    [45., 0.036585366, 0.028947368, 0.031393568, 0.290964778, 0.014366],
    [50., 0.032926829, 0.026315789, 0.032488114, 0.30110935, 0.01262],
    [60., 0.0279, 0.0223, 0.0275, 0.2545, 0.0114],
    [70., 0.024, 0.019, 0.032488114, 0.30110935, 0.0083],
    [80., 0.0207, 0.0174, 0.0251, 0.2326, 0.0088],
    [90., 0.0187, 0.0154, 0.0244, 0.2256, 0.0079],
    [100., 0.017, 0.014, 0.028021016, 0.288966725, 0.006281],
    [120., 0.0142, 0.0115, 0.0280, 0.2700, 0.0052]])


def _search_quality_table(Q: np.ndarray) -> np.ndarray:
    """Search_Quality_Table (quake_util.c:128-163), vectorised: the first row whose distance to Q
    stops decreasing ends the scan and the row before it is returned; -1 for Q > 500; -2 = none."""
    Q = np.asarray(Q, np.float64)
    out = np.full(Q.shape, -2, np.int64)
    done = Q > 500
    out[done] = -1
    mn = np.full(Q.shape, 1000.0)
    for i in range(_QTABLE.shape[0]):
        diff = np.abs(Q - _QTABLE[i, 0])
        stop = ~done & ~(diff < mn)
        out[stop] = i - 1
        done |= stop
        mn = np.where(~done & (diff < mn), diff, mn)
    return out


def bkt_coefficients(Vp, Vs, use_inf_qk: bool = False) -> np.ndarray:
    """The ten BKT floats of edata_t (a0, a1, b, g0, g1 for shear, then for kappa; psolve.h:95-97)
    as mesh_correct_properties derives them from the element's float Vp, Vs (psolve.c:7239-7310,
    simulation_velocity_profile_freq_hz = 0).  Returns float32 [E][10] in edata order
    a0_shear a1_shear b_shear g0_shear g1_shear a0_kappa a1_kappa b_kappa g0_kappa g1_kappa."""
    Vp = np.asarray(Vp, np.float32).astype(np.float64)
    Vs = np.asarray(Vs, np.float32).astype(np.float64)
    vs_vp = (np.asarray(Vs, np.float32) / np.asarray(Vp, np.float32)).astype(np.float64)   # float division
    vs = Vs * 0.001
    L = 4. / 3. * vs_vp * vs_vp
    Qs = 10.5 + vs * (-16. + vs * (153. + vs * (-103. + vs * (34.7 + vs * (-5.29 + vs * 0.31)))))
    Qp = 2. * Qs
    Qk = np.full(Qs.shape, 1000.0) if use_inf_qk else (1. - L) / (1. / Qp - L / Qs)
    out = np.zeros((Vs.size, 10), np.float32)
    for col, Q in ((0, Qs), (5, Qk)):
        idx = _search_quality_table(Q)
        if (idx == -2).any() or (idx >= _QTABLE.shape[0]).any():
            raise ValueError("Problem with the Quality Factor Table (psolve.c:7266)")
        ok = idx >= 0
        row = _QTABLE[np.where(ok, idx, 0)]
        # table columns 1..5 = a0, a1, g0, g1, b  ->  edata order a0, a1, b, g0, g1
        vals = np.stack([row[:, 1], row[:, 2], row[:, 5], row[:, 3], row[:, 4]], axis=1)
        out[:, col:col + 5] = np.where(ok[:, None], vals, 0.0).astype(np.float32)
    return out


def _elem_props(ex, ey, ez, dims, h, dt, layers, abase, bbase, thr_damping, thr_vpvs, size=None, mat=None):
    """Per-element quantities of solver_init (psolve.c:3360-3473) for elements whose lowest corner
    is grid point (ex, ey, ez) of the h-grid and whose edge is size * h (size = 1 when None):
    eTable rows, lumped mass, a, dashpot terms.  Everything but the dashpots depends on the
    element's (material layer, size) class only: evaluated once per class with the reference's
    float / double sequence, then spread over the elements."""
    f32 = np.float32
    nx, ny, nz = dims
    E = ex.size
    sz = 1 if size is None else np.asarray(size, np.int64)
    dt, h = np.float64(dt), np.float64(h)      # a Python float would leave dt2 * edge in float32
    zc = (ez + 0.5 * sz) * h
    if mat is not None:                                           # material index given per element
        li = np.asarray(mat, np.int64)
    else:
        li = np.zeros(E, np.int64)                                # the last layer whose top is above the centre
        for k, (zt, _, _, _) in enumerate(layers):
            li[zc >= zt] = k
    del zc
    szs = np.unique(sz) if size is not None else np.array([1], np.int64)
    si = np.searchsorted(szs, sz) if size is not None else np.zeros(E, np.int64)
    cls = li * szs.size + si                                      # class of every element
    del li, si
    C = len(layers) * szs.size
    cl, cs = np.divmod(np.arange(C), szs.size)
    # ---- per class ------------------------------------------------------------------------------
    Vp, Vs, rho = np.empty(C, f32), np.empty(C, f32), np.empty(C, f32)
    for k, (_, vp, vs, r) in enumerate(layers):
        Vp[cl == k], Vs[cl == k], rho[cl == k] = vp, vs, r
    edge = (szs[cs] * h).astype(f32) if size is not None else np.full(C, h, f32)
    # mu_and_lambda (psolve.c:3236-3272): float products, then double
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
    eT = np.empty((C, 4))
    eT[:, 0] = dt2 * edge * mu / 9
    eT[:, 1] = dt2 * edge * lam / 9
    zeta = (f32(10) / Vs).astype(np.float64)            # 10 / edata->Vs is a float division
    zeta = np.minimum(zeta, thr_damping)
    a, b = zeta * abase, zeta * bbase
    eT[:, 2] = b * dt * edge * mu / 9
    eT[:, 3] = b * dt * edge * lam / 9
    # lumped mass and dashpots (psolve.c:3411-3473, 5752-5804)
    M = (rho * edge * edge * edge).astype(np.float64) / 8
    scale = (rho * (edge / f32(2)) * (edge / f32(2))).astype(np.float64)
    # ---- per element ----------------------------------------------------------------------------
    cVp, cVs, cscale = Vp, Vs, scale
    Vp, Vs, rho, edge, eT, M, a = Vp[cls], Vs[cls], rho[cls], edge[cls], eT[cls], M[cls], a[cls]
    # absorbing faces: x near/far, y near/far, z far; the top (z near) is free under HALFSPACE
    touch = ((ex == 0, ex + sz == nx), (ey == 0, ey + sz == ny), (np.zeros(E, bool), ez + sz == nz))
    boundary = touch[0][0] | touch[0][1] | touch[1][0] | touch[1][1] | touch[2][1]   # flag != 13
    # dashpots only on boundary elements: keep them sparse
    bi = np.nonzero(boundary)[0]
    nb_ = bi.size
    dash = np.zeros((nb_, 8, 3))
    if nb_:
        bits = np.zeros((nb_, 8), np.int64)
        for j in range(8):
            for ax in range(3):
                far = (j >> ax) & 1
                bits[:, j] |= (touch[ax][far][bi].astype(np.int64) << ax)
        nb = (bits & 1) + ((bits >> 1) & 1) + ((bits >> 2) & 1)
        Vpb, Vsb, sc = cVp[cls[bi]], cVs[cls[bi]], cscale[cls[bi]]
        vp_plus_2vs = (Vpb + f32(2) * Vsb).astype(np.float64)
        for c in range(3):
            on = ((bits >> c) & 1).astype(bool)
            vsel = np.where(on, Vpb[:, None], Vsb[:, None])                   # float
            two = (Vsb[:, None] + vsel).astype(np.float64) * sc[:, None]      # (Vs + Vp|Vs) float, * double
            one = vsel.astype(np.float64) * sc[:, None]
            three = np.broadcast_to((vp_plus_2vs * sc)[:, None], (nb_, 8))
            dash[:, :, c] = np.where(nb == 3, three, np.where(nb == 2, two, np.where(nb == 1, one, 0.0)))
    return dict(Vp=Vp, Vs=Vs, rho=rho, edge=edge, eT=eT, M=M, a=a, bidx=bi, dash=dash)


def _accumulate(nT, nodes8, pr, dt, exact, keep=None):
    """Add the elements' lumped-mass / damping / dashpot terms into nT (psolve.c:3445-3473).
    nodes8 = [E][8] node rows; keep = optional [E][8] mask of corners to include."""
    E = nodes8.shape[0]
    N = nT.shape[0]
    M, a = pr["M"], pr["a"]
    bi, dash = pr["bidx"], pr["dash"]
    flat = nodes8.reshape(-1)
    kf = None if keep is None else keep.reshape(-1)
    if exact:
        # the reference's exact sequence per (element, corner, axis): -= dt a M; -= dt dashpot
        # (boundary elements only); += M  |  mass2: -= dt a M; -= dt dashpot; += 2 M
        dd = np.zeros((E, 8, 3)); dd[bi] = dt * dash
        dd = dd.reshape(-1, 3)
        bnd = np.zeros(E, bool); bnd[bi] = True
        bnd = np.repeat(bnd, 8)
        ones = np.ones_like(bnd) if kf is None else kf
        np.add.at(nT[:, 0], flat[ones], np.repeat(M, 8)[ones])
        daM = np.repeat(dt * a * M, 8)
        for ax in range(3):
            for col, mult in ((4 + ax, 1.0), (1 + ax, 2.0)):
                idx = np.stack([flat, flat, flat], 1)
                val = np.stack([-daM, np.where(bnd, -dd[:, ax], 0.0), np.repeat(M * mult, 8)], 1)
                k3 = np.stack([ones, bnd & ones, ones], 1)
                np.add.at(nT[:, col], idx[k3], val[k3])
        return
    # grouped sums: per node sum(M) and sum(dt a M) over the incident (element, corner) pairs, one
    # corner column at a time (no 8 E temporaries); mass_minusaM = sum(M) - sum(dt a M) - dashpots
    sumM, sumaM = np.zeros(N), np.zeros(N)
    daM = dt * a * M
    for j in range(8):
        idx = np.ascontiguousarray(nodes8[:, j], dtype=np.intp)
        if keep is None:
            sumM += np.bincount(idx, M, N)
            sumaM += np.bincount(idx, daM, N)
        else:
            w = keep[:, j].astype(np.float64)
            sumM += np.bincount(idx, M * w, N)
            sumaM += np.bincount(idx, daM * w, N)
    nT[:, 0] += sumM
    base1, base2 = sumM - sumaM, 2 * sumM - sumaM
    bflat = nodes8[bi].reshape(-1)
    bw = None if keep is None else keep[bi].reshape(-1).astype(np.float64)
    for ax in range(3):
        d = dt * dash[:, :, ax].reshape(-1)
        dsum = np.bincount(bflat, d if bw is None else d * bw, N) if bflat.size else 0.0
        nT[:, 4 + ax] += base1 - dsum
        nT[:, 1 + ax] += base2 - dsum


def uniform_halfspace(nx: int, ny: int, nz: int, h: float, dt: float, freq: float = 1.0,
                      damping: int = RAYLEIGH, layers=((0.0, 6000.0, 3464.0, 2700.0),),
                      thr_damping: float = 0.05, thr_vpvs: float = 3.0, exact: bool = False,
                      part: tuple | None = None):
    """nx x ny x nz elements of edge h (x = north, y = east, z = depth).  layers = (ztop, Vp, Vs,
    rho) by depth of the element centre.  exact=True accumulates nTable in the reference's exact
    operation order (slow, for the bit-for-bit test); otherwise sums are grouped per element.

    part = (rank, nranks) returns that rank's share under octor's partition rule: the Morton
    ordered leaf list is cut into contiguous blocks (octor.c:685-746, 4940-4941); a rank harbors
    the corner nodes of its elements; a node belongs to the rank that owns the element whose
    lowest corner it is (far-boundary nodes pulled in by one tick, octor.c:5466-5475); the halo
    schedules follow schedule_build (psolve.c:4705-4863) and the owners' nTable rows include the
    sharers' partial sums in messenger order (mass exchange, psolve.c:3498-3507).
    Returns (HostMesh, info)."""
    rank, world = part if part is not None else (0, 1)
    mi = _MortonIndex(nx, ny, nz)
    Etot = mi.total
    lo, hi = block_low(rank, world, Etot), block_low(rank + 1, world, Etot)
    ex, ey, ez = mi.coords(np.arange(lo, hi, dtype=np.int64))
    E = ex.size
    # ---- harbored nodes in Morton order ----------------------------------------------------------
    x0, y0, z0 = int(ex.min()), int(ey.min()), int(ez.min())
    X, Y, Z = int(ex.max()) - x0 + 2, int(ey.max()) - y0 + 2, int(ez.max()) - z0 + 2
    lx, ly, lz = ex - x0, ey - y0, ez - z0
    full = E == (X - 1) * (Y - 1) * (Z - 1)
    if full:
        hf = np.ones((X, Y, Z), bool)
    else:
        hf = np.zeros((X, Y, Z), bool)
        for j in range(8):
            hf[lx + (j & 1), ly + ((j >> 1) & 1), lz + ((j >> 2) & 1)] = True
    def key(g, n):
        return np.where(g == n, 2 * n - 1, 2 * g)
    # harbored nodes in octor's order: Morton codes of the half-step keys, sorted in place and decoded
    if full:
        kx, ky, kz = np.meshgrid(key(np.arange(x0, x0 + X, dtype=np.int64), nx), key(np.arange(y0, y0 + Y, dtype=np.int64), ny),
                                 key(np.arange(z0, z0 + Z, dtype=np.int64), nz), indexing="ij")
        kx, ky, kz = kx.ravel(), ky.ravel(), kz.ravel()
    else:
        ix, iy, iz = np.nonzero(hf)
        kx, ky, kz = key(ix + x0, nx), key(iy + y0, ny), key(iz + z0, nz)
        del ix, iy, iz
    del hf
    small = 2 * max(nx, ny, nz) < 4096
    codes = morton3_fast(kx, ky, kz) if small else morton3(kx, ky, kz)
    del kx, ky, kz
    codes.sort()
    if small:
        kx, ky, kz = demorton3_fast(codes)
    else:
        kx, ky, kz = _compact1by2(codes), _compact1by2(codes >> np.uint64(1)), _compact1by2(codes >> np.uint64(2))
    del codes
    gx, gy, gz = ((kx + 1) >> 1).astype(np.int64), ((ky + 1) >> 1).astype(np.int64), ((kz + 1) >> 1).astype(np.int64)
    del kx, ky, kz
    ix, iy, iz = gx - x0, gy - y0, gz - z0
    N = gx.size
    SY, SX = Z, Y * Z
    nrank = np.full(X * SX, -1, np.int32)
    nrank[ix * SX + iy * SY + iz] = np.arange(N, dtype=np.int32)
    lnid = np.empty((E, 8), np.int32)
    base = lx * SX + ly * SY + lz
    for j in range(8):
        lnid[:, j] = nrank.take(base + ((j & 1) * SX + ((j >> 1) & 1) * SY + ((j >> 2) & 1)))
    del base
    nrank = nrank.reshape(X, Y, Z)
    # ---- solver tables from my own elements ------------------------------------------------------
    abase, bbase = compute_setab(damping, freq)
    args = ((nx, ny, nz), h, dt, layers, abase, bbase, thr_damping, thr_vpvs)
    pr = _elem_props(ex, ey, ez, *args)
    nT = np.zeros((N, 7))
    _accumulate(nT, lnid, pr, dt, exact)
    # ---- ownership, sharers, schedules -----------------------------------------------------------
    owner = np.full(N, rank, np.int32)
    msg = {k: MsgList() for k in ("dn_c", "dn_s", "an_c", "an_s")}
    share = np.zeros((0, 2), np.int32)
    if world > 1:
        # nodes all of whose in-domain adjacent elements are mine need no look-up
        mine = np.zeros((X + 1, Y + 1, Z + 1), bool)          # element grid x0-1 .. x0+X-1
        mine[lx + 1, ly + 1, lz + 1] = True
        gxs, gys, gzs = np.arange(x0 - 1, x0 + X), np.arange(y0 - 1, y0 + Y), np.arange(z0 - 1, z0 + Z)
        outside = ((gxs < 0) | (gxs >= nx))[:, None, None] | ((gys < 0) | (gys >= ny))[None, :, None] | \
                  ((gzs < 0) | (gzs >= nz))[None, None, :]
        ok = mine | outside
        all8 = np.ones((X, Y, Z), bool)
        for j in range(8):
            dx, dy, dz = j & 1, (j >> 1) & 1, (j >> 2) & 1
            all8 &= ok[1 - dx:1 - dx + X, 1 - dy:1 - dy + Y, 1 - dz:1 - dz + Z]
        cand = np.nonzero(~all8[ix, iy, iz])[0]               # ascending lnid
        del mine, outside, ok, all8
        cx, cy, cz = gx[cand], gy[cand], gz[cand]
        owner[cand] = block_owner(mi.index(np.minimum(cx, nx - 1), np.minimum(cy, ny - 1),
                                           np.minimum(cz, nz - 1)), world, Etot)
        # adjacent elements of the candidate nodes and their ranks
        adj_rank = np.full((cand.size, 8), -1, np.int64)
        adj_gidx = np.full((cand.size, 8), -1, np.int64)
        for j in range(8):
            ax_, ay_, az_ = cx - (j & 1), cy - ((j >> 1) & 1), cz - ((j >> 2) & 1)
            inb = (ax_ >= 0) & (ax_ < nx) & (ay_ >= 0) & (ay_ < ny) & (az_ >= 0) & (az_ < nz)
            g = mi.index(ax_[inb], ay_[inb], az_[inb])
            adj_gidx[inb, j] = g
            adj_rank[inb, j] = block_owner(g, world, Etot)
        mine_c = owner[cand] == rank
        # c-lists: nodes I harbor but do not own, per owner, ascending lnid; messengers are pushed
        # to the front of the list as they first appear (psolve.c:4733-4746)
        def make_list(nodes, peers):
            if nodes.size == 0:
                return MsgList()
            firsts = {}
            for nd, p in zip(nodes.tolist(), peers.tolist()):
                firsts.setdefault(p, None)
            order = list(firsts.keys())[::-1]
            maps = [nodes[peers == p] for p in order]
            return MsgList(np.array(order, np.int32), np.array([m.size for m in maps], np.int32),
                           np.concatenate(maps).astype(np.int32))
        notmine = cand[~mine_c]
        msg["an_c"] = make_list(notmine, owner[notmine].astype(np.int64))
        # s-lists: (node, sharer) for nodes I own.  A node's share list holds its sharers in the
        # order this rank discovered its neighbour ranks (octor.c:5701-5785 pushes onto the list
        # while walking the pctl list, itself pushed in discovery order by com_allocpctl,
        # octor.c:2639-2742: leaves in Morton order, 4x4x4 probe points per leaf, z outermost).
        disc = _discovery_order(mi, ex, ey, ez, (x0, y0, z0), (X, Y, Z), rank, world)
        pos = {p: i for i, p in enumerate(disc)}
        sh_nodes, sh_peers = [], []
        ar = adj_rank[mine_c]
        own_nodes = cand[mine_c]
        ar_sorted = np.sort(ar, axis=1)
        for col in range(8):
            r = ar_sorted[:, col]
            new = (r >= 0) & (r != rank)
            if col:
                new &= r != ar_sorted[:, col - 1]
            sh_nodes.append(own_nodes[new]); sh_peers.append(r[new])
        sh_nodes, sh_peers = np.concatenate(sh_nodes), np.concatenate(sh_peers)
        sh_pos = np.array([pos[int(p)] for p in np.unique(sh_peers)], np.int64)
        lut = np.zeros(world, np.int64); lut[np.unique(sh_peers)] = sh_pos
        o = np.lexsort((lut[sh_peers], sh_nodes))             # by node, then by discovery order
        sh_nodes, sh_peers = sh_nodes[o], sh_peers[o]
        share = np.stack([sh_nodes, sh_peers], 1).astype(np.int32)
        msg["an_s"] = make_list(sh_nodes, sh_peers)
        # mass exchange (psolve.c:3505-3507): owners add each sharer's partial sums, messenger
        # after messenger; a sharer's partial is its own element loop over the elements that
        # touch the node, in its local (Morton) element order
        for p in msg["an_s"].peer.tolist():
            sel = (adj_rank == p) & mine_c[:, None]           # [cand][8]: ghost element of rank p
            ci, cj = np.nonzero(sel)
            gid = adj_gidx[ci, cj]
            ug, inv = np.unique(gid, return_inverse=True)     # ghost elements in p's element order
            gex, gey, gez = mi.coords(ug)
            gpr = _elem_props(gex, gey, gez, *args)
            # corner c of ghost element g is node (g + corner offset); corner index = the j whose
            # offset leads from the element to the node: node = elem + (j&1, ...)  =>  j = cj
            nodes8 = np.zeros((ug.size, 8), np.int64)
            keep = np.zeros((ug.size, 8), bool)
            nodes8[inv, cj] = cand[ci]
            keep[inv, cj] = True
            partial = np.zeros((N, 7))
            _accumulate(partial, nodes8, gpr, dt, exact, keep)
            rows = msg["an_s"].mapping[_slice_of(msg["an_s"], p)]
            nT[rows] += partial[rows]
    edata = np.zeros((E, 14), np.float32)
    edata[:, 0], edata[:, 1], edata[:, 2], edata[:, 3] = pr["edge"], pr["Vp"], pr["Vs"], pr["rho"]
    if damping == BKT:
        edata[:, 4:14] = bkt_coefficients(pr["Vp"], pr["Vs"])
    K1, K2 = compute_K()
    mesh = HostMesh(lnid, pr["eT"], nT, np.zeros((0, 6), np.int32), edata, K1, K2,
                    msg["dn_c"], msg["dn_s"], msg["an_c"], msg["an_s"])
    node_lin = (gx * (ny + 1) + gy) * (nz + 1) + gz
    info = dict(E=E, N=N, abase=abase, bbase=bbase, node_order=node_lin, node_xyz=(gx, gy, gz),
                elem_xyz=(ex, ey, ez), elem_geid=np.arange(lo, hi, dtype=np.int64), origin=(x0, y0, z0),
                dims=(nx, ny, nz), h=h, owner=owner, share=share, rank=rank, nranks=world,
                etotal=Etot)
    return mesh, info


def column_regions(nx: int, ny: int, nz: int, world: int):
    """octor's partition of a laterally uniform (banded) mesh over `world` ranks, as rectangles of
    the h-grid: the leaves in Morton order are cut into equal blocks (octor.c:685-746); with every
    vertical column of aligned c x c cells (c a power of two >= nz) holding the same number of leaves
    and being contiguous in Morton order, rank r gets `per` consecutive cells.  Returns
    [(x0, x1, y0, y1)] by rank; raises when the blocks would not be whole rectangles."""
    S = 1
    while S < max(nx, ny, nz):
        S *= 2
    c = S
    while c >= max(nz, 1):
        if nx % c == 0 and ny % c == 0:
            ncell = (nx // c) * (ny // c)
            if ncell >= world and ncell % world == 0:
                break
        c //= 2
    else:
        raise ValueError(f"no column partition of {nx}x{ny}x{nz} over {world} ranks (cells must be >= nz wide)")
    cx, cy = np.meshgrid(np.arange(nx // c), np.arange(ny // c), indexing="ij")
    cx, cy = cx.ravel(), cy.ravel()
    o = np.argsort(morton3(cx, cy, np.zeros_like(cx)), kind="stable")
    cx, cy = cx[o], cy[o]
    per = cx.size // world
    out = []
    for r in range(world):
        gx, gy = cx[r * per:(r + 1) * per], cy[r * per:(r + 1) * per]
        x0, x1, y0, y1 = int(gx.min()) * c, (int(gx.max()) + 1) * c, int(gy.min()) * c, (int(gy.max()) + 1) * c
        if (x1 - x0) * (y1 - y0) != per * c * c:
            raise ValueError("a rank's block of columns is not a rectangle")
        out.append((x0, x1, y0, y1))
    return out


def _graded_leaves(bands, ztop, x0, x1, y0, y1):
    """Leaves of the banded mesh inside [x0,x1) x [y0,y1), Morton order of the lowest corner: the
    Morton codes are sorted in place and decoded again (no index sort, no gathers)."""
    codes = []
    small = max(x1, y1, ztop[-1]) < 4096
    enc = morton3_fast if small else morton3
    for (nl, sz), z0 in zip(bands, ztop):
        gx, gy, gz = np.meshgrid(np.arange(x0, x1, sz, dtype=np.int32), np.arange(y0, y1, sz, dtype=np.int32),
                                 np.arange(z0, z0 + nl * sz, sz, dtype=np.int32), indexing="ij")
        codes.append(enc(gx.ravel(), gy.ravel(), gz.ravel()))
        del gx, gy, gz
    codes = np.concatenate(codes)
    codes.sort()
    if small:
        ex, ey, ez = demorton3_fast(codes)
    else:
        ex, ey, ez = _compact1by2(codes), _compact1by2(codes >> np.uint64(1)), _compact1by2(codes >> np.uint64(2))
    del codes
    size_of_z = np.zeros(ztop[-1] + 1, ex.dtype)
    for (nl, sz), z0 in zip(bands, ztop):
        size_of_z[z0:z0 + nl * sz] = sz
    return ex, ey, ez, size_of_z[ez]


def _graded_nodes(bands, ztop, x0, x1, y0, y1, dims):
    """Mesh nodes inside the closed box [x0,x1] x [y0,y1], in octor's node order (Z-order of the
    half-step keys 2 g, or 2 n - 1 on a far domain face): keys sorted in place and decoded."""
    nx, ny, nz = dims

    def key(g, n):
        return np.where(g == n, 2 * n - 1, 2 * g)

    def unkey(k):
        return (k + 1) >> 1                                     # 2 g -> g, 2 n - 1 -> n
    codes = []
    small = 2 * max(nx, ny, nz) < 4096
    enc = morton3_fast if small else morton3
    for k, ((nl, sz), z0) in enumerate(zip(bands, ztop)):
        first = z0 if k == 0 else z0 + sz                       # a band's top plane belongs to the finer band above
        gx, gy, gz = np.meshgrid(key(np.arange(x0, x1 + 1, sz, dtype=np.int32), nx),
                                 key(np.arange(y0, y1 + 1, sz, dtype=np.int32), ny),
                                 key(np.arange(first, z0 + nl * sz + 1, sz, dtype=np.int32), nz), indexing="ij")
        codes.append(enc(gx.ravel(), gy.ravel(), gz.ravel()))
        del gx, gy, gz
    codes = np.concatenate(codes)
    codes.sort()
    if small:
        kx, ky, kz = demorton3_fast(codes)
    else:
        kx, ky, kz = _compact1by2(codes), _compact1by2(codes >> np.uint64(1)), _compact1by2(codes >> np.uint64(2))
    return unkey(kx), unkey(ky), unkey(kz)


def _graded_dangling(px, py, pz, bands, ztop):
    """Dangling nodes among the given nodes of the banded mesh: (idx, deps, ax, ay) with idx ascending,
    deps = 2 (hangs on an edge) or 4 (in a face) and the anchors' x, y ([n][4], same z) in the
    reference's list order (descending Z-order).  Only nodes on the planes between bands qualify."""
    I, D, AX, AY = [], [], [], []
    for (_, sf), (_, sc), zp in zip(bands[:-1], bands[1:], ztop[1:]):
        on = np.nonzero(pz == zp)[0]
        qx, qy = px[on].astype(np.int64), py[on].astype(np.int64)
        ox, oy = (qx % sc) != 0, (qy % sc) != 0
        sel = ox | oy
        on, qx, qy, ox, oy = on[sel], qx[sel], qy[sel], ox[sel], oy[sel]
        both = ox & oy
        xh, xl = np.where(ox, qx + sf, qx), np.where(ox, qx - sf, qx)
        yh, yl = np.where(oy, qy + sf, qy), np.where(oy, qy - sf, qy)
        # 2 anchors: (high, low) along the hanging axis; 4 anchors: (xh,yh) (xl,yh) (xh,yl) (xl,yl)
        ax = np.stack([xh, xl, np.where(both, xh, 0), np.where(both, xl, 0)], 1)
        ay = np.stack([yh, np.where(both, yh, yl), np.where(both, yl, 0), np.where(both, yl, 0)], 1)
        I.append(on); D.append(np.where(both, 4, 2).astype(np.int32)); AX.append(ax); AY.append(ay)
    if not I:
        z = np.zeros(0, np.int64)
        return z, z.astype(np.int32), np.zeros((0, 4), np.int64), np.zeros((0, 4), np.int64)
    I, D, AX, AY = np.concatenate(I), np.concatenate(D), np.concatenate(AX), np.concatenate(AY)
    o = np.argsort(I, kind="stable")
    return I[o], D[o], AX[o], AY[o]


def graded_halfspace(nx: int, ny: int, bands, h: float, dt: float, freq: float = 1.0,
                     damping: int = RAYLEIGH, layers=((0.0, 6000.0, 3464.0, 2700.0),),
                     thr_damping: float = 0.05, thr_vpvs: float = 3.0, exact: bool = False,
                     part: tuple | None = None):
    """Adaptive (2:1 balanced) octree mesh of a depth-banded half-space, in the layout octor +
    solver_init produce for it (bit-exact on tests/golden/graded{2,3}_*.npz, tests/test_meshgen.py).
    The h-grid has nx x ny points per horizontal plane; bands = ((nlayers, size), ...) from the
    surface down: nlayers layers of cubic elements of edge size * h, size doubling from one band to
    the next (what octor's refinement + balancing yield when Vs grows with depth by band).  layers =
    (ztop, Vp, Vs, rho) by the depth of the element centre.

    * leaves in preorder = Morton order of their lowest corner (octor.c:5362-5374, 5505);
    * nodes = the distinct leaf corners in Z-order with the far faces pulled in by one tick
      (octor.c:3034, 5466-5475, 6166);
    * a node on the plane between two bands that is not a corner of the coarse side is dangling:
      on a coarse edge it hangs on that edge's 2 end nodes, inside a coarse face on its 4 corners;
      the anchor list is in descending Z-order (octor pushes at the head, octor.c:5863-5991);
      dnodeTable holds the OWNED dangling nodes in ascending ldnid;
    * nTable: element loop of solver_init, then compute_adjust(DISTRIBUTION) over all 7 columns
      in dnodeTable order (psolve.c:3503-3504, 5936-5978).

    part = (rank, nranks): that rank's share under octor's block partition of the Morton-ordered
    leaf list, for domains whose blocks are whole columns of cells (column_regions): local
    elements and geid, harbored nodes (the corners of the local elements), ownership (the rank whose
    region holds the node's pixel, far faces pulled in, octor.c:5466-5475), share lists in the
    order this rank discovers its neighbours (com_allocpctl, octor.c:2639-2742), dangling / anchored
    schedules (schedule_build, psolve.c:4705-4863).  On a partitioned mesh nTable holds, for every
    harbored node, the COMPLETE sums (all ranks' elements, all hanging-node transfers) evaluated in
    single-rank order: owned rows equal the reference's up to the order of additions (1e-15), rows of
    non-owned nodes are whole instead of partial (the reference overwrites those nodes after every
    update, psolve.c:4130-4154, so their mass never reaches a result).
    Returns (HostMesh, info)."""
    bands = [(int(a), int(b)) for a, b in bands]
    for (_, s0), (_, s1) in zip(bands[:-1], bands[1:]):
        if s1 != 2 * s0:
            raise ValueError("element size must double from one band to the next")
    smax = bands[-1][1]
    if nx % smax or ny % smax:
        raise ValueError("nx, ny must be multiples of the coarsest element size")
    ztop = [0]
    for nl, sz in bands:
        if ztop[-1] % sz:
            raise ValueError("a band must start on a multiple of its element size")
        ztop.append(ztop[-1] + nl * sz)
    nz = ztop[-1]
    dims = (nx, ny, nz)
    rank, world = part if part is not None else (0, 1)
    regions = column_regions(nx, ny, nz, world) if world > 1 else [(0, nx, 0, ny)]
    x0, x1, y0, y1 = regions[rank]
    if (x0 % smax) or (x1 % smax) or (y0 % smax) or (y1 % smax):
        raise ValueError("partition columns must be aligned to the coarsest element size")
    # ---- leaves -----------------------------------------------------------------------------------
    ex, ey, ez, es = _graded_leaves(bands, ztop, x0, x1, y0, y1)
    E = ex.size
    # ---- nodes ------------------------------------------------------------------------------------
    px, py, pz = _graded_nodes(bands, ztop, x0, x1, y0, y1, dims)
    N = px.size
    SY, SX = nz + 1, (y1 - y0 + 1) * (nz + 1)                    # strides of the dense node look-up
    nrank = np.full((x1 - x0 + 1) * SX, -1, np.int32)
    nrank[(px - x0).astype(np.int64) * SX + (py - y0) * SY + pz] = np.arange(N, dtype=np.int32)
    lnid = np.empty((E, 8), np.int32)
    base = (ex - x0).astype(np.int64) * SX + (ey - y0) * SY + ez
    es64 = es.astype(np.int64)
    for j in range(8):
        lnid[:, j] = nrank.take(base + es64 * ((j & 1) * SX + ((j >> 1) & 1) * SY + ((j >> 2) & 1)))
    del base, es64
    assert lnid.min() >= 0
    # ---- ownership --------------------------------------------------------------------------------
    owner = np.full(N, rank, np.int32)
    if world > 1:
        qx, qy = np.minimum(px, nx - 1), np.minimum(py, ny - 1)
        for r, (a0, a1, b0, b1) in enumerate(regions):
            owner[(qx >= a0) & (qx < a1) & (qy >= b0) & (qy < b1)] = r
    mine = owner == rank
    # ---- dangling nodes ---------------------------------------------------------------------------
    didx, ddeps, ax, ay = _graded_dangling(px, py, pz, bands, ztop)
    deps = np.zeros(N, np.int32)
    deps[didx] = ddeps
    om = mine[didx]                                              # the table holds the OWNED dangling nodes
    dsel, dd, ax, ay = didx[om], ddeps[om], ax[om], ay[om]       # ascending lnid
    dnode = np.full((dsel.size, 6), -1, np.int32)
    dnode[:, 0], dnode[:, 1] = dsel, dd
    for a in range(4):
        has = dd > a
        dnode[has, 2 + a] = nrank.take((ax[has, a] - x0) * SX + (ay[has, a] - y0) * SY + pz[dsel[has]])
    assert (dnode[:, 2:][dnode[:, 2:] != -1] >= 0).all()
    # ---- share lists and schedules ----------------------------------------------------------------
    msg = {k: MsgList() for k in ("dn_c", "dn_s", "an_c", "an_s")}
    share = np.zeros((0, 2), np.int32)
    if world > 1:
        def make_list(nodes, peers):
            # messengers are pushed at the head of the list as they first appear (psolve.c:4733-4746)
            if nodes.size == 0:
                return MsgList()
            order = list(dict.fromkeys(peers.tolist()))[::-1]
            maps = [nodes[peers == p_] for p_ in order]
            return MsgList(np.array(order, np.int32), np.array([m.size for m in maps], np.int32),
                           np.concatenate(maps).astype(np.int32))
        anch = deps == 0
        nm = np.nonzero(~mine)[0]
        msg["an_c"] = make_list(nm[anch[nm]], owner[nm[anch[nm]]].astype(np.int64))
        msg["dn_c"] = make_list(nm[~anch[nm]], owner[nm[~anch[nm]]].astype(np.int64))
        disc = _column_discovery_order(ex, ey, es, regions, rank, dims)
        pos = {p_: i for i, p_ in enumerate(disc)}
        sh_nodes, sh_peers, sh_pos = [], [], []
        own = np.nonzero(mine)[0]
        for r, (a0, a1, b0, b1) in enumerate(regions):
            if r == rank:
                continue
            hit = own[(px[own] >= a0) & (px[own] <= a1) & (py[own] >= b0) & (py[own] <= b1)]
            if hit.size:
                sh_nodes.append(hit); sh_peers.append(np.full(hit.size, r, np.int64))
                sh_pos.append(np.full(hit.size, pos[r], np.int64))
        if sh_nodes:
            sh_nodes, sh_peers, sh_pos = np.concatenate(sh_nodes), np.concatenate(sh_peers), np.concatenate(sh_pos)
            o = np.lexsort((sh_pos, sh_nodes))                   # by node, then by discovery order
            sh_nodes, sh_peers = sh_nodes[o], sh_peers[o]
            share = np.stack([sh_nodes, sh_peers], 1).astype(np.int32)
            msg["an_s"] = make_list(sh_nodes[anch[sh_nodes]], sh_peers[anch[sh_nodes]])
            msg["dn_s"] = make_list(sh_nodes[~anch[sh_nodes]], sh_peers[~anch[sh_nodes]])
    # ---- solver tables ----------------------------------------------------------------------------
    abase, bbase = compute_setab(damping, freq)
    args = (dims, h, dt, layers, abase, bbase, thr_damping, thr_vpvs)
    pr = _elem_props(ex, ey, ez, *args, size=es)
    if world == 1:
        nT = np.zeros((N, 7))
        _accumulate(nT, lnid, pr, dt, exact)
        _distribute(nT, dnode)
    else:
        # complete sums on the region widened by two coarse elements: every element and every
        # dangling node that feeds a harbored node lies inside
        m = 2 * smax
        X0, X1, Y0, Y1 = max(0, x0 - m), min(nx, x1 + m), max(0, y0 - m), min(ny, y1 + m)
        gx, gy, gz, gs = _graded_leaves(bands, ztop, X0, X1, Y0, Y1)
        qx, qy, qz = _graded_nodes(bands, ztop, X0, X1, Y0, Y1, dims)
        BY, BX = nz + 1, (Y1 - Y0 + 1) * (nz + 1)
        big = np.full((X1 - X0 + 1) * BX, -1, np.int32)
        big[(qx - X0).astype(np.int64) * BX + (qy - Y0) * BY + qz] = np.arange(qx.size, dtype=np.int32)
        l8 = np.empty((gx.size, 8), np.int32)
        gb, gs64 = (gx - X0).astype(np.int64) * BX + (gy - Y0) * BY + gz, gs.astype(np.int64)
        for j in range(8):
            l8[:, j] = big.take(gb + gs64 * ((j & 1) * BX + ((j >> 1) & 1) * BY + ((j >> 2) & 1)))
        del gb, gs64
        full = np.zeros((qx.size, 7))
        _accumulate(full, l8, _elem_props(gx, gy, gz, *args, size=gs), dt, exact)
        qi, dq, bx_, by_ = _graded_dangling(qx, qy, qz, bands, ztop)
        inside = np.ones(qi.size, bool)
        for a in range(4):                                       # anchors inside the widened box
            inside &= (dq <= a) | ((bx_[:, a] >= X0) & (bx_[:, a] <= X1) & (by_[:, a] >= Y0) & (by_[:, a] <= Y1))
        qi, dq, bx_, by_ = qi[inside], dq[inside], bx_[inside], by_[inside]
        dn = np.full((qi.size, 6), -1, np.int32)
        dn[:, 0], dn[:, 1] = qi, dq
        for a in range(4):
            has = dq > a
            dn[has, 2 + a] = big.take((bx_[has, a] - X0) * BX + (by_[has, a] - Y0) * BY + qz[qi[has]])
        _distribute(full, dn)
        nT = full[big.take((px - X0).astype(np.int64) * BX + (py - Y0) * BY + pz)]
    edata = np.zeros((E, 14), np.float32)
    edata[:, 0], edata[:, 1], edata[:, 2], edata[:, 3] = pr["edge"], pr["Vp"], pr["Vs"], pr["rho"]
    if damping == BKT:
        edata[:, 4:14] = bkt_coefficients(pr["Vp"], pr["Vs"])
    K1, K2 = compute_K()
    mesh = HostMesh(lnid, pr["eT"], nT, dnode, edata, K1, K2, msg["dn_c"], msg["dn_s"], msg["an_c"], msg["an_s"])
    etotal = sum(nl * (nx // sz) * (ny // sz) for nl, sz in bands)
    lo = block_low(rank, world, etotal)
    assert block_low(rank + 1, world, etotal) - lo == E
    info = dict(E=E, N=N, D=int(dnode.shape[0]), abase=abase, bbase=bbase, node_xyz=(px, py, pz),
                node_order=(px * (ny + 1) + py) * (nz + 1) + pz, elem_xyz=(ex, ey, ez), elem_size=es,
                elem_geid=np.arange(lo, lo + E, dtype=np.int64), origin=(x0, y0, 0), dims=dims, h=h,
                owner=owner, share=share, anchored=deps == 0, rank=rank, nranks=world, etotal=etotal,
                bands=bands, region=(x0, x1, y0, y1))
    return mesh, info


def _distribute(nT, dnode):
    """compute_adjust(DISTRIBUTION) on nTable (psolve.c:3503-3504, 5943-5978): every dangling node of
    the table adds value / deps to each of its anchors, in table order, anchors in list order."""
    if not dnode.shape[0]:
        return
    deps = dnode[:, 1].astype(np.int64)
    d = nT[dnode[:, 0]] / deps[:, None].astype(np.float64)               # darray = myvalue / deps
    slot = np.arange(4)[None, :] < deps[:, None]
    np.add.at(nT, dnode[:, 2:6][slot], np.repeat(d, deps, axis=0))


def _column_discovery_order(ex, ey, es, regions, rank, dims):
    """Neighbour ranks in the order com_allocpctl (octor.c:2639-2742) first meets them on a column
    partition: local leaves in Morton order, per leaf 4 x 4 x 4 probe points half an edge apart starting
    half an edge below the lowest corner (z outermost, x innermost), points outside the domain
    skipped.  The rank of a probe point depends on x, y only."""
    nx, ny, _ = dims
    x0, x1, y0, y1 = regions[rank]
    near = np.nonzero((ex - es < x0) | (ex + 2 * es > x1) | (ey - es < y0) | (ey + 2 * es > y1))[0]
    first = {}
    for j in range(4):
        for i in range(4):
            # doubled coordinates: 2 x - s + s i
            p2x, p2y = 2 * ex[near] - es[near] + es[near] * i, 2 * ey[near] - es[near] + es[near] * j
            ok = (p2x >= 0) & (p2x < 2 * nx) & (p2y >= 0) & (p2y < 2 * ny)
            fx, fy = p2x // 2, p2y // 2
            for r, (a0, a1, b0, b1) in enumerate(regions):
                if r == rank:
                    continue
                hit = ok & (fx >= a0) & (fx < a1) & (fy >= b0) & (fy < b1)
                if hit.any():
                    kmin = int((near[hit].astype(np.int64) * 16 + (j * 4 + i)).min())
                    if r not in first or kmin < first[r]:
                        first[r] = kmin
    return [r for r, _ in sorted(first.items(), key=lambda kv: kv[1])]


def _discovery_order(mi, ex, ey, ez, origin, shape, rank, world):
    """Neighbour ranks in the order com_allocpctl (octor.c:2639-2742) first meets them."""
    nx, ny, nz = mi.dims
    x0, y0, z0 = origin
    X, Y, Z = shape
    mine = np.zeros((X + 1, Y + 1, Z + 1), bool)
    mine[ex - x0 + 1, ey - y0 + 1, ez - z0 + 1] = True
    gxs, gys, gzs = np.arange(x0 - 1, x0 + X), np.arange(y0 - 1, y0 + Y), np.arange(z0 - 1, z0 + Z)
    ok = mine | ((gxs < 0) | (gxs >= nx))[:, None, None] | ((gys < 0) | (gys >= ny))[None, :, None] | \
        ((gzs < 0) | (gzs >= nz))[None, None, :]
    # elements with a foreign in-domain neighbour among the 26 around them
    pad = np.pad(ok, 1, constant_values=True)
    inner = np.ones_like(ok)
    for dz in (0, 1, 2):
        for dy in (0, 1, 2):
            for dx in (0, 1, 2):
                inner &= pad[dx:dx + X + 1, dy:dy + Y + 1, dz:dz + Z + 1]
    bsel = np.nonzero(~inner[ex - x0 + 1, ey - y0 + 1, ez - z0 + 1])[0]     # local element order
    bx, by, bz = ex[bsel], ey[bsel], ez[bsel]
    first = {}
    d = (-1, 0, 0, 1)
    for k in range(4):
        for j in range(4):
            for i in range(4):
                px, py, pz = bx + d[i], by + d[j], bz + d[k]
                inb = (px >= 0) & (px < nx) & (py >= 0) & (py < ny) & (pz >= 0) & (pz < nz)
                if not inb.any():
                    continue
                r = block_owner(mi.index(px[inb], py[inb], pz[inb]), world, mi.total)
                keyv = bsel[inb].astype(np.int64) * 64 + (k * 16 + j * 4 + i)
                for p in np.unique(r):
                    if p == rank:
                        continue
                    m = int(keyv[r == p].min())
                    if p not in first or m < first[p]:
                        first[int(p)] = m
    return [p for p, _ in sorted(first.items(), key=lambda kv: kv[1])]


def _slice_of(ml: MsgList, peer: int) -> slice:
    i = int(np.nonzero(ml.peer == peer)[0][0])
    off = int(ml.nodes[:i].sum())
    return slice(off, off + int(ml.nodes[i]))


def node_index(info: dict, ix: int, iy: int, iz: int) -> int:
    """Local node id of grid point (ix, iy, iz)."""
    nx, ny, nz = info["dims"]
    lin = (ix * (ny + 1) + iy) * (nz + 1) + iz
    return int(np.nonzero(info["node_order"] == lin)[0][0])


def element_index(info: dict, ex: int, ey: int, ez: int) -> int:
    """Local element id of the element whose lowest corner is grid point (ex, ey, ez), counted
    from the lowest corner of this rank's bounding box."""
    X, Y, Z = info["elem_xyz"]
    x0, y0, z0 = info["origin"]
    hit = np.nonzero((X == ex + x0) & (Y == ey + y0) & (Z == ez + z0))[0]
    if hit.size != 1:
        raise ValueError(f"element ({ex},{ey},{ez}) is not on this rank")
    return int(hit[0])
