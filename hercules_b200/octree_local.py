"""Per-rank octree meshes: every rank builds ITS Morton block of the mesh and the one-cell
neighbourhood it needs, never the whole mesh.

octree.octree_halfspace_part builds the whole mesh on every rank and cuts it (bit-exact against 2-, 3- and
4-rank runs of the unmodified reference, tests/test_octree.py) -- fine up to a few 10 M elements.  The
reference itself never holds the whole mesh on one rank: octor refines, balances, partitions and extracts
in a distributed fashion (octor.c:4337-4700, 685-746, 5268-6645).  This module reaches the same tables from
local work only:

1. the domain is cut into *coarse cells* of the largest admissible leaf edge S (octor's multi-rank bootstrap
   limit, octree.bootstrap_size); in Morton order the global leaf list is the concatenation of the cells'
   leaves, so a per-cell leaf COUNT fixes every leaf's global index (geid) and with it octor's equal blocks;
2. the balanced refinement inside a cell depends on the material model within one cell edge of it only
   (a leaf of edge s can only be split by finer leaves closer than s: the ripple grows by a factor of two per
   hop), so refining + balancing a set of cells together with their 26-neighbour ring gives the EXACT
   leaves of the whole-domain mesh on that set.  Counts: every rank does 1/world of the cells, the counts are
   all-gathered (one int per coarse cell);
3. a rank then builds the exact leaves of the cells its block touches plus one ring (X).  Everything octor
   derives for a rank -- the corners of its elements, the nodes it owns (held by one of its leaves), the
   anchors of the dangling nodes it owns, who else harbors an owned node (directly, or only as an anchor of a
   dangling node IT owns), the neighbour discovery order of com_allocpctl -- involves only leaves that touch
   the closure of the rank's cells, i.e. leaves of X;
4. extraction, solver_init's tables and the partition bookkeeping then run on X exactly as
   octree.extract / octree.partition run on the whole mesh, with the global leaf index taken from (1).

The heavy steps have native counterparts in csrc/hmesh.cpp (libhercules_mesh.so, include/hercules_mesh.h): refine +
balance per chunk, node extraction per chunk, lnid, the per-node mass sums and the neighbour discovery order.  They
are used when the material model is given on a grid (`model_cell` / `model`); the numpy restatement of every step
stays here and in octree.py as the checker.  Chunks of cells are independent, so the passes run on a thread pool
(ctypes drops the GIL).

tests/test_octree_local.py pins the result against the reference's multi-rank goldens and against the whole-mesh
cut, table by table, on the reference's models and on random ones.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from itertools import product

import numpy as np

from . import _lib
from . import meshgen as mg
from . import octree as oc
from .solver import HostMesh, MsgList, RAYLEIGH, BKT


class CoarseGrid:
    """The coarse cells (edge S, a power of two, in units of h) of a box, in Morton order of their lowest corners."""

    def __init__(self, dims, S: int):
        nx, ny, nz = dims
        if S & (S - 1) or nx % S or ny % S or nz % S:
            raise ValueError("the coarse edge must be a power of two that divides the domain")
        self.dims, self.S = dims, S
        self.shift = np.uint64(3 * (S.bit_length() - 1))
        gx, gy, gz = np.meshgrid(np.arange(0, nx, S), np.arange(0, ny, S), np.arange(0, nz, S), indexing="ij")
        gx, gy, gz = (g.ravel().astype(np.int64) for g in (gx, gy, gz))
        key = oc._code(gx, gy, gz) >> self.shift
        o = np.argsort(key, kind="stable")
        self.x, self.y, self.z, self.key = gx[o], gy[o], gz[o], key[o]
        self.n = self.key.size
        self.idx3 = np.empty((nx // S, ny // S, nz // S), np.int32)              # cell coordinates -> Morton position
        self.idx3[self.x // S, self.y // S, self.z // S] = np.arange(self.n, dtype=np.int32)

    def cell_of_code(self, codes):
        """Index of the cell that holds the octant with this Morton code."""
        return np.searchsorted(self.key, codes >> self.shift)

    def ring(self, ids):
        """ids and their 26 neighbours inside the domain: sorted, unique."""
        cx, cy, cz = self.idx3.shape
        i, j, k = self.x[ids] // self.S, self.y[ids] // self.S, self.z[ids] // self.S
        mark = np.zeros(self.n, bool)
        for dx, dy, dz in product((-1, 0, 1), repeat=3):
            qi, qj, qk = i + dx, j + dy, k + dz
            ok = (qi >= 0) & (qi < cx) & (qj >= 0) & (qj < cy) & (qk >= 0) & (qk < cz)
            mark[self.idx3[qi[ok], qj[ok], qk[ok]]] = True
        return np.nonzero(mark)[0]


def _chunk_leaves(grid: CoarseGrid, sub, vs_of, factor_h):
    """Exact leaves of the whole-domain mesh inside the cells `sub` (ascending): refine + balance on sub and
    its ring, keep what lies in sub.  Returns (codes ascending, sizes, leaves per cell of sub)."""
    reg = grid.ring(sub)
    sets = oc.balance(oc.refine_cells((grid.x[reg], grid.y[reg], grid.z[reg]), grid.S, vs_of, factor_h), grid.dims,
                      as_codes=True)
    codes = np.concatenate(list(sets.values()))
    sizes = np.concatenate([np.full(c.size, s, np.int64) for s, c in sets.items()])
    skey = grid.key[sub]
    ck = codes >> grid.shift
    pos = np.minimum(np.searchsorted(skey, ck), sub.size - 1)
    keep = skey[pos] == ck
    codes, sizes, pos = codes[keep], sizes[keep], pos[keep]
    o = np.argsort(codes, kind="stable")
    return codes[o], sizes[o], np.bincount(pos, minlength=sub.size)


# ---- native primitives (csrc/hmesh.cpp -> libhercules_mesh.so) ------------------------------------------

_MESH_SO = _lib.PKG / "libhercules_mesh.so"
_mesh = None


def mesh_lib() -> C.CDLL:
    global _mesh
    if _mesh is None:
        if not _MESH_SO.exists():
            raise RuntimeError(f"{_MESH_SO} is missing: run hercules_b200.build()")
        L = C.CDLL(str(_MESH_SO))
        L.hmesh_chunk_leaves.restype = C.c_int
        L.hmesh_chunk_leaves.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p]
        L.hmesh_chunk_nodes.restype = C.c_int
        L.hmesh_chunk_nodes.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
        L.hmesh_lnid.restype = C.c_int
        L.hmesh_lnid.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.hmesh_corner_sums.restype = C.c_int
        L.hmesh_corner_sums.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.hmesh_discovery.restype = C.c_int
        L.hmesh_discovery.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.hmesh_abi_version.restype, L.hmesh_abi_version.argtypes = C.c_int, []
        L.hmesh_free.restype, L.hmesh_free.argtypes = None, [C.c_void_p]
        _mesh = L
    return _mesh


class GridModel:
    """A material model that is piecewise constant on cells of `cl` h (what a CVM etree is to octor): the
    material index per cell as a dense uint8 array, plus Vs per material -- the form the native primitives read."""

    def __init__(self, dims, cl: int, mat_of, vs_tab, slab: int = 16):
        nx, ny, nz = dims
        self.cl = int(cl)
        g = tuple(-(-n // cl) for n in dims)
        self.grid = np.empty(g, np.uint8)
        yc, zc = (np.arange(g[1]) + 0.5) * cl, (np.arange(g[2]) + 0.5) * cl
        for i0 in range(0, g[0], slab):                           # slabs keep the temporaries of mat_of small
            xc = (np.arange(i0, min(i0 + slab, g[0])) + 0.5) * cl
            X, Y, Z = np.meshgrid(xc, yc, zc, indexing="ij")
            self.grid[i0:i0 + slab] = mat_of(X, Y, Z)
        self.gdims = np.array(g, np.int64)
        self.vs = np.ascontiguousarray(vs_tab, np.float64)


def _chunk_leaves_native(grid: CoarseGrid, sub, model: GridModel, factor_h: float, want_leaves: bool = True):
    L = mesh_lib()
    subc = np.ascontiguousarray(np.stack([grid.x[sub], grid.y[sub], grid.z[sub]], 1), np.int32)
    dims = np.array(grid.dims, np.int32)
    per_cell = np.zeros(sub.size, np.int64)
    pc, ps, n = C.c_void_p(), C.c_void_p(), C.c_int64()
    rc = L.hmesh_chunk_leaves(dims.ctypes.data, grid.S, 1, model.grid.ctypes.data, model.gdims.ctypes.data, model.cl,
                              model.vs.ctypes.data, float(factor_h), None, 0, subc.ctypes.data, sub.size,
                              int(want_leaves), C.byref(pc), C.byref(ps), C.byref(n), per_cell.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"hmesh_chunk_leaves failed ({rc})")
    if not want_leaves:
        return None, None, per_cell
    codes = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_uint64)), (n.value,)).copy()
    sizes = np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_int32)), (n.value,)).astype(np.int64)
    L.hmesh_free(pc); L.hmesh_free(ps)
    return codes, sizes, per_cell


def _take(ptr, ctype, shape, L):
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape).copy()
    L.hmesh_free(ptr)
    return a


def extract_native(grid: CoarseGrid, X, codes, sizes, lstart, chunk: int = 4096, threads: int = 1):
    """octree.extract on the leaves of the cells X (ascending codes; lstart[X.size + 1] = first leaf of every
    cell) by the native primitives, chunk by chunk.  Same return values plus the hanging flags and the leaf
    that holds every node, with ONE difference: corners located in cells outside X (on the outer faces of X)
    are not numbered; they all map to a single extra node at the end (index N, `trash`) that no rank-relevant
    quantity reads.  A hanging node with an anchor out there keeps its flag but gets no dnode row.
    Returns ((ex, ey, ez, es), (px, py, pz), lnid, dnode, dang, holder, trash)."""
    L = mesh_lib()
    dims = np.array(grid.dims, np.int32)
    xkeys = np.ascontiguousarray(grid.key[X], np.uint64)
    lstart = np.ascontiguousarray(lstart, np.int64)
    codes = np.ascontiguousarray(codes, np.uint64)
    sizes32 = np.ascontiguousarray(sizes, np.int32)
    nX, E = X.size, codes.size

    def nodes_of(r):
        i0, i1 = r
        pn, px_, ph, pd, pa = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n, nd = C.c_int64(), C.c_int64()
        per_cell = np.zeros(i1 - i0, np.int64)
        rc = L.hmesh_chunk_nodes(dims.ctypes.data, grid.S, xkeys.ctypes.data, nX, lstart.ctypes.data, codes.ctypes.data,
                                 sizes32.ctypes.data, i0, i1, C.byref(pn), C.byref(px_), C.byref(ph), C.byref(pd), C.byref(pa),
                                 C.byref(n), C.byref(nd), per_cell.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"hmesh_chunk_nodes failed ({rc})")
        m = n.value
        return (_take(pn, C.c_uint64, (m,), L), _take(px_, C.c_int32, (m, 3), L), _take(ph, C.c_int64, (m,), L),
                _take(pd, C.c_uint8, (m,), L), _take(pa, C.c_uint64, (nd.value, 4), L), per_cell)
    ranges = [(i, min(i + chunk, nX)) for i in range(0, nX, chunk)]
    with ThreadPoolExecutor(max(threads, 1)) as ex:
        res = list(ex.map(nodes_of, ranges))
    ncodes = np.concatenate([r[0] for r in res])
    xyz = np.concatenate([r[1] for r in res])
    holder = np.concatenate([r[2] for r in res])
    dangf = np.concatenate([r[3] for r in res])
    acode = np.concatenate([r[4] for r in res])
    nstart = np.concatenate([[0], np.cumsum(np.concatenate([r[5] for r in res]))]).astype(np.int64)
    del res
    N = ncodes.size
    assert N < 2 ** 31 - 1
    lnid = np.empty((E, 8), np.int32)
    exyz = np.empty((E, 3), np.int32)
    step = max(1, -(-E // max(4 * threads, 1)))

    def lnid_of(e0):
        e1 = min(e0 + step, E)
        rc = L.hmesh_lnid(dims.ctypes.data, grid.S, xkeys.ctypes.data, nX, nstart.ctypes.data, ncodes.ctypes.data,
                          codes.ctypes.data, sizes32.ctypes.data, e0, e1, N, lnid[e0:e1].ctypes.data, exyz[e0:e1].ctypes.data)
        if rc != 0:
            raise RuntimeError(f"hmesh_lnid failed ({rc})")
    with ThreadPoolExecutor(max(threads, 1)) as ex:
        list(ex.map(lnid_of, range(0, E, step)))
    # hanging nodes: anchors from codes to ids; rows with an anchor outside X are dropped (flag kept)
    didx = np.nonzero(dangf)[0]
    deps = dangf[didx].astype(np.int64)
    used = np.arange(4)[None, :] < deps[:, None]
    ids = np.minimum(np.searchsorted(ncodes, acode), N - 1)
    found = ncodes[ids] == acode
    okrow = (found | ~used).all(1)
    dnode = np.full((int(okrow.sum()), 6), -1, np.int32)
    dnode[:, 0] = didx[okrow]
    dnode[:, 1] = deps[okrow]
    dnode[:, 2:6] = np.where(used[okrow], ids[okrow], -1)
    del ncodes
    px, py, pz = (np.append(xyz[:, c], 0).astype(np.int64) for c in range(3))      # + the trash node
    dang = np.append(dangf != 0, False)
    holder = np.append(holder, -1)
    ex_, ey_, ez_ = (exyz[:, c].astype(np.int64) for c in range(3))
    return (ex_, ey_, ez_, np.asarray(sizes, np.int64)), (px, py, pz), lnid, dnode, dang, holder, N


def _accumulate_native(nT, lnid, pr, dt, threads: int = 1):
    """meshgen._accumulate (grouped form, exact = False) with the two big per-node sums -- sum(M) and sum(dt a M)
    over the incident (element, corner) pairs -- done by hmesh_corner_sums in the same summation order; the
    dashpot terms (boundary elements only) stay in numpy.  Same doubles as the numpy restatement."""
    L = mesh_lib()
    E, N = lnid.shape[0], nT.shape[0]
    lnid = np.ascontiguousarray(lnid, np.int32)
    M = np.ascontiguousarray(pr["M"], np.float64)
    daM = np.ascontiguousarray(dt * pr["a"] * pr["M"], np.float64)
    sumM, sumaM = np.zeros(N), np.zeros(N)

    def one(args):
        w, out = args
        wp, op = (C.c_void_p * 1)(w.ctypes.data), (C.c_void_p * 1)(out.ctypes.data)
        rc = L.hmesh_corner_sums(E, lnid.ctypes.data, N, 1, wp, op)
        if rc != 0:
            raise RuntimeError(f"hmesh_corner_sums failed ({rc})")
    with ThreadPoolExecutor(2 if threads > 1 else 1) as ex:
        list(ex.map(one, [(M, sumM), (daM, sumaM)]))
    nT[:, 0] += sumM
    base1, base2 = sumM - sumaM, 2 * sumM - sumaM
    bi, dash = pr["bidx"], pr["dash"]
    bflat = lnid[bi].reshape(-1)
    for ax in range(3):
        d = dt * dash[:, :, ax].reshape(-1)
        dsum = np.bincount(bflat, d, N) if bflat.size else 0.0
        nT[:, 4 + ax] += base1 - dsum
        nT[:, 1 + ax] += base2 - dsum


def _chunks(ids, chunk):
    return [ids[i:i + chunk] for i in range(0, ids.size, chunk)]


def leaf_counts(grid: CoarseGrid, ids, vs_of, factor_h, chunk: int = 4096, threads: int = 1, model: GridModel | None = None):
    """Leaves of the whole-domain mesh per coarse cell, for the cells `ids` (ascending).  model: run the native
    primitive on it (vs_of is then not used); otherwise the numpy restatement evaluates vs_of."""
    parts = _chunks(np.asarray(ids, np.int64), chunk)
    if model is not None:
        def one(sub):
            return _chunk_leaves_native(grid, sub, model, factor_h, False)[2]
    else:
        def one(sub):
            return _chunk_leaves(grid, sub, vs_of, factor_h)[2]
    with ThreadPoolExecutor(max(threads, 1)) as ex:
        res = list(ex.map(one, parts))
    return np.concatenate(res) if res else np.zeros(0, np.int64)


def exact_leaves(grid: CoarseGrid, ids, vs_of, factor_h, chunk: int = 4096, threads: int = 1, model: GridModel | None = None):
    """(codes ascending, sizes, leaves per cell) of the whole-domain mesh inside the cells `ids` (ascending)."""
    parts = _chunks(np.asarray(ids, np.int64), chunk)
    if model is not None:
        def one(sub):
            return _chunk_leaves_native(grid, sub, model, factor_h)
    else:
        def one(sub):
            return _chunk_leaves(grid, sub, vs_of, factor_h)
    with ThreadPoolExecutor(max(threads, 1)) as ex:
        res = list(ex.map(one, parts))
    return (np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res]),
            np.concatenate([r[2] for r in res]).astype(np.int64))


def _msglist(nd, peers):
    if nd.size == 0:
        return MsgList()
    order = list(dict.fromkeys(peers.tolist()))[::-1]            # messengers are pushed at the head (psolve.c:4733)
    maps = [nd[peers == p_] for p_ in order]
    return MsgList(np.array(order, np.int32), np.array([m.size for m in maps], np.int32),
                   np.concatenate(maps).astype(np.int32))


def partition_local(dims, leaves, gidx, etotal: int, nodes, lnid, dnode, rank: int, world: int, dang=None, trash=None,
                    holder=None, lcode=None, native=None):
    """octree.partition on the leaves of X (the rank's cells and one ring) instead of the whole mesh:
    gidx = global Morton index of every leaf of X (ascending), etotal = leaves of the whole mesh.  Same rules,
    same return values; ranks other than `rank` are only described where they meet nodes `rank` owns."""
    nx, ny, nz = dims
    ex, ey, ez, es = leaves
    px, py, pz = nodes
    N = px.size
    if lcode is None:
        lcode = oc._code(ex, ey, ez)
    if holder is None:
        qx, qy, qz = np.minimum(px, nx - 1), np.minimum(py, ny - 1), np.minimum(pz, nz - 1)
        hl = np.searchsorted(lcode, oc._code(qx, qy, qz), side="right") - 1
        inside = (ex[hl] <= qx) & (qx < ex[hl] + es[hl]) & (ey[hl] <= qy) & (qy < ey[hl] + es[hl]) & \
                 (ez[hl] <= qz) & (qz < ez[hl] + es[hl])
    else:                                                        # extract_native looked the holders up cell by cell
        hl, inside = np.maximum(holder, 0), holder >= 0
    # a node on the outer faces of X may be held by a leaf that is not in X: nobody's here (it cannot be
    # this rank's: the rank's leaves and every leaf touching them are in X)
    owner = np.where(inside, mg.block_owner(gidx[hl], world, etotal), -1).astype(np.int32)
    if trash is not None:
        owner[trash] = -1                                        # extract_native's stand-in for corners outside X
    eown = mg.block_owner(gidx, world, etotal)
    lo, hi = mg.block_low(rank, world, etotal), mg.block_low(rank + 1, world, etotal)
    a, b = int(np.searchsorted(gidx, lo)), int(np.searchsorted(gidx, hi))
    assert b - a == hi - lo, "X does not hold the whole block of the rank"
    if dang is None:
        dang = np.zeros(N, bool)
        dang[dnode[:, 0]] = True
    ranks = [int(r) for r in np.unique(eown)]
    direct, harbor = {}, {}
    downer = owner[dnode[:, 0]]
    for r in ranks:
        d = np.zeros(N, bool)
        d[(lnid[a:b] if r == rank else lnid[np.nonzero(eown == r)[0]]).reshape(-1)] = True
        hb = d | (owner == r)
        anc = dnode[downer == r][:, 2:6]
        hb[anc[anc >= 0]] = True
        direct[r], harbor[r] = d, hb
    H = np.nonzero(harbor[rank])[0]                              # ascending = ascending Z-order
    local = np.full(N, -1, np.int64)
    local[H] = np.arange(H.size)
    l_lnid = local[lnid[a:b]].astype(np.int32)
    rows = dnode[downer == rank]
    l_dnode = rows.copy()
    l_dnode[:, 0] = local[rows[:, 0]]
    for k in range(4):
        has = rows[:, 2 + k] >= 0
        l_dnode[has, 2 + k] = local[rows[has, 2 + k]]
    assert l_lnid.min() >= 0 and (l_dnode[:, 0] >= 0).all()
    assert (owner[H] >= 0).all(), "a harbored node has no holder inside X"
    mine = owner[H] == rank
    anch = ~dang[H]
    msg = {}
    ln = np.arange(H.size)
    nm = ln[~mine]
    an, dn = nm[anch[nm]], nm[~anch[nm]]
    msg["an_c"] = _msglist(an, owner[H[an]].astype(np.int64))
    msg["dn_c"] = _msglist(dn, owner[H[dn]].astype(np.int64))
    disc = _discovery_order_local(dims, leaves, lcode, gidx, etotal, a, b, rank, world, harbor, lnid, native)
    pos = {p_: i for i, p_ in enumerate(disc)}
    sh_n, sh_p, sh_k = [], [], []
    own_l = ln[mine]
    for s in ranks:
        if s == rank:
            continue
        hit = own_l[harbor[s][H[own_l]]]
        if not hit.size:
            continue
        ind = ~direct[s][H[hit]]                                 # harbored by s only as an anchor: listed first,
        key = np.where(ind, -1 - s, pos.get(s, world))           # highest rank first; then discovery order
        sh_n.append(hit); sh_p.append(np.full(hit.size, s, np.int64)); sh_k.append(key.astype(np.int64))
    if sh_n:
        sh_n, sh_p, sh_k = np.concatenate(sh_n), np.concatenate(sh_p), np.concatenate(sh_k)
        o = np.lexsort((sh_k, sh_n))
        sh_n, sh_p = sh_n[o], sh_p[o]
        share = np.stack([sh_n, sh_p], 1).astype(np.int32)
        msg["an_s"] = _msglist(sh_n[anch[sh_n]], sh_p[anch[sh_n]])
        msg["dn_s"] = _msglist(sh_n[~anch[sh_n]], sh_p[~anch[sh_n]])
    else:
        share = np.zeros((0, 2), np.int32)
        msg["an_s"], msg["dn_s"] = MsgList(), MsgList()
    return (a, b), H, l_lnid, l_dnode, owner[H], share, anch, msg


def _discovery_order_local(dims, leaves, lcode, gidx, etotal, a, b, rank, world, harbor, lnid, native=None):
    """octree._discovery_order with the rank of a probe point taken from the global index of the X leaf
    that holds it (every probe lies within half an edge of one of the rank's leaves: inside X).
    native = (S, xkeys, lstart): the probes run in hmesh_discovery."""
    nx, ny, nz = dims
    ex, ey, ez, es = leaves
    shared_node = np.sum([h for h in harbor.values()], axis=0) > 1
    cand = a + np.nonzero(shared_node[lnid[a:b]].any(1))[0]
    if native is not None:
        S, xkeys, lstart = native
        L = mesh_lib()
        d32 = np.array(dims, np.int32)
        xkeys = np.ascontiguousarray(xkeys, np.uint64)
        lstart = np.ascontiguousarray(lstart, np.int64)
        codes = np.ascontiguousarray(lcode, np.uint64)
        sizes32 = np.ascontiguousarray(es, np.int32)
        g64 = np.ascontiguousarray(gidx, np.int64)
        cand64 = np.ascontiguousarray(cand, np.int64)
        base = np.ascontiguousarray(cand - a, np.int64)
        first = np.full(world, np.iinfo(np.int64).max, np.int64)
        rc = L.hmesh_discovery(d32.ctypes.data, S, xkeys.ctypes.data, xkeys.size, lstart.ctypes.data, codes.ctypes.data,
                               sizes32.ctypes.data, g64.ctypes.data, etotal, world, rank, cand64.ctypes.data, base.ctypes.data,
                               cand64.size, first.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"hmesh_discovery failed ({rc})")
        met = np.nonzero(first != np.iinfo(np.int64).max)[0]
        return [int(p_) for p_ in met[np.argsort(first[met], kind="stable")]]
    x, y, z, s = ex[cand], ey[cand], ez[cand], es[cand]
    first = {}
    for k in range(4):
        for j in range(4):
            for i in range(4):
                p2 = (2 * x - s + s * i, 2 * y - s + s * j, 2 * z - s + s * k)        # doubled coordinates
                ok = (p2[0] >= 0) & (p2[0] < 2 * nx) & (p2[1] >= 0) & (p2[1] < 2 * ny) & (p2[2] >= 0) & (p2[2] < 2 * nz)
                if not ok.any():
                    continue
                q = tuple(np.where(ok, c // 2, 0) for c in p2)
                h_ = np.searchsorted(lcode, oc._code(*q), side="right") - 1
                r = mg.block_owner(gidx[h_], world, etotal)
                keyv = (cand - a).astype(np.int64) * 64 + (k * 16 + j * 4 + i)
                for p_ in np.unique(r[ok]):
                    if p_ == rank:
                        continue
                    m = int(keyv[ok & (r == p_)].min())
                    if int(p_) not in first or m < first[int(p_)]:
                        first[int(p_)] = m
    return [p_ for p_, _ in sorted(first.items(), key=lambda kv: kv[1])]


def octree_halfspace_local(dims, smax: int, h: float, dt: float, materials, mat_of, ppw: float, fmax: float,
                           rank: int, world: int, root: int | None = None, freq: float | None = None,
                           damping: int = RAYLEIGH, thr_damping: float = 0.05, thr_vpvs: float = 3.0,
                           vs_min: float = 0.0, exact: bool = False, counts=None, allgather=None,
                           chunk: int = 4096, threads: int = 1, model_cell: int | None = None, model: GridModel | None = None):
    """One rank's mesh and solver tables, as octree.octree_halfspace_part returns them, from local work.

    counts    leaves per coarse cell of the whole domain (Morton order), if the caller already has them;
    allgather callable(np.ndarray) -> list of the arrays of all ranks in rank order (torch.distributed /
              MPI): with it every rank counts 1/world of the cells.  With neither, this rank counts them all.
    model_cell edge (in h) of the cells the material model is constant on: with it (or a ready GridModel in
              `model`) refinement and balancing run in the native primitives (csrc/hmesh.cpp) on a rasterised
              model; without it the numpy restatement evaluates mat_of directly.
    Returns (HostMesh of the rank, info); info["counts"] / info["model"] can be handed to the other ranks of
    one process."""
    if root is None:
        root = 1
        while root < max(dims):
            root *= 2
    S = min(smax, oc.bootstrap_size(dims, root, world))
    grid = CoarseGrid(dims, S)
    vs_tab = np.array([max(m[1], vs_min) for m in materials], np.float64)
    factor_h = h * ppw * fmax

    def vs_of(x, y, z):
        return vs_tab[mat_of(x, y, z)]
    if model is None and model_cell is not None:
        model = GridModel(dims, model_cell, mat_of, vs_tab)
    if counts is None:
        if allgather is None:
            counts = leaf_counts(grid, np.arange(grid.n), vs_of, factor_h, chunk, threads, model)
        else:
            c0, c1 = rank * grid.n // world, (rank + 1) * grid.n // world
            counts = np.concatenate(allgather(leaf_counts(grid, np.arange(c0, c1), vs_of, factor_h, chunk, threads, model)))
    counts = np.asarray(counts, np.int64)
    assert counts.size == grid.n and counts.min() >= 1
    prefix = np.concatenate([[0], np.cumsum(counts)])
    etotal = int(prefix[-1])
    lo, hi = mg.block_low(rank, world, etotal), mg.block_low(rank + 1, world, etotal)
    c0 = int(np.searchsorted(prefix, lo, side="right")) - 1
    c1 = int(np.searchsorted(prefix, hi, side="left"))
    X = grid.ring(np.arange(c0, max(c1, c0 + 1)))
    codes, sizes, per_cell = exact_leaves(grid, X, vs_of, factor_h, chunk, threads, model)
    assert np.array_equal(per_cell, counts[X]), "leaf counts of the two passes disagree"
    start = np.concatenate([[0], np.cumsum(per_cell)])[:-1]
    # global Morton index of every leaf of X: first leaf of its cell in the whole mesh + position inside the cell
    gidx = np.repeat(prefix[X] - start, per_cell) + np.arange(codes.size)
    dang = trash = holder = native = None
    if model is not None:
        lstart = np.concatenate([start, [codes.size]])
        (ex, ey, ez, es), (px, py, pz), lnid, dnode, dang, holder, trash = extract_native(
            grid, X, codes, sizes, lstart, chunk, threads)
        native = (S, grid.key[X], lstart)
    else:
        leaves = {int(s): oc._decode(codes[sizes == s]) for s in np.unique(sizes)}
        (ex, ey, ez, es), (px, py, pz), lnid, dnode = oc.extract(leaves, dims)
    del sizes
    # solver_init's tables on X (rows of the nodes this rank owns are complete: every element that touches
    # them, and every dangling node anchored at them with all ITS elements, is in X)
    abase, bbase = mg.compute_setab(damping, fmax if freq is None else freq)
    layers = [(0.0, vp, vs, rho) for (vp, vs, rho) in materials]
    if model is not None:                                        # the model cell that holds the element's centre
        d = 2 * model.cl
        mat = model.grid[(2 * ex + es) // d, (2 * ey + es) // d, (2 * ez + es) // d].astype(np.int64)
    else:
        mat = mat_of(ex + 0.5 * es, ey + 0.5 * es, ez + 0.5 * es).astype(np.int64)
    pr = mg._elem_props(ex, ey, ez, dims, h, dt, layers, abase, bbase, thr_damping, thr_vpvs, size=es, mat=mat)
    nT = np.zeros((px.size, 7))
    if model is not None and not exact:
        _accumulate_native(nT, lnid, pr, dt, threads)
    else:
        mg._accumulate(nT, lnid, pr, dt, exact)
    mg._distribute(nT, dnode)
    (a, b), H, l_lnid, l_dnode, owner, share, anch, msg = partition_local(dims, (ex, ey, ez, es), gidx, etotal,
                                                                          (px, py, pz), lnid, dnode, rank, world,
                                                                          dang, trash, holder, codes, native)
    edata = np.zeros((b - a, 14), np.float32)
    edata[:, 0], edata[:, 1], edata[:, 2], edata[:, 3] = pr["edge"][a:b], pr["Vp"][a:b], pr["Vs"][a:b], pr["rho"][a:b]
    if damping == BKT:
        mab = mat[a:b]                                           # one table search per material, not per element
        mu = np.flatnonzero(np.bincount(mab))
        first = np.array([int(np.argmax(mab == m)) for m in mu], np.int64)
        lut = np.zeros(int(mu.max()) + 1 if mu.size else 1, np.int64)
        lut[mu] = np.arange(mu.size)
        edata[:, 4:14] = mg.bkt_coefficients(pr["Vp"][a:b][first], pr["Vs"][a:b][first])[lut[mab]]
    K1, K2 = mg.compute_K()
    part = HostMesh(l_lnid, pr["eT"][a:b], nT[H], l_dnode, edata, K1, K2, msg["dn_c"], msg["dn_s"], msg["an_c"], msg["an_s"])
    info = dict(E=b - a, N=H.size, D=int(l_dnode.shape[0]), node_xyz=(px[H], py[H], pz[H]),
                elem_xyz=(ex[a:b], ey[a:b], ez[a:b]), elem_size=es[a:b], elem_geid=gidx[a:b].astype(np.int64), owner=owner,
                share=share, anchored=anch, dims=dims, h=h, rank=rank, nranks=world, etotal=etotal, origin=(0, 0, 0),
                abase=abase, bbase=bbase, counts=counts, model=model, local_region_elements=int(ex.size), coarse_edge=S)
    return part, info
