"""General single-rank octree meshes in the reference's layout.

meshgen.py covers the two families the benchmarks need (uniform, depth-banded).  This module takes
ANY material model that is piecewise constant on a grid and does what the reference's mesher does
with it, vectorised in numpy:

  refine     octor_refinetree + toexpand / vsrule (octor.c:4337, psolve.c:2185, quake_util.c:215):
             an octant is split while its edge exceeds Vs(centre) / (points-per-wavelength * f_max)
  balance    octor_balancetree (octor.c:4398-4700): 2:1 across faces AND edges (18 directions, not
             corners), by prioritised ripple propagation from the finest level up -- the result is
             the unique smallest balanced refinement, so any algorithm that reaches it agrees
  extract    octor_extractmesh (octor.c:5268-6645): leaves in Morton order of their lowest corner,
             nodes = distinct corners in Z-order with the far domain faces pulled in by one tick,
             elem_t.lnid, and the hanging nodes with their anchors (node_setproperty octor.c:3294,
             anchor lists octor.c:5863-5991)

and then solver_init's tables through meshgen's helpers.  tests/test_octree.py pins all three stages
bit for bit on meshes the unmodified reference produced (banded and laterally varying models).
Coordinates are integers in units of the finest admissible edge h.
"""
from __future__ import annotations

import numpy as np

from . import meshgen as mg
from .solver import HostMesh, MsgList, RAYLEIGH, BKT

_DIRS18 = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)
           if 1 <= abs(dx) + abs(dy) + abs(dz) <= 2]


def _code(x, y, z):
    """Morton code, x the least significant interleaved bit (coordinates below 2^16)."""
    return mg.morton3_fast(x, y, z)


def _decode(codes):
    x, y, z = mg.demorton3_fast(codes, 4)
    return x.astype(np.int64), y.astype(np.int64), z.astype(np.int64)


def refine(dims, smax: int, vs_of, factor_h: float, smin: int = 1):
    """Top-down refinement.  dims = (nx, ny, nz) in units of h (multiples of smax, the largest octant
    that fits the domain's alignment); vs_of(xc, yc, zc) -> Vs at points given in units of h;
    factor_h = h * ppw * f_max: an octant of edge s is split when s * factor_h > Vs (vsrule).
    Returns {size: (x, y, z)} of leaves."""
    nx, ny, nz = dims
    gx, gy, gz = np.meshgrid(np.arange(0, nx, smax), np.arange(0, ny, smax), np.arange(0, nz, smax), indexing="ij")
    cur = (gx.ravel().astype(np.int64), gy.ravel().astype(np.int64), gz.ravel().astype(np.int64))
    return refine_cells(cur, smax, vs_of, factor_h, smin)


def refine_cells(cur, smax: int, vs_of, factor_h: float, smin: int = 1):
    """refine() from a given set of octants of edge smax: cur = (x, y, z) of their lowest corners."""
    leaves, s = {}, smax
    while cur[0].size:
        x, y, z = cur
        split = (s * factor_h > vs_of(x + 0.5 * s, y + 0.5 * s, z + 0.5 * s)) & (s > smin)
        leaves[s] = (x[~split], y[~split], z[~split])
        x, y, z = x[split], y[split], z[split]
        hs = s // 2
        cur = tuple(np.concatenate([c + hs * ((j >> k) & 1) for j in range(8)]) for k, c in enumerate((x, y, z)))
        s = hs
        if s == 0:
            break
    return {k: v for k, v in leaves.items() if v[0].size}


def balance(leaves: dict, dims, as_codes: bool = False):
    """2:1 balance across faces and edges, finest level first (prioritised ripple propagation).
    The leaf set may cover only part of the domain: octants that are not there constrain nothing.
    as_codes: return {size: sorted Morton codes} instead of coordinates."""
    nx, ny, nz = dims
    sets = {s: np.unique(_code(*v)) for s, v in leaves.items()}
    sizes = sorted(sets)
    smax = max(sizes)
    s = min(sizes)
    while s * 2 < smax:
        if s in sets and sets[s].size:
            x, y, z = _decode(sets[s])
            need = []
            for dx, dy, dz in _DIRS18:
                qx, qy, qz = x + dx * s, y + dy * s, z + dz * s
                ok = (qx >= 0) & (qx < nx) & (qy >= 0) & (qy < ny) & (qz >= 0) & (qz < nz)
                t = 2 * s
                need.append(_code((qx[ok] // t) * t, (qy[ok] // t) * t, (qz[ok] // t) * t))
            need = np.unique(np.concatenate(need))                 # cells of size 2 s that must not lie in a larger leaf
            t = 4 * s
            while t <= smax:
                if t in sets and sets[t].size:
                    nxq, nyq, nzq = _decode(need)
                    holder = _code((nxq // t) * t, (nyq // t) * t, (nzq // t) * t)
                    hit = np.isin(holder, sets[t])
                    if hit.any():
                        # split every such leaf down to size 2 s along the way to each needed cell
                        split = np.unique(holder[hit])
                        sets[t] = np.setdiff1d(sets[t], split, assume_unique=True)
                        targets = need[hit]
                        u = t
                        parents = split
                        while u > 2 * s:
                            hu = u // 2
                            px, py, pz = _decode(parents)
                            kids = np.concatenate([_code(px + hu * (j & 1), py + hu * ((j >> 1) & 1), pz + hu * ((j >> 2) & 1))
                                                   for j in range(8)])
                            tx, ty, tz = _decode(targets)
                            on_path = np.unique(_code((tx // hu) * hu, (ty // hu) * hu, (tz // hu) * hu))
                            if hu > 2 * s:
                                stay = np.setdiff1d(kids, on_path)
                                parents = np.intersect1d(kids, on_path)
                            else:
                                stay, parents = kids, np.zeros(0, np.uint64)
                            sets[hu] = np.union1d(sets.get(hu, np.zeros(0, np.uint64)), stay)
                            u = hu
                t *= 2
        s *= 2
    if as_codes:
        return {s: c for s, c in sets.items() if c.size}
    return {s: _decode(c) for s, c in sets.items() if c.size}


def extract(leaves: dict, dims):
    """Leaf set -> (ex, ey, ez, es) in Morton order, nodes (px, py, pz) in octor's order, lnid [E][8],
    dnode [D][6]."""
    nx, ny, nz = dims
    ex = np.concatenate([v[0] for v in leaves.values()]).astype(np.int64)
    ey = np.concatenate([v[1] for v in leaves.values()]).astype(np.int64)
    ez = np.concatenate([v[2] for v in leaves.values()]).astype(np.int64)
    es = np.concatenate([np.full(v[0].size, s, np.int64) for s, v in leaves.items()])
    o = np.argsort(_code(ex, ey, ez), kind="stable")
    ex, ey, ez, es = ex[o], ey[o], ez[o], es[o]
    E = ex.size

    def key(g, n):
        return np.where(g == n, 2 * n - 1, 2 * g)

    def ncode(x, y, z):
        return _code(key(x, nx), key(y, ny), key(z, nz))
    corner = [(ex + es * (j & 1), ey + es * ((j >> 1) & 1), ez + es * ((j >> 2) & 1)) for j in range(8)]
    ccode = [ncode(*c) for c in corner]
    codes, first = np.unique(np.concatenate(ccode), return_index=True)     # ascending = octor's node order
    N = codes.size
    lnid = np.stack([np.searchsorted(codes, c) for c in ccode], 1).astype(np.int32)
    del ccode
    px = np.concatenate([c[0] for c in corner])[first]
    py = np.concatenate([c[1] for c in corner])[first]
    pz = np.concatenate([c[2] for c in corner])[first]
    # smallest leaf that has the node as a corner
    smin = np.full(N, np.iinfo(np.int64).max, np.int64)
    for j in range(8):
        np.minimum.at(smin, lnid[:, j], es)
    # ---- hanging nodes: a node is dangling when a leaf twice the size of its smallest leaf holds it
    #      inside an edge (one coordinate off the 2 s grid) or inside a face (two coordinates off) ----
    lcode = _code(ex, ey, ez)                                    # ascending (Morton order)

    def is_leaf(x, y, z, s):
        inb = (x >= 0) & (x + s <= nx) & (y >= 0) & (y + s <= ny) & (z >= 0) & (z + s <= nz)
        c = _code(np.where(inb, x, 0), np.where(inb, y, 0), np.where(inb, z, 0))
        i = np.minimum(np.searchsorted(lcode, c), E - 1)
        return inb & (lcode[i] == c) & (es[i] == s)
    P = np.stack([px, py, pz], 1)
    t = 2 * smin
    odd = (P % t[:, None]) != 0                                   # [N][3]
    k = odd.sum(1)
    # a node that is a corner of 8 leaves is interior and anchored (node_setproperty, octor.c:3316): only
    # the others -- interfaces between levels and the domain boundary -- need the look-ups
    touches = np.bincount(lnid.reshape(-1), minlength=N)
    ci = np.nonzero((touches < 8) & (k >= 1) & (k <= 2))[0]
    Pc, sc, tc, oc = P[ci], smin[ci], t[ci], odd[ci]
    hit = np.zeros(ci.size, bool)
    for m in range(8):
        # the candidate leaves of size 2 s around the node: an off-grid coordinate fixes the candidate's
        # start (s below the node); an on-grid one leaves two choices, the cell below or the cell at the node
        # (combinations that only differ in an off-grid axis' bit repeat a candidate: harmless)
        cand = np.empty((ci.size, 3), np.int64)
        for c in range(3):
            cand[:, c] = np.where(oc[:, c], Pc[:, c] - sc, Pc[:, c] - tc if (m >> c) & 1 else Pc[:, c])
        hit |= is_leaf(cand[:, 0], cand[:, 1], cand[:, 2], tc)
    dang = np.zeros(N, bool)
    dang[ci[hit]] = True
    didx = np.nonzero(dang)[0]
    dnode = np.full((didx.size, 6), -1, np.int32)
    dnode[:, 0] = didx
    dnode[:, 1] = np.where(k[didx] == 1, 2, 4)
    # anchors in descending Z-order: +-s along the off-grid axes, the higher axis (z > y > x) varying slowest
    for a in range(4):
        q = P[didx].copy()
        sm = smin[didx]
        o_ = odd[didx]
        # rank the off-grid axes: first (lowest) and second
        ax1 = np.argmax(o_, axis=1)                               # lowest off-grid axis
        ax2 = 2 - np.argmax(o_[:, ::-1], axis=1)                  # highest off-grid axis (== ax1 when k == 1)
        two = k[didx] == 2
        use = (a < 2) | two
        s1 = np.where(a % 2 == 0, 1, -1)                          # lowest axis: +, -, +, -
        s2 = np.where(a < 2, 1, -1)                               # highest axis: +, +, -, -
        rows = np.arange(didx.size)
        q[rows, ax1] += s1 * sm
        q[rows[two], ax2[two]] += s2 * sm[two]
        ids = np.searchsorted(codes, ncode(q[:, 0], q[:, 1], q[:, 2]))
        ok = use & (ids < N)
        ids = np.where(ok, ids, 0)
        assert (codes[ids[use]] == ncode(q[use, 0], q[use, 1], q[use, 2])).all(), "an anchor is not a mesh node"
        dnode[use, 2 + a] = ids[use]
    return (ex, ey, ez, es), (px, py, pz), lnid, dnode


def octree_halfspace(dims, smax: int, h: float, dt: float, materials, mat_of, ppw: float, fmax: float,
                     freq: float | None = None, damping: int = RAYLEIGH, thr_damping: float = 0.05,
                     thr_vpvs: float = 3.0, vs_min: float = 0.0, exact: bool = False):
    """Mesh + solver tables for a model that is piecewise constant: materials = [(Vp, Vs, rho)],
    mat_of(xc, yc, zc) -> material index at points in units of h.  Returns (HostMesh, info)."""
    vs_tab = np.array([max(m[1], vs_min) for m in materials], np.float64)
    leaves = balance(refine(dims, smax, lambda x, y, z: vs_tab[mat_of(x, y, z)], h * ppw * fmax), dims)
    (ex, ey, ez, es), (px, py, pz), lnid, dnode = extract(leaves, dims)
    abase, bbase = mg.compute_setab(damping, fmax if freq is None else freq)
    layers = [(0.0, vp, vs, rho) for (vp, vs, rho) in materials]
    mat = mat_of(ex + 0.5 * es, ey + 0.5 * es, ez + 0.5 * es).astype(np.int64)
    pr = mg._elem_props(ex, ey, ez, dims, h, dt, layers, abase, bbase, thr_damping, thr_vpvs, size=es, mat=mat)
    nT = np.zeros((px.size, 7))
    mg._accumulate(nT, lnid, pr, dt, exact)
    mg._distribute(nT, dnode)
    edata = np.zeros((ex.size, 14), np.float32)
    edata[:, 0], edata[:, 1], edata[:, 2], edata[:, 3] = pr["edge"], pr["Vp"], pr["Vs"], pr["rho"]
    if damping == BKT:                                            # one table search per material, not per element
        mu, first, inv = np.unique(mat, return_index=True, return_inverse=True)
        edata[:, 4:14] = mg.bkt_coefficients(pr["Vp"][first], pr["Vs"][first])[inv]
    K1, K2 = mg.compute_K()
    mesh = HostMesh(lnid, pr["eT"], nT, dnode, edata, K1, K2, MsgList(), MsgList(), MsgList(), MsgList())
    info = dict(E=ex.size, N=px.size, D=int(dnode.shape[0]), node_xyz=(px, py, pz), elem_xyz=(ex, ey, ez),
                elem_size=es, dims=dims, h=h, origin=(0, 0, 0), abase=abase, bbase=bbase)
    return mesh, info


# ---- partition ----------------------------------------------------------------------------------------

def bootstrap_size(dims, root: int, world: int) -> int:
    """octor_newtree (octor.c:4170-4200) pushes the tree down until there are at least 10 tasks per rank,
    and octor_refinetree splits every one of those data-less task octants once more (toexpand returns 1
    for a leaf without data, psolve.c:2187): with `world` > 1 ranks no leaf is larger than half a task
    octant.  root = edge of the root octant in units of h.  Returns that largest admissible leaf edge."""
    if world == 1:
        return root
    s = root
    while True:
        tasks = 1
        for n in dims:
            tasks *= -(-n // s)
        if tasks >= 10 * world or s == 1:
            return max(s // 2, 1)
        s //= 2


def partition(dims, leaves, nodes, lnid, dnode, nT, rank: int, world: int):
    """One rank's share of a mesh extracted on the whole domain, as octor_partitiontree +
    octor_extractmesh + schedule_build make it:

    * elements: the rank's block of the Morton-ordered leaf list (octor.c:685-746);
    * a node belongs to the rank whose block holds the leaf that contains it (far faces pulled in,
      octor.c:5466-5475); a rank harbors the corners of its elements, every node it owns (neighbours
      report them, octor.c:5700-5760) and the anchors of the dangling nodes it owns (octor.c:5863-5991);
    * share list of an owned node: the ranks that harbor it only as such an anchor, highest rank first,
      then the ranks whose elements touch it in the order this rank discovered them (com_allocpctl);
    * dnodeTable = the owned dangling nodes; schedules by schedule_build (psolve.c:4705-4863).
    nT = nTable of the whole mesh: rows are taken over (complete sums, see meshgen.graded_halfspace).
    Returns (elem range, harbored global node ids, local lnid, local dnode, owner, share, MsgLists)."""
    nx, ny, nz = dims
    ex, ey, ez, es = leaves
    px, py, pz = nodes
    E, N = ex.size, px.size
    lcode = _code(ex, ey, ez)
    holder = np.searchsorted(lcode, _code(np.minimum(px, nx - 1), np.minimum(py, ny - 1), np.minimum(pz, nz - 1)),
                             side="right") - 1
    owner = mg.block_owner(holder, world, E).astype(np.int32)
    lo = [mg.block_low(r, world, E) for r in range(world + 1)]
    dang = np.zeros(N, bool)
    dang[dnode[:, 0]] = True
    # what every rank harbors: corners of its elements (direct), owned nodes, anchors of owned dangling nodes
    direct = np.zeros((world, N), bool)
    harbor = np.zeros((world, N), bool)
    for r in range(world):
        direct[r, np.unique(lnid[lo[r]:lo[r + 1]])] = True
        harbor[r] = direct[r] | (owner == r)
        rows = dnode[owner[dnode[:, 0]] == r]
        anc = rows[:, 2:6]
        harbor[r, anc[anc >= 0]] = True
    H = np.nonzero(harbor[rank])[0]                              # ascending global id = ascending Z-order
    local = np.full(N, -1, np.int64)
    local[H] = np.arange(H.size)
    l_lnid = local[lnid[lo[rank]:lo[rank + 1]]].astype(np.int32)
    rows = dnode[owner[dnode[:, 0]] == rank]
    l_dnode = rows.copy()
    l_dnode[:, 0] = local[rows[:, 0]]
    for a in range(4):
        has = rows[:, 2 + a] >= 0
        l_dnode[has, 2 + a] = local[rows[has, 2 + a]]
    assert l_lnid.min() >= 0 and (l_dnode[:, 0] >= 0).all()
    mine = owner[H] == rank
    anch = ~dang[H]

    def make_list(nd, peers):
        if nd.size == 0:
            return MsgList()
        order = list(dict.fromkeys(peers.tolist()))[::-1]        # messengers are pushed at the head (psolve.c:4733)
        maps = [nd[peers == p_] for p_ in order]
        return MsgList(np.array(order, np.int32), np.array([m.size for m in maps], np.int32),
                       np.concatenate(maps).astype(np.int32))
    msg = {}
    ln = np.arange(H.size)
    nm = ln[~mine]
    an, dn = nm[anch[nm]], nm[~anch[nm]]
    msg["an_c"] = make_list(an, owner[H[an]].astype(np.int64))
    msg["dn_c"] = make_list(dn, owner[H[dn]].astype(np.int64))
    # share lists of the nodes I own
    disc = _discovery_order(dims, leaves, lcode, lo, rank, world, harbor, lnid)
    pos = {p_: i for i, p_ in enumerate(disc)}
    sh_n, sh_p, sh_k = [], [], []
    own_l = ln[mine]
    for s in range(world):
        if s == rank:
            continue
        hit = own_l[harbor[s, H[own_l]]]
        if not hit.size:
            continue
        ind = ~direct[s, H[hit]]                                 # harbored by s only as an anchor: listed first,
        key = np.where(ind, -1 - s, pos.get(s, world))           # highest rank first; then discovery order
        sh_n.append(hit); sh_p.append(np.full(hit.size, s, np.int64)); sh_k.append(key.astype(np.int64))
    if sh_n:
        sh_n, sh_p, sh_k = np.concatenate(sh_n), np.concatenate(sh_p), np.concatenate(sh_k)
        o = np.lexsort((sh_k, sh_n))
        sh_n, sh_p = sh_n[o], sh_p[o]
        share = np.stack([sh_n, sh_p], 1).astype(np.int32)
        msg["an_s"] = make_list(sh_n[anch[sh_n]], sh_p[anch[sh_n]])
        msg["dn_s"] = make_list(sh_n[~anch[sh_n]], sh_p[~anch[sh_n]])
    else:
        share = np.zeros((0, 2), np.int32)
        msg["an_s"], msg["dn_s"] = MsgList(), MsgList()
    return (lo[rank], lo[rank + 1]), H, l_lnid, l_dnode, owner[H], share, anch, msg


def _discovery_order(dims, leaves, lcode, lo, rank, world, harbor, lnid):
    """Neighbour ranks in the order com_allocpctl (octor.c:2639-2742) first meets them: local leaves in
    Morton order; per leaf 4 x 4 x 4 probe points half an edge apart, starting half an edge below the
    lowest corner (z outermost, x innermost); a point outside the domain is skipped; the rank of a point
    is the rank of the leaf that contains it."""
    nx, ny, nz = dims
    ex, ey, ez, es = leaves
    E = ex.size
    a, b = lo[rank], lo[rank + 1]
    # only leaves with a corner that somebody else harbors can see a foreign leaf
    shared_node = harbor.sum(0) > 1
    cand = a + np.nonzero(shared_node[lnid[a:b]].any(1))[0]
    x, y, z, s = ex[cand], ey[cand], ez[cand], es[cand]
    first = {}
    for k in range(4):
        for j in range(4):
            for i in range(4):
                # doubled coordinates: 2 x - s + s i
                p2 = (2 * x - s + s * i, 2 * y - s + s * j, 2 * z - s + s * k)
                ok = (p2[0] >= 0) & (p2[0] < 2 * nx) & (p2[1] >= 0) & (p2[1] < 2 * ny) & (p2[2] >= 0) & (p2[2] < 2 * nz)
                if not ok.any():
                    continue
                q = tuple(np.where(ok, c // 2, 0) for c in p2)
                h_ = np.searchsorted(lcode, _code(*q), side="right") - 1
                r = mg.block_owner(h_, world, E)
                keyv = (cand - a).astype(np.int64) * 64 + (k * 16 + j * 4 + i)
                for p_ in np.unique(r[ok]):
                    if p_ == rank:
                        continue
                    m = int(keyv[ok & (r == p_)].min())
                    if int(p_) not in first or m < first[int(p_)]:
                        first[int(p_)] = m
    return [p_ for p_, _ in sorted(first.items(), key=lambda kv: kv[1])]


def octree_halfspace_part(dims, smax, h, dt, materials, mat_of, ppw, fmax, rank: int, world: int, root: int | None = None,
                          **kw):
    """octree_halfspace on `world` ranks: the whole mesh is built (with the coarsest leaves octor's
    multi-rank bootstrap allows), then cut.  Returns (HostMesh of the rank, info)."""
    if root is None:
        root = 1
        while root < max(dims):
            root *= 2
    smax = min(smax, bootstrap_size(dims, root, world))
    mesh, info = octree_halfspace(dims, smax, h, dt, materials, mat_of, ppw, fmax, **kw)
    leaves = (*info["elem_xyz"], info["elem_size"])
    (a, b), H, l_lnid, l_dnode, owner, share, anch, msg = partition(dims, leaves, info["node_xyz"], mesh.elem_lnid,
                                                                    mesh.dnode, mesh.nTable, rank, world)
    part = HostMesh(l_lnid, mesh.eTable[a:b], mesh.nTable[H], l_dnode, mesh.edata[a:b], mesh.K1, mesh.K2,
                    msg["dn_c"], msg["dn_s"], msg["an_c"], msg["an_s"])
    pinfo = dict(E=b - a, N=H.size, D=int(l_dnode.shape[0]), node_xyz=tuple(c[H] for c in info["node_xyz"]),
                 elem_xyz=tuple(c[a:b] for c in info["elem_xyz"]), elem_size=info["elem_size"][a:b],
                 elem_geid=np.arange(a, b, dtype=np.int64), owner=owner, share=share, anchored=anch, dims=dims, h=h,
                 rank=rank, nranks=world, etotal=info["E"], origin=(0, 0, 0), gnid=H)
    return part, pinfo
