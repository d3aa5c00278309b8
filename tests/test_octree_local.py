"""CPU tests: hercules_b200.octree_local -- a rank's mesh from LOCAL work (its Morton block plus a one-cell
ring; leaf counts per coarse cell instead of the global leaf list) against (a) multi-rank runs of the
unmodified reference (the same goldens and the same assertions as tests/test_octree.py::
test_partition_reproduces_octor) and (b) the whole-mesh-then-cut build on the bench's basin model."""
import numpy as np
import pytest

from conftest import load_golden, params_of, rank_view
from test_octree import CASES, PART_CASES, _mat_of


@pytest.mark.parametrize("native", [False, True], ids=["numpy", "native"])
@pytest.mark.parametrize("name", sorted(PART_CASES))
def test_local_build_reproduces_octor(name, native):
    """native: refinement, balancing and extraction by csrc/hmesh.cpp on the rasterised model (cells of the
    CVM leaf edge); otherwise the numpy restatement on the model function."""
    from hercules_b200 import octree_local as ol
    base, world = PART_CASES[name]
    dims, h, smax, cl, mats, box, tops, vs_min, ppw, fmax = CASES[base]
    g = load_golden(name)
    counts = None
    for r in range(world):
        v = rank_view(g, r); P = params_of(v)
        mesh, info = ol.octree_halfspace_local(dims, smax, h, P["dt"], mats, _mat_of(CASES[base]), ppw, fmax, r, world,
                                               freq=P["freq"], damping=P["damping"], vs_min=vs_min, exact=True,
                                               counts=counts, chunk=5, threads=2, model_cell=cl if native else None)
        counts = info["counts"]
        assert np.array_equal(info["elem_geid"], v["elem_geid"])
        assert mesh.elem_lnid.shape == v["elem_lnid"].shape and np.array_equal(mesh.elem_lnid, v["elem_lnid"])
        xyz = np.stack(info["node_xyz"], 1)
        tick = int(v["node_ticks"][v["node_ticks"] > 0].min()) // int(xyz[xyz > 0].min())
        assert np.array_equal(xyz * tick, v["node_ticks"])
        ismine = v["node_flags"][:, 0].astype(bool)
        assert np.array_equal(info["owner"] == r, ismine)
        assert np.array_equal(info["owner"][~ismine], v["node_owner"][~ismine])
        assert np.array_equal(info["anchored"], v["node_flags"][:, 1].astype(bool))
        assert info["share"].shape == v["node_share"].shape and np.array_equal(info["share"], v["node_share"])
        assert mesh.dnode.shape == v["dnode"].shape and np.array_equal(mesh.dnode, v["dnode"])
        for side in ("dn_c", "dn_s", "an_c", "an_s"):
            ml = getattr(mesh, side)
            hdr = np.stack([ml.peer, ml.nodes], 1).reshape(-1, 2)
            assert np.array_equal(hdr, v[side + "_hdr"].reshape(-1, 2)), (r, side)
            assert np.array_equal(ml.mapping, v[side + "_map"]), (r, side)
        assert np.array_equal(mesh.eTable, v["eTable"])
        assert np.allclose(mesh.nTable[ismine], v["nTable"][ismine], rtol=2e-15, atol=0)


def _same_tables(a, ai, b, bi):
    assert np.array_equal(ai["elem_geid"], bi["elem_geid"]) and ai["etotal"] == bi["etotal"]
    for k in ("elem_lnid", "eTable", "dnode", "edata"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(np.stack(ai["node_xyz"]), np.stack(bi["node_xyz"]))
    assert np.array_equal(ai["owner"], bi["owner"]) and np.array_equal(ai["share"], bi["share"])
    assert np.array_equal(ai["anchored"], bi["anchored"])
    mine = ai["owner"] == ai["rank"]
    assert np.array_equal(a.nTable[mine], b.nTable[mine])       # complete sums, same summation order: bit-equal
    for side in ("dn_c", "dn_s", "an_c", "an_s"):
        x, y = getattr(a, side), getattr(b, side)
        assert np.array_equal(x.peer, y.peer) and np.array_equal(x.nodes, y.nodes) and np.array_equal(x.mapping, y.mapping)


@pytest.mark.parametrize("native", [False, True], ids=["numpy", "native"])
@pytest.mark.parametrize("world", [2, 5])
def test_local_build_equals_whole_mesh_cut_on_the_basin_model(world, native):
    """bench.py's configs[4] model at --edge 128 (57 k elements, 4 octree levels, BKT): every rank's tables
    from the local build equal the cut of the whole mesh; the counts come from the ranks' own shares of the
    coarse cells, gathered (here: in process)."""
    import bench
    import hercules_b200 as hb
    from hercules_b200 import octree, octree_local as ol
    n = 128
    whole = [bench.basin_workload(n, hb.BKT, (r, world))[:2] for r in range(world)]
    mat_of = whole[0][1]["mat_of"]
    dims = whole[0][1]["dims"]
    h, ppw = 300000.0 / n, 8.0
    fmax, dt = 499.0 / (ppw * h), 0.2 * h / 1500.0
    g = int(np.gcd.reduce(dims)); smax = min(g & -g, 8)
    # pass 1 as the ranks would do it: each counts its share of the coarse cells
    S = min(smax, octree.bootstrap_size(dims, 1 << (max(dims) - 1).bit_length(), world))
    grid = ol.CoarseGrid(dims, S)
    vs_tab = np.array([m[1] for m in bench.BASIN_MATS])
    model = ol.GridModel(dims, 4, mat_of, vs_tab) if native else None
    shares = [ol.leaf_counts(grid, np.arange(r * grid.n // world, (r + 1) * grid.n // world),
                             lambda x, y, z: vs_tab[mat_of(x, y, z)], h * ppw * fmax, chunk=300, threads=2, model=model)
              for r in range(world)]
    for r in range(world):
        mesh, info = ol.octree_halfspace_local(dims, smax, h, dt, list(bench.BASIN_MATS), mat_of, ppw, fmax, r, world,
                                               damping=hb.BKT, allgather=lambda mine: shares, chunk=300, threads=2, model=model)
        assert info["local_region_elements"] < info["etotal"]
        _same_tables(mesh, info, *whole[r])


def test_local_build_on_one_rank_equals_the_whole_mesh():
    """world = 1 (bench.py --workload basin on one GPU): X is the whole domain and the tables are those of the
    numpy whole-mesh build, nTable included bit for bit."""
    import bench
    import hercules_b200 as hb
    a, ai = bench.basin_workload(256, hb.BKT)[:2]
    b, bi = bench.basin_workload(256, hb.BKT, (0, 1), local=True, threads=2)[:2]
    assert bi["local_region_elements"] == bi["E"] == ai["E"] and bi["N"] == ai["N"]
    for k in ("elem_lnid", "eTable", "nTable", "dnode", "edata"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(np.stack(ai["node_xyz"]), np.stack(bi["node_xyz"]))
    assert np.array_equal(np.stack(ai["elem_xyz"]), np.stack(bi["elem_xyz"])) and np.array_equal(ai["elem_size"], bi["elem_size"])
    assert b.dn_c.peer.size == b.dn_s.peer.size == b.an_c.peer.size == b.an_s.peer.size == 0 and (bi["owner"] == 0).all()


@pytest.mark.parametrize("n", [128, 256])
def test_bench_basin_elements_do_not_straddle_materials(n):
    """mesh_correct_properties (psolve.c:7103-7200) gives an element the mean of 27 samples (0.005 / 0.5 / 0.995 of
    its edge along every axis); the meshers here sample the centre.  The two agree exactly when no element sees
    two materials -- which is what makes the bench's basin tables the reference's: checked here."""
    import bench
    import hercules_b200 as hb
    mesh, info = bench.basin_workload(n, hb.BKT, (0, 1), local=True, threads=2)[:2]
    mat_of = info["mat_of"]
    ex, ey, ez = info["elem_xyz"]
    es = info["elem_size"]
    centre = mat_of(ex + 0.5 * es, ey + 0.5 * es, ez + 0.5 * es)
    for a in (0.005, 0.5, 0.995):
        for b in (0.005, 0.5, 0.995):
            for c in (0.005, 0.5, 0.995):
                assert np.array_equal(mat_of(ex + a * es, ey + b * es, ez + c * es), centre), (a, b, c)


@pytest.mark.parametrize("seed", range(6))
def test_native_primitives_equal_the_numpy_restatement_on_random_models(seed):
    """Random piecewise-constant models (random cells of edge cl, four materials whose Vs make the vs rule ask for
    every level, so isolated fine octants sit in coarse surroundings and balancing has to ripple): the native
    refine + balance on random chunks, and the native extraction on the whole domain, against octree.py's numpy
    restatement (itself pinned on the reference's meshes) -- leaves, node order, lnid, hanging-node table, all exact."""
    from hercules_b200 import octree as oc, octree_local as ol
    rng = np.random.default_rng(100 + seed)
    S = 8
    dims = tuple(int(S * rng.integers(2, 5)) for _ in range(3))
    cl = int(rng.choice([1, 2, 4]))
    g = tuple(-(-n // cl) for n in dims)
    # stiff cells (leaves of edge 8) mixed with softer ones down to edge 1: the vs rule looks at octant centres only
    grid_mat = rng.choice(4, size=g, p=[0.55, 0.2, 0.15, 0.1]).astype(np.int64)
    vs_tab = np.array([8.5, 4.5, 2.5, 1.2])                      # factor_h = 1: an octant of edge s splits while s > Vs

    def mat_of(x, y, z):
        return grid_mat[np.floor(np.asarray(x) / cl).astype(int), np.floor(np.asarray(y) / cl).astype(int),
                        np.floor(np.asarray(z) / cl).astype(int)]

    def vs_of(x, y, z):
        return vs_tab[mat_of(x, y, z)]
    grid = ol.CoarseGrid(dims, S)
    model = ol.GridModel(dims, cl, mat_of, vs_tab)
    assert np.array_equal(model.grid, grid_mat.astype(np.uint8))
    # random chunks
    for _ in range(4):
        sub = np.sort(rng.choice(grid.n, size=int(rng.integers(1, grid.n + 1)), replace=False))
        a = ol._chunk_leaves(grid, sub, vs_of, 1.0)
        b = ol._chunk_leaves_native(grid, sub, model, 1.0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert np.array_equal(ol._chunk_leaves_native(grid, sub, model, 1.0, want_leaves=False)[2], a[2])
    # the whole domain: the chunked leaves are the whole-domain balanced refinement
    whole = oc.balance(oc.refine(dims, S, vs_of, 1.0), dims)
    X = np.arange(grid.n)
    codes, sizes, per_cell = ol.exact_leaves(grid, X, None, 1.0, chunk=5, threads=2, model=model)
    wc = np.concatenate([oc._code(*v) for v in whole.values()])
    ws = np.concatenate([np.full(v[0].size, s, np.int64) for s, v in whole.items()])
    o = np.argsort(wc)
    assert np.array_equal(codes, wc[o]) and np.array_equal(sizes, ws[o])
    assert len(set(sizes.tolist())) >= 3                         # several levels, or the test says nothing
    # extraction
    (ex, ey, ez, es), (px, py, pz), lnid, dnode = oc.extract(whole, dims)
    lstart = np.concatenate([[0], np.cumsum(per_cell)])
    (nex, ney, nez, nes), (npx, npy, npz), nlnid, ndnode, dang, holder, trash = ol.extract_native(grid, X, codes, sizes, lstart,
                                                                                               chunk=3, threads=2)
    assert np.array_equal(np.stack([ex, ey, ez, es]), np.stack([nex, ney, nez, nes]))
    assert trash == px.size and np.array_equal(np.stack([px, py, pz]), np.stack([npx, npy, npz])[:, :-1])
    assert np.array_equal(lnid, nlnid) and nlnid.max() < trash   # X is the whole domain: nothing is outside
    assert dnode.shape == ndnode.shape and np.array_equal(dnode, ndnode) and dnode.shape[0] > 0
    assert np.array_equal(np.nonzero(dang)[0], dnode[:, 0])
    # holders: the leaf whose half-open box holds the node (far faces pulled in)
    q = [np.minimum(p, n - 1) for p, n in zip((px, py, pz), dims)]
    h = holder[:-1]
    assert (h >= 0).all()
    assert all(((c[h] <= qq) & (qq < c[h] + es[h])).all() for c, qq in zip((ex, ey, ez), q))


def test_native_primitives_reject_bad_input():
    from hercules_b200 import octree_local as ol
    dims = (16, 16, 8)
    grid = ol.CoarseGrid(dims, 8)
    model = ol.GridModel(dims, 4, lambda x, y, z: np.zeros(np.shape(x), np.int64), np.array([100.0]))
    with pytest.raises(RuntimeError, match="hmesh_chunk_leaves failed"):
        ol._chunk_leaves_native(grid, np.array([1, 0]), model, 1.0)          # cells not in ascending Morton order
    with pytest.raises(ValueError, match="power of two"):
        ol.CoarseGrid(dims, 6)
    with pytest.raises(ValueError, match="power of two"):
        ol.CoarseGrid((16, 16, 12), 8)                                       # edge does not divide the domain
    codes, sizes, per_cell = ol.exact_leaves(grid, np.arange(grid.n), None, 1.0, model=model)
    assert codes.size == grid.n and (sizes == 8).all() and (per_cell == 1).all()      # stiff everywhere: the coarse cells


@pytest.mark.parametrize("seed,world", [(0, 2), (1, 3), (2, 5), (3, 7)])
def test_local_build_equals_whole_mesh_cut_on_random_models(seed, world):
    """The one-ring argument under stress: random models with isolated fine octants anywhere, block boundaries
    wherever the leaf counts put them.  Every rank's tables from its own neighbourhood (native path) must be the
    cut of the whole mesh: elements, numbering, hanging nodes with anchors on other ranks, ownership, share
    lists, all four schedules, owned nTable rows."""
    from hercules_b200 import octree as oc, octree_local as ol
    rng = np.random.default_rng(500 + seed)
    dims, cl = (64, 32, 32), 2
    g = tuple(n // cl for n in dims)
    grid_mat = rng.choice(4, size=g, p=[0.6, 0.2, 0.12, 0.08]).astype(np.int64)
    mats = [(3.0 * v, v, 2000.0) for v in (8.5, 4.5, 2.5, 1.2)]             # (Vp, Vs, rho); h * ppw * fmax = 1

    def mat_of(x, y, z):
        return grid_mat[np.floor(np.asarray(x) / cl).astype(int), np.floor(np.asarray(y) / cl).astype(int),
                        np.floor(np.asarray(z) / cl).astype(int)]
    counts = model = None
    for r in range(world):
        whole = oc.octree_halfspace_part(dims, 8, 1.0, 1e-3, mats, mat_of, 1.0, 1.0, r, world)
        mesh, info = ol.octree_halfspace_local(dims, 8, 1.0, 1e-3, mats, mat_of, 1.0, 1.0, r, world, counts=counts,
                                               model=model, model_cell=cl, chunk=11, threads=2)
        counts, model = info["counts"], info["model"]
        assert len(set(info["elem_size"].tolist())) >= 2 and info["D"] > 0
        assert world == 2 or info["local_region_elements"] < 0.9 * info["etotal"]      # a real subset of the mesh
        _same_tables(mesh, info, *whole)
