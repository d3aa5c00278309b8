"""CPU tests: hercules_b200.octree (refine by the vs rule, 2:1 balance across faces and edges, mesh
extraction with hanging nodes, solver tables) against meshes the unmodified reference produced with
octor from synthetic material etrees -- layered models and a laterally varying one (a soft box in a
stiff half-space: hanging nodes on faces and edges of all three directions, a refinement level that
only exists because of balancing)."""
import numpy as np
import pytest

from conftest import load_golden, params_of


def _leaves_of(g):
    tick = int(g["node_ticks"][g["node_ticks"] > 0].min())
    nt, ln = g["node_ticks"] // tick, g["elem_lnid"]
    unit = int((nt[ln[:, 1], 0] - nt[ln[:, 0], 0]).min())       # finest edge in node-tick units
    nt = nt // unit
    lv = g["elem_level"].astype(int)
    sz = 2 ** (lv.max() - lv)
    return {int(s): tuple(nt[ln[sz == s, 0]][:, c] for c in range(3)) for s in np.unique(sz)}, nt


CASES = {   # name: (dims in h, h, smax, cvm leaf in h, materials, basin box or None, layer tops, vs_min, ppw, fmax)
    "basin_rayleigh_eff": ((32, 32, 16), 31.25, 8, 2, [(6000., 3464., 2700.), (1800., 866., 1800.)],
                           (375., 750., 250., 625., 125.), [0.0], 800., 8., 2.5),
    # the box in a corner of the domain, through its whole depth: hanging nodes ON absorbing faces and domain edges
    "basin_corner_rayleigh_eff": ((32, 32, 16), 31.25, 8, 2, [(6000., 3464., 2700.), (1800., 866., 1800.)],
                                  (0., 250., 0., 375., 500.), [0.0], 800., 8., 2.5),
    "graded3_rayleigh_eff": ((32, 32, 16), 31.25, 8, 2, [(1800., 866., 1800.), (3000., 1732., 2000.), (6000., 3464., 2700.)],
                             None, [0.0, 62.5, 250.0], 800., 8., 2.5),
    "graded2_rayleigh_eff": ((16, 16, 8), 62.5, 4, 2, [(3000., 1732., 2000.), (6000., 3464., 2700.)],
                             None, [0.0, 125.0], 800., 8., 2.5),
    "test1_homogeneous": ((32, 32, 12), 3125.0, 4, 4, [(6000., 3464., 2700.)], None, [0.0], 500., 8., 0.1),
}


def _mat_of(case):
    dims, h, smax, cl, mats, box, tops, *_ = case

    def f(x, y, z):
        # cvm_query: the material of the etree leaf (edge cl * h) that holds the point; x = north, y = east
        cx, cy, cz = ((np.floor(np.asarray(v) / cl) + 0.5) * cl * h for v in (x, y, z))
        if box is not None:
            e0, e1, n0, n1, zb = box
            return ((cy >= e0) & (cy < e1) & (cx >= n0) & (cx < n1) & (cz < zb)).astype(np.int64)
        m = np.zeros(np.shape(cz), np.int64)
        for k, zt in enumerate(tops):
            m[cz >= zt] = k
        return m
    return f


@pytest.mark.parametrize("name", ["basin_rayleigh_eff", "basin_corner_rayleigh_eff", "graded3_rayleigh_eff",
                                  "graded2_rayleigh_eff", "uniform_rayleigh_eff", "test1_homogeneous"])
def test_extract_reproduces_octor(name):
    """octor's own leaves in: leaf order, node numbering, elem_t.lnid and the dangling-node table
    (ids, deps, anchors in list order) out, bit for bit."""
    from hercules_b200 import octree
    g = load_golden(name)
    leaves, nt = _leaves_of(g)
    dims = tuple(int(v) for v in nt.max(0))
    (ex, ey, ez, es), (px, py, pz), lnid, dnode = octree.extract(leaves, dims)
    assert np.array_equal(lnid, g["elem_lnid"])
    assert np.array_equal(np.stack([px, py, pz], 1), nt)
    assert dnode.shape == g["dnode"].shape and np.array_equal(dnode, g["dnode"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_mesher_reproduces_octor_and_solver_init(name):
    """Material model in: the refined + balanced leaf set, the mesh and solver_init's tables out, bit
    for bit what the unmodified reference built from the same model."""
    from hercules_b200 import octree
    g = load_golden(name); P = params_of(g)
    dims, h, smax, cl, mats, box, tops, vs_min, ppw, fmax = CASES[name]
    mesh, info = octree.octree_halfspace(dims, smax, h, P["dt"], mats, _mat_of(CASES[name]), ppw, fmax, freq=P["freq"],
                                         damping=P["damping"], vs_min=vs_min, exact=True)
    assert mesh.elem_lnid.shape == g["elem_lnid"].shape and np.array_equal(mesh.elem_lnid, g["elem_lnid"])
    assert np.array_equal(mesh.dnode, g["dnode"])
    lv = g["elem_level"].astype(int)
    assert np.array_equal(info["elem_size"], 2 ** (lv.max() - lv))
    assert np.array_equal(mesh.edata[:, :4], g["elem_edata"][:, :4])
    assert np.array_equal(mesh.eTable, g["eTable"])
    assert np.array_equal(mesh.nTable, g["nTable"])


def test_balance_is_two_to_one_across_faces_and_edges():
    """Random refinement in: after balance every pair of leaves sharing a face or an edge differs by at
    most one level, nothing was coarsened, and a second pass changes nothing."""
    from hercules_b200 import octree
    rng = np.random.default_rng(3)
    dims, smax = (32, 32, 32), 16
    spots = rng.integers(0, 32, (6, 3))

    def vs(x, y, z):
        # twice the Chebyshev distance from the octant's centre to the nearest spot: an octant of edge s
        # is split (s > vs) exactly when a spot lies inside it, so isolated unit octants sit in coarse ones
        return 2.0 * np.min([np.maximum(np.maximum(np.abs(x - a - 0.25), np.abs(y - b - 0.25)), np.abs(z - c - 0.25))
                             for a, b, c in spots], axis=0)
    lv = octree.refine(dims, smax, vs, 1.0)
    lb = octree.balance(lv, dims)
    assert sum(v[0].size * s ** 3 for s, v in lb.items()) == 32 ** 3
    again = octree.balance(lb, dims)
    assert {s: v[0].size for s, v in again.items()} == {s: v[0].size for s, v in lb.items()}
    size = np.zeros((32, 32, 32), np.int64)
    for s, (x, y, z) in lb.items():
        for a, b, c in zip(x, y, z):
            size[a:a + s, b:b + s, c:c + s] = s
    for dx, dy, dz in octree._DIRS18:
        a = size[max(dx, 0):32 + min(dx, 0), max(dy, 0):32 + min(dy, 0), max(dz, 0):32 + min(dz, 0)]
        b = size[max(-dx, 0):32 + min(-dx, 0), max(-dy, 0):32 + min(-dy, 0), max(-dz, 0):32 + min(-dz, 0)]
        assert (np.maximum(a, b) <= 2 * np.minimum(a, b)).all(), (dx, dy, dz)
    before = np.zeros((32, 32, 32), np.int64)
    for s, (x, y, z) in lv.items():
        for a, b, c in zip(x, y, z):
            before[a:a + s, b:b + s, c:c + s] = s
    assert (size <= before).all() and (size < before).any()


PART_CASES = {   # golden: (single-rank case it derives from, ranks)
    "basin_rayleigh_eff_np2": ("basin_rayleigh_eff", 2), "basin_rayleigh_eff_np3": ("basin_rayleigh_eff", 3),
    "basin_rayleigh_eff_np4": ("basin_rayleigh_eff", 4), "graded3_rayleigh_eff_np2": ("graded3_rayleigh_eff", 2),
    "graded3_rayleigh_eff_np4": ("graded3_rayleigh_eff", 4), "graded3_rayleigh_eff_np8": ("graded3_rayleigh_eff", 8),
}


@pytest.mark.parametrize("name", sorted(PART_CASES))
def test_partition_reproduces_octor(name):
    """The general partition against multi-rank runs of the unmodified reference: Morton blocks that cut
    through refinement levels, the coarser-leaf limit of octor's multi-rank bootstrap (4 ranks: no 125 m
    octants), nodes harbored only because they are owned or anchor an owned dangling node, share lists
    with indirect sharers first, owned dangling tables, all four schedules -- bit for bit; nTable rows of
    owned nodes to rounding (complete sums)."""
    from conftest import rank_view
    from hercules_b200 import octree
    base, world = PART_CASES[name]
    dims, h, smax, cl, mats, box, tops, vs_min, ppw, fmax = CASES[base]
    g = load_golden(name)
    for r in range(world):
        v = rank_view(g, r); P = params_of(v)
        mesh, info = octree.octree_halfspace_part(dims, smax, h, P["dt"], mats, _mat_of(CASES[base]), ppw, fmax, r, world,
                                                  freq=P["freq"], damping=P["damping"], vs_min=vs_min, exact=True)
        assert np.array_equal(info["elem_geid"], v["elem_geid"])
        assert mesh.elem_lnid.shape == v["elem_lnid"].shape and np.array_equal(mesh.elem_lnid, v["elem_lnid"])
        tick = int(v["node_ticks"][v["node_ticks"] > 0].min()) // int(np.stack(info["node_xyz"], 1)[np.stack(info["node_xyz"], 1) > 0].min())
        assert np.array_equal(np.stack(info["node_xyz"], 1) * tick, v["node_ticks"])
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


@pytest.mark.parametrize("local", [False, True], ids=["whole-mesh-cut", "per-rank"])
@pytest.mark.parametrize("world", [2, 3])
def test_bench_basin_partitions_are_consistent(world, local):
    """bench.py --workload basin --gpus N at --edge 128: every node owned once, every element on one rank,
    the halo schedules of every pair of ranks name the same nodes in the same order.  local: the per-rank
    mesher bench.py uses (hercules_b200.octree_local), otherwise the whole-mesh cut."""
    import bench
    import hercules_b200 as hb
    from test_meshgen import _pairwise_schedules_match
    parts = [bench.basin_workload(128, hb.BKT, (r, world), local=local)[:2] for r in range(world)]
    whole, winfo = bench.basin_workload(128, hb.BKT)[:2]
    assert _pairwise_schedules_match([p[0] for p in parts], [p[1] for p in parts]) > 0
    assert sum(p[1]["E"] for p in parts) == winfo["E"] == parts[0][1]["etotal"]
    assert sum(int((p[1]["owner"] == r).sum()) for r, p in enumerate(parts)) == winfo["N"]
    assert sum(p[1]["D"] for p in parts) == winfo["D"]
