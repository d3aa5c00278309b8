"""CPU test: the synthetic uniform-mesh generator reproduces, bit for bit, the mesh and solver
tables the unmodified reference produced for the same domain (octor ordering + solver_init)."""
import numpy as np

from conftest import load_golden, params_of


def test_uniform_mesh_matches_octor_and_solver_init():
    from hercules_b200 import meshgen
    g = load_golden("uniform_rayleigh_eff"); P = params_of(g)
    assert g["elem_lnid"].shape[0] == 16 * 16 * 8
    mesh, info = meshgen.uniform_halfspace(16, 16, 8, h=62.5, dt=P["dt"], freq=P["freq"],
                                           damping=P["damping"], exact=True)
    assert np.array_equal(mesh.elem_lnid, g["elem_lnid"])          # indexing bit-exact
    assert info["abase"] == P["abase"] and info["bbase"] == P["bbase"]
    assert np.array_equal(mesh.edata[:, :4], g["elem_edata"][:, :4])
    assert np.array_equal(mesh.eTable, g["eTable"])
    assert np.array_equal(mesh.nTable, g["nTable"])
    assert np.array_equal(mesh.K1.reshape(8, 8, 9), g["K1"]) and np.array_equal(mesh.K2.reshape(8, 8, 9), g["K2"])
    # node coordinates follow the same order
    nx, ny, nz = 16, 16, 8
    lin = info["node_order"]
    ix, iy, iz = lin // ((ny + 1) * (nz + 1)), (lin // (nz + 1)) % (ny + 1), lin % (nz + 1)
    h = int(g["node_ticks"][g["node_ticks"] > 0].min())
    assert np.array_equal(np.stack([ix, iy, iz], 1) * h, g["node_ticks"])
    # the fast (grouped) accumulation differs only by rounding
    fast, _ = meshgen.uniform_halfspace(16, 16, 8, h=62.5, dt=P["dt"], freq=P["freq"], damping=P["damping"])
    assert np.allclose(fast.nTable, g["nTable"], rtol=1e-13, atol=0)


import pytest


@pytest.mark.parametrize("name,world", [("uniform_rayleigh_eff_np3", 3), ("uniform_rayleigh_eff_np4", 4)])
def test_partition_matches_octor(name, world):
    """Mesh and partition indexing bit-exact with octor_partitiontree / octor_extractmesh and
    schedule_build on a multi-rank run of the unmodified reference: element blocks (geid), local
    node numbering, ownership, sharer lists, halo schedules; and the exchanged nTable."""
    from conftest import rank_view
    from hercules_b200 import meshgen
    g = load_golden(name)
    for r in range(world):
        v = rank_view(g, r); P = params_of(v)
        mesh, info = meshgen.uniform_halfspace(16, 16, 8, h=62.5, dt=P["dt"], freq=P["freq"],
                                               damping=P["damping"], exact=True, part=(r, world))
        assert np.array_equal(info["elem_geid"], v["elem_geid"])
        assert np.array_equal(mesh.elem_lnid, v["elem_lnid"])
        h = int(v["node_ticks"][v["node_ticks"] > 0].min())
        assert np.array_equal(np.stack(info["node_xyz"], 1) * h, v["node_ticks"])
        ismine = v["node_flags"][:, 0].astype(bool)
        assert np.array_equal(info["owner"] == r, ismine)
        assert np.array_equal(info["owner"][~ismine], v["node_owner"][~ismine])
        assert np.array_equal(info["share"], v["node_share"])
        for side in ("dn_c", "dn_s", "an_c", "an_s"):
            ml = getattr(mesh, side)
            hdr = np.stack([ml.peer, ml.nodes], 1).reshape(-1, 2)
            assert np.array_equal(hdr, v[side + "_hdr"].reshape(-1, 2)), (r, side)
            assert np.array_equal(ml.mapping, v[side + "_map"]), (r, side)
        assert np.array_equal(mesh.eTable, v["eTable"])
        assert np.array_equal(mesh.nTable, v["nTable"]), r


def test_bkt_coefficients_match_reference_edata():
    """meshgen.bkt_coefficients restates mesh_correct_properties' BKT block (psolve.c:7239-7310,
    Search_Quality_Table quake_util.c:128-163, the 18-of-26-row table psolve.c:5575-5616): bit-exact
    with the edata the unmodified reference produced for its BKT run (use_infinite_qk = yes there)."""
    import numpy as np
    from conftest import load_golden
    from hercules_b200 import meshgen
    ed = load_golden("graded2_bkt")["elem_edata"]
    c = meshgen.bkt_coefficients(ed[:, 1], ed[:, 2], use_inf_qk=True)
    assert np.array_equal(c, ed[:, 4:14])
    # finite Qk (use_infinite_qk = no) on a soft column: both families non-zero, bit-exact again
    ed = load_golden("graded2_bkt_qk")["elem_edata"]
    c = meshgen.bkt_coefficients(ed[:, 1], ed[:, 2], use_inf_qk=False)
    assert np.array_equal(c, ed[:, 4:14]) and (c[:, 5:] != 0).all()
    # finite Qk: every element of a stiff layer still finds a table row, soft ones differ
    c2 = meshgen.bkt_coefficients(np.float32([4000, 6000, 1500]), np.float32([2000, 3464, 500]))
    assert c2.shape == (3, 10) and (c2[:, :5] > 0).all()


@pytest.mark.parametrize("name,grid,bands,h,layers", [
    ("graded3_rayleigh_eff", (32, 32), ((2, 1), (3, 2), (2, 4)), 31.25,
     [(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)]),
    ("graded2_rayleigh_eff", (16, 16), ((2, 1), (3, 2)), 62.5,
     [(0, 3000, 1732, 2000), (125, 6000, 3464, 2700)]),
])
def test_graded_mesh_matches_octor_and_solver_init(name, grid, bands, h, layers):
    """The adaptive (hanging-node) generator reproduces octor's refined + balanced mesh of a banded
    half-space bit for bit: leaf order, node numbering, elem_t.lnid, dnodeTable (ids, deps, anchor
    list order), and solver_init's eTable / nTable after compute_adjust(DISTRIBUTION)."""
    from hercules_b200 import meshgen
    g = load_golden(name); P = params_of(g)
    mesh, info = meshgen.graded_halfspace(*grid, bands, h=h, dt=P["dt"], freq=P["freq"],
                                          damping=P["damping"], layers=layers, exact=True)
    assert np.array_equal(mesh.elem_lnid, g["elem_lnid"])
    assert np.array_equal(mesh.dnode, g["dnode"]) and info["D"] == g["dnode"].shape[0] > 0
    tick = int(g["node_ticks"][g["node_ticks"] > 0].min())
    assert np.array_equal(np.stack(info["node_xyz"], 1) * tick, g["node_ticks"])
    lvl = g["elem_level"].astype(np.int64)
    assert np.array_equal(info["elem_size"], 2 ** (lvl.max() - lvl))
    assert np.array_equal(mesh.edata[:, :4], g["elem_edata"][:, :4])
    assert np.array_equal(mesh.eTable, g["eTable"])
    assert np.array_equal(mesh.nTable, g["nTable"])
    fast, _ = meshgen.graded_halfspace(*grid, bands, h=h, dt=P["dt"], freq=P["freq"],
                                       damping=P["damping"], layers=layers)
    assert np.allclose(fast.nTable, g["nTable"], rtol=1e-13, atol=0)


L3 = [(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)]
L2 = [(0, 3000, 1732, 2000), (125, 6000, 3464, 2700)]


@pytest.mark.parametrize("name,world,grid,bands,h,layers", [
    ("graded3_rayleigh_eff_np2", 2, (32, 32), ((2, 1), (3, 2), (2, 4)), 31.25, L3),
    # with 4 ranks the reference's octor leaves no level-3 octants: two levels, 3840 leaves
    ("graded3_rayleigh_eff_np4", 4, (32, 32), ((2, 1), (7, 2)), 31.25, L3),
    ("graded2_bkt_np2", 2, (16, 16), ((2, 1), (3, 2)), 62.5, L2),
])
def test_graded_partition_matches_octor(name, world, grid, bands, h, layers):
    """Partitioned adaptive meshes: element blocks, local node numbering, ownership, anchored flags,
    share lists, the OWNED dangling-node table with its anchor lists, and all four halo schedules
    (dangling / anchored x contribute / share) bit-exact with multi-rank runs of the unmodified
    reference; nTable rows of owned nodes equal to rounding (complete sums in single-rank order)."""
    from conftest import rank_view
    from hercules_b200 import meshgen
    g = load_golden(name)
    for r in range(world):
        v = rank_view(g, r); P = params_of(v)
        mesh, info = meshgen.graded_halfspace(*grid, bands, h=h, dt=P["dt"], freq=P["freq"], damping=P["damping"],
                                              layers=layers, exact=True, part=(r, world))
        assert np.array_equal(info["elem_geid"], v["elem_geid"])
        assert np.array_equal(mesh.elem_lnid, v["elem_lnid"])
        tick = int(v["node_ticks"][v["node_ticks"] > 0].min())
        assert np.array_equal(np.stack(info["node_xyz"], 1) * tick, v["node_ticks"])
        ismine = v["node_flags"][:, 0].astype(bool)
        assert np.array_equal(info["owner"] == r, ismine)
        assert np.array_equal(info["owner"][~ismine], v["node_owner"][~ismine])
        assert np.array_equal(info["anchored"], v["node_flags"][:, 1].astype(bool))
        assert np.array_equal(info["share"], v["node_share"])
        assert np.array_equal(mesh.dnode, v["dnode"])
        for side in ("dn_c", "dn_s", "an_c", "an_s"):
            ml = getattr(mesh, side)
            hdr = np.stack([ml.peer, ml.nodes], 1).reshape(-1, 2)
            assert np.array_equal(hdr, v[side + "_hdr"].reshape(-1, 2)), (r, side)
            assert np.array_equal(ml.mapping, v[side + "_map"]), (r, side)
        assert np.array_equal(mesh.eTable, v["eTable"])
        if P["damping"] != 3:
            assert np.array_equal(mesh.edata[:, :4], v["elem_edata"][:, :4])
        assert np.allclose(mesh.nTable[ismine], v["nTable"][ismine], rtol=2e-15, atol=0)
        assert (mesh.nTable[:, 0] > 0).all()


def test_column_regions():
    from hercules_b200 import meshgen
    assert meshgen.column_regions(32, 32, 16, 2) == [(0, 32, 0, 16), (0, 32, 16, 32)]
    assert meshgen.column_regions(32, 32, 16, 4) == [(0, 16, 0, 16), (16, 32, 0, 16), (0, 16, 16, 32), (16, 32, 16, 32)]
    assert meshgen.column_regions(1024, 512, 512, 2) == [(0, 512, 0, 512), (512, 1024, 0, 512)]
    r8 = meshgen.column_regions(2048, 1024, 512, 8)
    assert r8[:4] == [(0, 512, 0, 512), (512, 1024, 0, 512), (0, 512, 512, 1024), (512, 1024, 512, 1024)] and r8[4][0] == 1024
    with pytest.raises(ValueError):
        meshgen.column_regions(32, 32, 32, 2)          # columns 16 wide would interleave in z


def test_uniform_mesh_configs0_domain():
    """BASELINE.json configs[0] (examples/test1): 100 x 100 x 37.5 km, tick ratio 8:8:3, 32 x 32 x 12
    elements -- a domain whose depth is not a power of two -- bit-exact with the reference's mesh."""
    from hercules_b200 import meshgen
    g = load_golden("test1_homogeneous"); P = params_of(g)
    mesh, info = meshgen.uniform_halfspace(32, 32, 12, h=3125.0, dt=P["dt"], freq=P["freq"], damping=P["damping"], exact=True)
    assert np.array_equal(mesh.elem_lnid, g["elem_lnid"])
    assert np.array_equal(mesh.eTable, g["eTable"]) and np.array_equal(mesh.nTable, g["nTable"])


def _pairwise_schedules_match(meshes, infos):
    """c-lists and s-lists of every pair of ranks name the same nodes (by coordinates) in the same
    order -- what the halo exchange relies on (psolve.c:4945-5079: the k-th value a messenger packs
    is the k-th value its counterpart unpacks)."""
    world = len(meshes)
    total = 0
    for side_c, side_s in (("an_c", "an_s"), ("dn_c", "dn_s")):
        for a in range(world):
            ca = getattr(meshes[a], side_c)
            off = np.concatenate([[0], np.cumsum(ca.nodes)]).astype(int)
            for i, b in enumerate(ca.peer.tolist()):
                sb = getattr(meshes[b], side_s)
                j = sb.peer.tolist().index(a)
                offb = np.concatenate([[0], np.cumsum(sb.nodes)]).astype(int)
                na, nb_ = ca.mapping[off[i]:off[i + 1]], sb.mapping[offb[j]:offb[j + 1]]
                assert na.size == nb_.size > 0
                xa = np.stack([c[na] for c in infos[a]["node_xyz"]], 1)
                xb = np.stack([c[nb_] for c in infos[b]["node_xyz"]], 1)
                assert np.array_equal(xa, xb), (side_c, a, b)
                assert (infos[a]["owner"][na] == b).all() and (infos[b]["owner"][nb_] == b).all()
                total += na.size
            # every s-list entry has a c-list counterpart
            sa = getattr(meshes[a], side_s)
            for b in sa.peer.tolist():
                assert a in getattr(meshes[b], side_c).peer.tolist()
    return total


@pytest.mark.parametrize("world", [2, 8])
def test_bench_partitions_are_consistent(world):
    """The weak-scaling domains bench.py builds for N GPUs (uniform: N Morton blocks; adaptive: N
    columns), at 1/16 of their edge: every node owned exactly once, schedules pairwise consistent."""
    import bench
    from hercules_b200 import meshgen
    n = 16
    bx, by, bz = bench.block_grid(world)
    parts = [meshgen.uniform_halfspace(n * bx, n * by, n * bz, h=25.0, dt=0.002, part=(r, world)) for r in range(world)]
    assert _pairwise_schedules_match([p[0] for p in parts], [p[1] for p in parts]) > 0
    owned = sum(int((p[1]["owner"] == r).sum()) for r, p in enumerate(parts))
    assert owned == (n * bx + 1) * (n * by + 1) * (n * bz + 1)
    assert all(p[1]["E"] == n ** 3 for p in parts)
    n = 64
    cx, cy = bench.column_grid(world)
    parts = [meshgen.graded_halfspace(n * cx, n * cy, bench.adaptive_bands(n), h=25.0, dt=0.002,
                                      layers=bench.adaptive_layers(n), part=(r, world)) for r in range(world)]
    assert _pairwise_schedules_match([p[0] for p in parts], [p[1] for p in parts]) > 0
    whole, winfo = meshgen.graded_halfspace(n * cx, n * cy, bench.adaptive_bands(n), h=25.0, dt=0.002,
                                            layers=bench.adaptive_layers(n))
    assert sum(int((p[1]["owner"] == r).sum()) for r, p in enumerate(parts)) == winfo["N"]
    assert sum(p[1]["D"] for p in parts) == winfo["D"]
    assert sum(p[1]["E"] for p in parts) == winfo["E"]


@pytest.mark.parametrize("world", [2, 4])
def test_bench_strong_scaling_partitions_are_consistent(world):
    """bench.py --workload adaptive --strong: ONE mesh of 8 columns cut over 2 or 4 ranks (2 x 2 and 2 x 1
    columns per rank), at 1/6 of its edge."""
    import bench
    from hercules_b200 import meshgen
    n = 64
    cx, cy = bench.column_grid(8)
    parts = [meshgen.graded_halfspace(n * cx, n * cy, bench.adaptive_bands(n), h=25.0, dt=0.002,
                                      layers=bench.adaptive_layers(n), part=(r, world)) for r in range(world)]
    assert _pairwise_schedules_match([p[0] for p in parts], [p[1] for p in parts]) > 0
    assert len({p[1]["E"] for p in parts}) == 1 and sum(p[1]["E"] for p in parts) == parts[0][1]["etotal"]
    whole = (n * cx + 1) * (n * cy + 1)            # nodes per fine plane; every node owned exactly once:
    owned = sum(int((p[1]["owner"] == r).sum()) for r, p in enumerate(parts))
    _, winfo = meshgen.graded_halfspace(n * cx, n * cy, bench.adaptive_bands(n), h=25.0, dt=0.002, layers=bench.adaptive_layers(n))
    assert owned == winfo["N"] and whole > 0
