"""CPU tests of the host-side index builders behind hgpu_init (no GPU needed): the tile plan is
built and independently re-checked against the mesh inside hgpu_plan_build (every node owned once,
every element the core element of exactly one tile, slots decode to the element's own corners, and
for every owned node the tile's own entries plus the partial forces it reads from lower tiles
account for each incident element exactly once)."""
import numpy as np
import pytest

from conftest import load_golden, rank_view


@pytest.mark.parametrize("tile_nodes", [0, 2, 64, 200])
@pytest.mark.parametrize("name", ["graded2_rayleigh_eff", "graded3_rayleigh_eff", "uniform_rayleigh_eff",
                                  "graded3_rayleigh_eff_np4", "basin_rayleigh_eff", "basin_corner_rayleigh_eff",
                                  "test1_homogeneous"])
def test_tile_plan_valid_on_octor_meshes(name, tile_nodes):
    from hercules_b200 import solver
    g = rank_view(load_golden(name), 1)
    E, N = g["elem_lnid"].shape[0], g["nTable"].shape[0]
    r = solver.plan_build(g["elem_lnid"], N, tile_nodes)
    assert r["ntiles"] >= 1 and r["tile_elems_total"] == E      # every element evaluated exactly once
    assert r["early_tiles"] == 0
    assert r["smem_bytes"] <= 115712                      # two CTAs per SM on a B200
    if tile_nodes:
        assert r["tile_nodes"] <= max(2, tile_nodes)


@pytest.mark.parametrize("tile_nodes", [0, 64])
@pytest.mark.parametrize("name,nranks", [("graded3_rayleigh_eff_np2", 2), ("graded3_rayleigh_eff_np4", 4),
                                         ("uniform_rayleigh_eff_np3", 3)])
def test_tile_plan_self_tiles_on_partitioned_meshes(name, nranks, tile_nodes):
    """Multi-rank meshes: tiles that own a node of a halo schedule or of the hanging-node lists
    evaluate the foreign elements incident to their nodes themselves and wait for nobody."""
    from hercules_b200 import solver
    for rank in range(nranks):
        g = rank_view(load_golden(name), rank)
        hm = solver.HostMesh.from_dump(g)
        E, N = g["elem_lnid"].shape[0], g["nTable"].shape[0]
        r = solver.plan_build(g["elem_lnid"], N, tile_nodes, mesh=hm)
        assert r["early_tiles"] >= 1 and r["tile_elems_total"] >= E
        assert r["smem_bytes"] <= 115712


def test_tile_plan_uniform_layout():
    """Aligned 8x8x8 element blocks on a uniform mesh: 8^3 core elements, 9^3 staged nodes of which
    8^3 are owned (interior tiles), no element evaluated twice, near conflict-free slots."""
    from hercules_b200 import meshgen, solver
    mesh, info = meshgen.uniform_halfspace(32, 32, 32, h=25.0, dt=0.002)
    r = solver.plan_build(mesh.elem_lnid, info["N"])
    assert r["max_tile_elems"] == 512
    assert r["tile_elems_total"] == info["E"]
    assert r["max_tile_nodes"] <= 752 and r["max_tile_acc"] <= 752
    assert r["tile_halo_total"] == r["partial_slots"] or r["partial_slots"] >= r["tile_halo_total"]
    assert r["est_gather_wavefronts"] < 3.3 and r["est_scatter_wavefronts"] < 3.3


def test_tile_plan_rejects_bad_mesh():
    from hercules_b200 import solver
    lnid = np.arange(8, dtype=np.int32).reshape(1, 8)
    with pytest.raises(solver.HerculesGpuError, match="out of range"):
        solver.plan_build(lnid, 4)


@pytest.mark.parametrize("world,rank", [(2, 0), (4, 3), (8, 4)])
def test_tile_plan_on_partitioned_adaptive_workload(world, rank):
    """bench.py --workload adaptive --gpus N (configs[3]) at 1/512 of its size: the plan builds and
    self-checks on a rank's share of the partitioned 3-level mesh; tiles owning shared or hanging
    nodes become self tiles."""
    import bench
    from hercules_b200 import meshgen, solver
    n = 64
    cx, cy = bench.column_grid(world)
    mesh, info = meshgen.graded_halfspace(n * cx, n * cy, bench.adaptive_bands(n), h=bench.H_M, dt=bench.DT,
                                          layers=bench.adaptive_layers(n), part=(rank, world))
    r = solver.plan_build(mesh.elem_lnid, info["N"], 0, mesh=mesh)
    assert r["tile_elems_total"] >= info["E"] and r["early_tiles"] >= 1
    assert r["smem_bytes"] <= 115712


@pytest.mark.parametrize("world,rank", [(2, 1), (4, 0), (8, 5)])
def test_tile_plan_on_per_rank_basin_workload(world, rank):
    """bench.py --workload basin --gpus N (configs[4]; hercules_b200.octree_local: every rank meshes only its block
    and one ring) at --edge 256: the tile plan builds and self-checks on the rank's mesh -- 4 octree levels, BKT,
    hanging nodes and anchors on other ranks -- and the mesh handed to hgpu_init is self-consistent."""
    import bench
    import hercules_b200 as hb
    from hercules_b200 import solver
    mesh, info = bench.basin_workload(256, hb.BKT, (rank, world), local=True, threads=2)[:2]
    E, N = info["E"], info["N"]
    assert mesh.elem_lnid.shape == (E, 8) and mesh.elem_lnid.min() == 0 and mesh.elem_lnid.max() == N - 1
    assert mesh.nTable.shape == (N, 7) and (mesh.nTable[info["owner"] == rank, 0] > 0).all()       # owned masses complete
    assert mesh.dnode.shape[0] == info["D"] > 0 and mesh.dnode[:, 0].max() < N
    for ml in (mesh.dn_c, mesh.dn_s, mesh.an_c, mesh.an_s):
        assert ml.mapping.size == ml.nodes.sum() and (ml.mapping.size == 0 or ml.mapping.max() < N)
        assert rank not in ml.peer.tolist()
    r = solver.plan_build(mesh.elem_lnid, N, 0, mesh=mesh)
    assert r["tile_elems_total"] >= E and r["early_tiles"] >= 1
    assert r["smem_bytes"] <= 115712
