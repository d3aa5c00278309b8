"""CPU tests of the host-side index builders behind hgpu_init (no GPU needed): the owner-computes
tile plan is built and independently re-checked against the mesh inside hgpu_plan_build (every
node owned once, every incident element evaluated once per owner tile, slots decode to the
element's own corners)."""
import numpy as np
import pytest

from conftest import load_golden, rank_view


@pytest.mark.parametrize("tile_nodes", [0, 2, 64, 200])
@pytest.mark.parametrize("name", ["graded2_rayleigh_eff", "graded3_rayleigh_eff", "uniform_rayleigh_eff",
                                  "graded3_rayleigh_eff_np4"])
def test_tile_plan_valid_on_octor_meshes(name, tile_nodes):
    from hercules_b200 import solver
    g = rank_view(load_golden(name), 1)
    E, N = g["elem_lnid"].shape[0], g["nTable"].shape[0]
    r = solver.plan_build(g["elem_lnid"], N, tile_nodes)
    assert r["ntiles"] >= 1 and r["tile_elems_total"] >= E
    assert r["smem_bytes"] <= 115712                      # two CTAs per SM on a B200
    if tile_nodes:
        assert r["tile_nodes"] <= max(2, tile_nodes)


def test_tile_plan_uniform_redundancy():
    """Aligned 8x8x8 node cells on a uniform mesh: 9^3 elements per 8^3 owned nodes at most."""
    from hercules_b200 import meshgen, solver
    mesh, info = meshgen.uniform_halfspace(32, 32, 32, h=25.0, dt=0.002)
    r = solver.plan_build(mesh.elem_lnid, info["N"])
    assert r["max_tile_elems"] <= 729 and r["max_tile_nodes"] <= 1008   # 10^3 staged + conflict-avoiding slack
    assert r["tile_elems_total"] / info["E"] < 1.43


def test_tile_plan_rejects_bad_mesh():
    from hercules_b200 import solver
    lnid = np.arange(8, dtype=np.int32).reshape(1, 8)
    with pytest.raises(solver.HerculesGpuError, match="out of range"):
        solver.plan_build(lnid, 4)
