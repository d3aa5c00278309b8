"""CPU worker of tests/test_halo_gloo.py: one process per rank over torch.distributed (gloo).

Every rank builds ITS share of a partitioned adaptive mesh (hercules_b200.meshgen, host logic),
steps it with the oracle's per-rank arithmetic and carries out the reference's four
schedule_senddata calls per step (psolve.c:4036-4154, 4945-5079) over gloo exactly as the
messenger lists prescribe: c-list -> owner with += in the owner's s-list order (contribution),
s-list -> sharers with = (sharing).  Rank 0 compares every rank's final displacement field with a
single-rank oracle run on the whole mesh, node by node (matched by coordinates).  This checks, with
no GPU, that partitioned meshes + schedules + complete nTable rows describe the same problem as the
whole mesh -- the host-side half of the multi-GPU path (the device half is tests/test_multirank_gpu.py).

usage: halo_worker.py <nx> <ny> <bands as n:s,n:s,...> <steps>     env: RANK WORLD_SIZE MASTER_ADDR MASTER_PORT
       halo_worker.py <nx> <ny> basin <steps>      the laterally varying model of tests/golden/basin_rayleigh_eff*.npz
                                                   through hercules_b200.octree (general mesher and partition)
       halo_worker.py <nx> <ny> basin-local <steps>  the same through hercules_b200.octree_local: per-rank meshing,
                                                   leaf counts all-gathered over gloo
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

H, DT, FREQ = 31.25, 1e-3, 2.5
LAYERS = [(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)]


BASIN_MATS = [(6000.0, 3464.0, 2700.0), (1800.0, 866.0, 1800.0)]


def basin_mat(x, y, z):
    cx, cy, cz = ((np.floor(np.asarray(v) / 2) + 0.5) * 2 * H for v in (x, y, z))
    return ((cy >= 375) & (cy < 750) & (cx >= 250) & (cx < 625) & (cz < 125)).astype(np.int64)


def exchange(dist, rank, world, snd, rcv, v, contribution):
    """One schedule_senddata: snd/rcv = the MsgLists that pack / unpack on this rank."""
    out = {int(p): v[snd.mapping[o:o + n]].copy()
           for p, o, n in zip(snd.peer, np.concatenate([[0], np.cumsum(snd.nodes)[:-1]]).astype(int), snd.nodes)}
    boxes = [None] * world
    dist.all_gather_object(boxes, out)
    off = 0
    for p, n in zip(rcv.peer.tolist(), rcv.nodes.tolist()):        # messenger after messenger, list order
        data = boxes[p][rank]
        assert data.shape[0] == n, (rank, p, data.shape, n)
        rows = rcv.mapping[off:off + n]
        if contribution:
            v[rows] += data
        else:
            v[rows] = data
        off += n


def main():
    import torch.distributed as dist
    import hercules_oracle as ho
    from hercules_b200 import meshgen, octree
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    basin = sys.argv[3] in ("basin", "basin-local")
    bands = None if basin else tuple(tuple(int(t) for t in b.split(":")) for b in sys.argv[3].split(","))
    steps = int(sys.argv[4])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ho.build() if rank == 0 else None
    dist.barrier()
    L = ho.lib()
    if sys.argv[3] == "basin-local":
        # every rank meshes only its Morton block and one ring of coarse cells (hercules_b200.octree_local, native
        # primitives); the leaf counts of the ranks' shares of the coarse cells travel over gloo
        from hercules_b200 import octree_local

        def allgather(mine):
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            return parts
        mesh, info = octree_local.octree_halfspace_local((nx, ny, 16), 8, H, DT, BASIN_MATS, basin_mat, 8.0, FREQ, rank, world,
                                                         vs_min=800.0, allgather=allgather, model_cell=2, chunk=7, threads=2)
        assert info["local_region_elements"] <= info["etotal"]
    elif basin:
        mesh, info = octree.octree_halfspace_part((nx, ny, 16), 8, H, DT, BASIN_MATS, basin_mat, 8.0, FREQ, rank, world, vs_min=800.0)
        info["dims"] = (nx, ny, 16)
    else:
        mesh, info = meshgen.graded_halfspace(nx, ny, bands, h=H, dt=DT, freq=FREQ, layers=LAYERS, part=(rank, world))
    m = ho.Mesh(mesh.elem_lnid, mesh.eTable, mesh.nTable, mesh.dnode, mesh.edata, mesh.K1, mesh.K2)
    st = ho.State(m)
    px, py, pz = info["node_xyz"]
    nz = info["dims"][2]
    gkey = (px * (ny + 1) + py) * (nz + 1) + pz
    # point source on the 8 nodes of the element whose lowest corner is (sx, sy, sz); the rank that has
    # the element loads it (a rank's loaded nodes are those of its own source elements)
    sx, sy, sz = nx // 2 - 1, ny // 2 - 1, 1
    ex, ey, ez = info["elem_xyz"]
    es_ = info["elem_size"]
    hit = np.nonzero((ex <= sx) & (sx < ex + es_) & (ey <= sy) & (sy < ey + es_) & (ez <= sz) & (sz < ez + es_))[0]
    rng = np.random.default_rng(7)
    F = 1e9 * rng.standard_normal((steps, 8, 3))
    loaded = np.sort(mesh.elem_lnid[hit[0]]).astype(np.int32) if hit.size else np.zeros(0, np.int32)
    order = np.argsort(mesh.elem_lnid[hit[0]]) if hit.size else None
    for k in range(steps):
        st.tm1, st.tm2 = st.tm2, st.tm1
        if loaded.size:
            L.ho_addforce_s(8, loaded, np.ascontiguousarray(F[k][order]).reshape(-1), DT * DT, st.force.reshape(-1))
        ho.add_forces(m, st, ho.RAYLEIGH, ho.EFFECTIVE, FREQ, DT)
        exchange(dist, rank, world, mesh.dn_c, mesh.dn_s, st.force, True)                   # phase 8
        L.ho_compute_adjust(m.D, m.dnode.reshape(-1), st.force.reshape(-1), 3, ho.DISTRIBUTION)
        exchange(dist, rank, world, mesh.an_c, mesh.an_s, st.force, True)                   # phase 10
        L.ho_compute_displacement(m.N, m.nTable.reshape(-1), st.tm1.reshape(-1), st.tm2.reshape(-1), None,
                                  st.force.reshape(-1))
        exchange(dist, rank, world, mesh.an_s, mesh.an_c, st.tm2, False)                    # phase 13
        L.ho_compute_adjust(m.D, m.dnode.reshape(-1), st.tm2.reshape(-1), 3, ho.ASSIGNMENT)
        exchange(dist, rank, world, mesh.dn_s, mesh.dn_c, st.tm2, False)                    # phase 15
    res = [None] * world
    dist.all_gather_object(res, (gkey, st.tm2, info["owner"] == rank))
    if rank == 0:
        if basin:
            # the whole mesh must be the one the ranks cut: same coarsest-leaf limit as the multi-rank bootstrap
            smax = min(8, octree.bootstrap_size((nx, ny, 16), 32, world))
            whole, winfo = octree.octree_halfspace((nx, ny, 16), smax, H, DT, BASIN_MATS, basin_mat, 8.0, FREQ, vs_min=800.0)
        else:
            whole, winfo = meshgen.graded_halfspace(nx, ny, bands, h=H, dt=DT, freq=FREQ, layers=LAYERS)
        wm = ho.Mesh(whole.elem_lnid, whole.eTable, whole.nTable, whole.dnode, whole.edata, whole.K1, whole.K2)
        ws = ho.State(wm)
        wx, wy, wz = winfo["node_xyz"]
        wkey = (wx * (ny + 1) + wy) * (nz + 1) + wz
        wex, wey, wez = winfo["elem_xyz"]
        wes = winfo["elem_size"]
        we = int(np.nonzero((wex <= sx) & (sx < wex + wes) & (wey <= sy) & (sy < wey + wes) & (wez <= sz) & (sz < wez + wes))[0][0])
        wl = np.sort(whole.elem_lnid[we]).astype(np.int32)
        wo = np.argsort(whole.elem_lnid[we])
        for k in range(steps):
            ho.step(wm, ws, ho.RAYLEIGH, ho.EFFECTIVE, FREQ, DT, wl, F[k][wo])
        lut = {int(kk): i for i, kk in enumerate(wkey)}
        scale = np.abs(ws.tm2).max()
        assert scale > 0
        worst, owned_total = 0.0, 0
        for r, (gk, tm2, minemask) in enumerate(res):
            rows = np.array([lut[int(kk)] for kk in gk])
            err = np.abs(tm2 - ws.tm2[rows]).max() / scale
            worst = max(worst, err)
            owned_total += int(minemask.sum())
        assert owned_total == wkey.size, (owned_total, wkey.size)     # every node owned exactly once
        print(f"HALO_RESULT world={world} nodes={wkey.size} worst_rel_err={worst:.3e}", flush=True)
        assert worst < 1e-12, worst
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
