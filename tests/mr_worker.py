"""Worker for the multi-rank tests: one process per rank (launched by torch.distributed.run or by
tests/test_multirank_gpu.py).  Steps a multi-rank golden case of the unmodified reference through
libhercules_gpu.so with the halo exchange replacing schedule_senddata, and checks every rank's
tm1 snapshots against what the reference's own MPI ranks held.

usage: mr_worker.py <golden name> <transport: nccl|p2p> [flags]
env:   RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT (gloo rendezvous), MR_DEVICES = "0,1,..." device per rank
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch.distributed as dist
    import hercules_b200 as hb
    import refdump
    from conftest import load_golden, rank_view, rel_l2

    name, transport = sys.argv[1], sys.argv[2]
    flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    devs = [int(x) for x in os.environ.get("MR_DEVICES", ",".join(map(str, range(world)))).split(",")]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(name)
    assert int(g["nranks"][0]) == world
    v = rank_view(g, rank)
    P = refdump.params(v)
    s = hb.Solver(hb.HostMesh.from_dump(v), dt=P["dt"], dt2=P["dt2"], damping=P["damping"],
                  stiffness=P["stiffness"], freq=P["freq"], loaded_lnid=v["loaded_lnid"],
                  rank=rank, nranks=world, device=devs[rank], flags=flags)
    if transport == "nccl":
        uid = [hb.Solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s.comm_init(uid[0])
    else:
        blob = s.p2p_export()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        s.p2p_connect(blobs)
    dist.barrier()
    snaps = {int(k[len("tm1_step"):]): a for k, a in v.items() if k.startswith("tm1_step")}
    worst = 0.0
    F = v["forces"]
    for k in range(P["steps"]):
        s.step_begin(k)
        if k in snaps:
            got, ref = s.fetch_all(hb.TM1), snaps[k]
            if np.abs(ref).max() == 0:
                assert not got.any(), (rank, k)
            else:
                worst = max(worst, rel_l2(got, ref))
        s.compute_force_source(F[k] if v["loaded_lnid"].size else None)
        s.compute_force_stiffness()
        s.compute_force_damping()
        s.send_force_and_adjust()
        s.compute_displacement()
        s.send_displacement_and_adjust()
    s.sync()
    lay = s.layout()
    s.close()
    res = [None] * world
    dist.all_gather_object(res, (rank, worst, lay["early_tiles"], lay["ntiles"]))
    dist.barrier()
    if rank == 0:
        print("MR_RESULT", name, transport, " ".join(f"r{r}:{w:.2e}(early {e}/{n})" for r, w, e, n in res), flush=True)
        assert max(w for _, w, _, _ in res) < 1e-10
        assert max(w for _, w, _, _ in res) > 0
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
