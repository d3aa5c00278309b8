"""GPU parity tests of the multi-rank path: the reference's own multi-rank runs (golden fixtures
made with 2 and 4 MPI ranks of the unmodified reference on a mesh with hanging nodes) replayed
with one process per rank through libhercules_gpu.so, the halo exchange replacing
schedule_senddata (psolve.c:4945-5079).  Every rank's full displacement field must stay within
1e-10 relative L2 of the field the reference rank held (north_star's FP64 bar).

nccl transport needs one GPU per rank; the p2p transport (CUDA IPC mailboxes) also runs with all
ranks on one device, which is what a single-GPU box exercises.
"""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(name, world, transport, devices, flags=0, timeout=180):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), MR_DEVICES=",".join(map(str, devices)))
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "mr_worker.py"), name, transport,
                                       str(flags)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            o, _ = p.communicate(timeout=timeout)
            outs.append(o)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    assert any("MR_RESULT" in o for o in outs)


# 4 = HGPU_FLAG_NO_OVERLAP.  8 = HGPU_FLAG_TAIL_OVERLAP (opt-in, written after this round's GPU budget
# was spent: it has not run on hardware yet, so its cases only run when asked for)
FLAGS = [0, 4] + ([8] if os.environ.get("HGPU_TEST_TAIL_OVERLAP") == "1" else [])


# the basin cases: Morton blocks cutting through refinement levels of a laterally varying model -- hanging nodes
# whose anchors belong to another rank, nodes a rank harbors without having an element on them
CASES = [("graded3_rayleigh_eff_np2", 2), ("graded3_rayleigh_eff_np4", 4), ("uniform_rayleigh_eff_np3", 3),
         ("graded2_bkt_np2", 2), ("basin_rayleigh_eff_np2", 2), ("basin_rayleigh_eff_np3", 3), ("basin_rayleigh_eff_np4", 4),
         ("basin_bkt_np3", 3), ("graded3_rayleigh_eff_np8", 8)]


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("name,world", CASES)
def test_nccl_halo_matches_reference_ranks(name, world, flags):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs for NCCL (one rank per device)")
    _run(name, world, "nccl", list(range(world)), flags)


@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("name,world", CASES)
def test_p2p_halo_matches_reference_ranks(name, world, flags):
    """Peer-memory transport; ranks are spread over the GPUs present (all on one device on a
    single-GPU box: the mailboxes are then IPC mappings of the same device's memory)."""
    n = max(1, _ngpu())
    _run(name, world, "p2p", [r % n for r in range(world)], flags)
