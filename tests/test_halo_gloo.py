"""CPU multi-process tests (gloo, world_size 2 and 4) of the host side of the multi-GPU path:
partitioned adaptive meshes, ownership, the four halo schedules and the hanging-node tables from
hercules_b200.meshgen drive a rank-per-process time loop (oracle arithmetic, gloo transport,
tests/halo_worker.py) whose result must equal the single-rank run on the whole mesh."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,bands", [(2, "2:1,3:2,2:4"), (4, "2:1,7:2"), (4, "4:1,2:2,2:4"),
                                         (3, "basin"), (4, "basin"),      # basin: general mesher + general partition
                                         (2, "basin-local"), (4, "basin-local")])   # per-rank meshing, counts over gloo
def test_partitioned_time_loop_equals_whole_mesh(world, bands):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "halo_worker.py"), "32", "32", bands, "12"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            o, _ = p.communicate(timeout=240)
            outs.append(o)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    assert any("HALO_RESULT" in o for o in outs)
