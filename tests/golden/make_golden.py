#!/usr/bin/env python
"""make_golden.py -- regenerate tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/ref_dump (reference host code compiled from /root/reference + the dump hook of
oracle/ref_dump.c) on small synthetic cases and stores, per case and per rank:

  * the mesh and solver tables the hot path consumes (elem_lnid, eTable, nTable, dnode, edata,
    K1, K2, node ownership, halo schedules, loaded-node list)
  * the source history the reference wrote to force_process.<rank>
  * tm1 snapshots taken at the top of selected steps (after the swap, psolve.c:4271-4275)
  * the station rows the reference printed (7 significant digits)

plus an excerpt of the reference's shipped goldens for examples/simple (expected-out/).

Needs /root/reference (run `make -C oracle ref` first).  The .npz files are committed; tests and
the GPU box never need the reference.  Usage: python tests/golden/make_golden.py
"""
from __future__ import annotations

import bz2
import gzip
import shutil
import struct
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import refcase  # noqa: E402
import refdump  # noqa: E402

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference")

TWO_LAYER = dict(cvm_level=3, cvm_n=(8, 8, 4), vs_min=800, freq_hz=2.5,
                 layers=[(0, 3000, 1732, 2000), (125, 6000, 3464, 2700)])
THREE_LAYER = dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5,
                   layers=[(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)])
SRC = dict(src_xyz=(437.5, 562.5, 140.0), src_strike_dip_rake=(30.0, 70.0, 20.0),
           stations=[(500.0, 500.0, 0.0), (700.0, 300.0, 50.0), (120.0, 880.0, 300.0)])

CASES = {
    # name: (Case kwargs, ranks, snapshot period)
    "graded2_rayleigh_eff": (dict(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.06), 1, 20),
    "graded2_rayleigh_conv": (dict(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="conventional", end_t=0.06), 1, 20),
    "graded2_none_eff": (dict(**TWO_LAYER, **SRC, damping="none", stiffness="effective", end_t=0.06), 1, 20),
    "graded2_mass_eff": (dict(**TWO_LAYER, **SRC, damping="mass", stiffness="effective", end_t=0.06), 1, 20),
    "graded2_bkt": (dict(**TWO_LAYER, **SRC, damping="bkt", stiffness="effective", end_t=0.06), 1, 20),
    # soft column with Vp/Vs = 3 and finite Qk (use_infinite_qk = no): BOTH memory-variable families of
    # calc_conv / constant_Q_addforce are active in a whole run of the reference (damping.c:126-216, 256-371)
    "graded2_bkt_qk": (dict(cvm_level=3, cvm_n=(8, 8, 4), vs_min=800, freq_hz=2.5,
                            layers=[(0, 2400, 800, 2000), (125, 4800, 1600, 2300)], **SRC, damping="bkt",
                            stiffness="effective", end_t=0.06, use_infinite_qk="no"), 1, 20),
    "graded2_bkt_np2": (dict(**TWO_LAYER, **SRC, damping="bkt", stiffness="effective", end_t=0.06), 2, 20),
    "graded2_accel": (dict(**TWO_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.03,
                           print_accel="yes"), 1, 10),
    "graded3_rayleigh_eff": (dict(**THREE_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.1), 1, 25),
    "graded3_rayleigh_eff_np2": (dict(**THREE_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.1), 2, 25),
    "graded3_rayleigh_eff_np4": (dict(**THREE_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.1), 4, 25),
    # 8 ranks = one per GPU of an 8 x B200 box: what bench.py replays on the real devices before it times N = 8
    "graded3_rayleigh_eff_np8": (dict(**THREE_LAYER, **SRC, damping="rayleigh", stiffness="effective", end_t=0.1), 8, 50),
    # BASELINE.json configs[0] (examples/test1): homogeneous half-space 100 x 100 x 37.5 km (tick ratio 8:8:3),
    # meshed for 0.1 Hz (32 x 32 x 12 elements of 3125 m), quadratic point source 1 km deep, 500 steps of 0.02 s;
    # the unshipped labase.e is stood in for by a homogeneous etree of the same values (SURVEY.md 8d cfg 1)
    "test1_homogeneous": (dict(cvm_level=3, cvm_n=(8, 8, 3), east_m=100000.0, layers=[(0.0, 6000.0, 3464.0, 2700.0)],
                               freq_hz=0.1, ppw=8.0, vs_min=500.0, dt=0.02, end_t=10.0, damping="rayleigh",
                               stiffness="effective", src_xyz=(50000.0, 50000.0, 1000.0),
                               src_strike_dip_rake=(0.0, 90.0, 0.0), src_risetime=0.5, src_moment=1e15,
                               stations=[(50000.0, 50000.0, 0.0), (62000.0, 41000.0, 0.0), (30000.0, 70000.0, 5000.0)]), 1, 100),
    # laterally varying model: a soft box (Vs 866) in the middle of the top of a stiff half-space, so that octor
    # refines sideways too -- hanging nodes on x-, y- and z-faces and on edges of all three directions
    "basin_rayleigh_eff": (dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5, layers=[(0, 6000, 3464, 2700)],
                                basin=(375.0, 750.0, 250.0, 625.0, 125.0, 1800.0, 866.0, 1800.0), **SRC,
                                damping="rayleigh", stiffness="effective", end_t=0.06), 1, 20),
    # the soft box in a corner of the domain, reaching two lateral absorbing faces and the bottom: hanging nodes ON
    # domain faces and edges (node_setproperty's boundary branches, octor.c:3294-3800)
    "basin_corner_rayleigh_eff": (dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5, layers=[(0, 6000, 3464, 2700)],
                                       basin=(0.0, 250.0, 0.0, 375.0, 500.0, 1800.0, 866.0, 1800.0), **SRC,
                                       damping="rayleigh", stiffness="effective", end_t=0.04), 1, 20),
    # the laterally varying model on 2, 3 and 4 ranks: Morton blocks that cut through refinement levels, hanging
    # nodes whose anchors belong to other ranks (indirect sharing, octor.c:5800-6000)
    "basin_rayleigh_eff_np2": (dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5, layers=[(0, 6000, 3464, 2700)],
                                    basin=(375.0, 750.0, 250.0, 625.0, 125.0, 1800.0, 866.0, 1800.0), **SRC,
                                    damping="rayleigh", stiffness="effective", end_t=0.06), 2, 20),
    "basin_rayleigh_eff_np3": (dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5, layers=[(0, 6000, 3464, 2700)],
                                    basin=(375.0, 750.0, 250.0, 625.0, 125.0, 1800.0, 866.0, 1800.0), **SRC,
                                    damping="rayleigh", stiffness="effective", end_t=0.06), 3, 20),
    "basin_rayleigh_eff_np4": (dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5, layers=[(0, 6000, 3464, 2700)],
                                    basin=(375.0, 750.0, 250.0, 625.0, 125.0, 1800.0, 866.0, 1800.0), **SRC,
                                    damping="rayleigh", stiffness="effective", end_t=0.06), 4, 20),
    "basin_bkt_np3": (dict(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=2.5, layers=[(0, 4800, 1600, 2300)],
                           basin=(375.0, 750.0, 250.0, 625.0, 125.0, 2400.0, 800.0, 2000.0), **SRC,
                           damping="bkt", stiffness="effective", end_t=0.06, use_infinite_qk="no"), 3, 20),
    "uniform_rayleigh_eff": (dict(**SRC, damping="rayleigh", stiffness="effective", end_t=0.05), 1, 25),
    "uniform_rayleigh_eff_np3": (dict(**SRC, damping="rayleigh", stiffness="effective", end_t=0.05), 3, 25),
    "uniform_rayleigh_eff_np4": (dict(**SRC, damping="rayleigh", stiffness="effective", end_t=0.05), 4, 25),
}

KEEP = ["params", "counts", "domain_ticks", "elem_lnid", "elem_geid", "elem_level", "elem_edata",
        "eTable", "node_ticks", "node_gnid", "node_flags", "node_owner", "node_share", "nTable",
        "dnode", "K1", "K2", "dn_c_hdr", "dn_c_map", "dn_s_hdr", "dn_s_map", "an_c_hdr", "an_c_map",
        "an_s_hdr", "an_s_map", "loaded_lnid", "station_nodes", "station_local"]


def read_station(path: Path) -> np.ndarray:
    rows = [list(map(float, ln.split())) for ln in path.read_text().splitlines()
            if ln.strip() and not ln.lstrip().startswith("#")]
    return np.array(rows)


def make_case(name: str, kw: dict, nranks: int, every: int) -> None:
    c = refcase.Case(**kw)
    with tempfile.TemporaryDirectory() as td:
        d = refcase.write_case(c, td)
        refcase.run("ref_dump", d, nranks=nranks, env_extra={"HDUMP_EVERY": str(every)})
        out = {}
        for r in range(nranks):
            D = refdump.read_dump(d / f"dump.{r}.bin")
            pre = f"r{r}_" if nranks > 1 else ""
            for k in KEEP:
                out[pre + k] = D[k]
            for k, v in D.items():
                if k.startswith("tm1_step") or (k.startswith("tm2_step") and kw.get("print_accel") == "yes"):
                    out[pre + k] = v
            fp = d / "out" / "srctmp" / f"force_process.{r}"
            if D["loaded_lnid"].size and fp.exists():
                ll, F = refdump.read_force_process(fp, c.steps)
                assert np.array_equal(ll, D["loaded_lnid"])
                out[pre + "forces"] = F
            else:
                out[pre + "forces"] = np.zeros((c.steps, 0, 3))
        for i in range(len(c.stations)):
            out[f"station{i}"] = read_station(d / "out" / "stations" / f"station.{i}")
        out["nranks"] = np.array([nranks])
        np.savez_compressed(OUT / f"{name}.npz", **out)
        sz = (OUT / f"{name}.npz").stat().st_size
        print(f"{name}: ranks={nranks} E={int(out.get('counts', out.get('r0_counts'))[0])} -> {sz/1024:.0f} KiB")


def make_shipped_excerpt(nsteps: int = 1500) -> None:
    """examples/simple: the reference's own goldens (expected-out/) + the tables ref_dump gives for
    that exact case.  Forces come from the shipped force_process.0.gz, stations from the shipped
    station.N.bz2 (7 digits), so this fixture pins oracle and GPU path against files the
    reference's authors committed, not against anything rebuilt here."""
    ex = REF / "examples" / "simple"
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for sub in ("out/checkpoints", "out/planes", "out/srctmp", "out/stations"):
            (td / sub).mkdir(parents=True)
        shutil.copy(ex / "simple_case.e", td / "simple_case.e")
        shutil.copytree(ex / "in" / "sourcefiles", td / "in" / "sourcefiles")
        params = (ROOT / "BASELINE.md").read_text().split("```")[1]      # the merged parameters.in
        params = params.replace("simulation_end_time_sec        = 3", "simulation_end_time_sec        = 20")
        (td / "parameters.in").write_text(params)
        # the reference regenerates force_process.0 itself (FFT filter over all 20000 steps)
        refcase.run("ref_dump", td, nranks=1, env_extra={"HDUMP_EVERY": "0", "HDUMP_STOP_AFTER_INIT": "1"},
                    timeout=3600)
        D = refdump.read_dump(td / "dump.0.bin")
        out = {k: D[k] for k in KEEP}
        # shipped goldens
        raw = gzip.decompress((ex / "expected-out" / "srctmp" / "force_process.0.gz").read_bytes())
        n = struct.unpack_from("<i", raw, 0)[0]
        ll = np.frombuffer(raw, np.int32, n, 4)
        F = np.frombuffer(raw, np.float64, offset=4 + 4 * n).reshape(-1, n, 3)
        assert np.array_equal(ll, D["loaded_lnid"]), "loaded-node list differs from the shipped golden"
        _, Fnew = refdump.read_force_process(td / "out" / "srctmp" / "force_process.0")
        out["forces_rebuilt_max_rel_err"] = np.array([np.abs(Fnew - F).max() / np.abs(F).max()])
        out["forces"] = F[:nsteps].copy()
        for i in range(5):
            txt = bz2.decompress((ex / "expected-out" / "stations" / f"station.{i}.bz2").read_bytes()).decode()
            rows = [list(map(float, ln.split())) for ln in txt.splitlines()
                    if ln.strip() and not ln.lstrip().startswith("#")]
            out[f"station{i}"] = np.array(rows)[:nsteps]
        np.savez_compressed(OUT / "shipped_simple.npz", **out)
        print("shipped_simple:", (OUT / "shipped_simple.npz").stat().st_size // 1024, "KiB",
              "rebuilt-vs-shipped force rel err", out["forces_rebuilt_max_rel_err"][0])


if __name__ == "__main__":
    if not refcase.have_ref("ref_dump"):
        sys.exit("oracle/_ref/ref_dump missing: run `make -C oracle ref` (needs /root/reference)")
    only = set(sys.argv[1:])
    for name, (kw, nr, ev) in CASES.items():
        if not only or name in only:
            make_case(name, kw, nr, ev)
    if not only or "shipped_simple" in only:
        make_shipped_excerpt()
