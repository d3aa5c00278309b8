#!/usr/bin/env python
"""bench.py -- element-updates/s of the explicit time-stepping hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[1]): layered half-space (LOH.1 values: 1000 m of rho 2600 /
Vp 4000 / Vs 2000 over rho 2700 / Vp 6000 / Vs 3464), uniform octree mesh of 256^3 = 16 777 216
hexahedra per GPU (h = 25 m), Rayleigh damping, effective stiffness, FP64, point source on the 8
nodes of one element, 5 stations.  The mesh tables are produced on the host in the reference's
own layout (hercules_b200/meshgen.py, bit-exact with octor + solver_init on a reference-made
mesh, tests/test_meshgen.py).  N > 1 is weak scaling: the domain grows to N Morton-contiguous
256^3 blocks, one per GPU, as octor_partitiontree would cut it; shared nodes are exchanged
through the halo schedules.

A step = one time step over the whole mesh = E element-updates.

  value  device-resident: source history preloaded in HBM, K steps through hgpu_run, timed with
         CUDA events on the solver's stream, max over ranks.
  e2e    the per-step C-ABI sequence with HOST buffers: every step copies that step's source
         forces host->device (hgpu_force_source) and interpolates the stations on
         the device (hgpu_stations_record; rows read back to the host every 50 steps), as solver_run
         does with read_myForces and interpolate_station_displacements; the final displacement
         field is read back once at the end (inside the timed region).
  roofline  the fused tile kernel (element force + central-difference update): algorithmic bytes
         = 64 E + 248 N per launch over its mean CUDA-event duration.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/psolve_ref_O3, built from /root/reference
         with the fork-based mini-MPI of oracle/mpistub) on this box's host cores, on a smaller
         mesh of the same material/damping/stiffness configuration; the reference's own
         "TOTAL SOLVER" timer.  Falls back to the oracle's C port on one core if the binary did
         not travel.

--impl reference prints the same line for the reference arm alone (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

T_START = time.time()
METRIC = "element-updates/sec"
UNIT = "element-updates/s"
LAYERS = ((0.0, 4000.0, 2000.0, 2600.0), (1000.0, 6000.0, 3464.0, 2700.0))   # ztop, Vp, Vs, rho
# --damping bkt: a soft sedimentary column (Vs 500..1500) so that Qs AND Qk fall inside the BKT
# table for every element (psolve.c:7255-7310): both memory-variable families are advanced everywhere
LAYERS_BKT = ((0.0, 1500.0, 500.0, 2000.0), (500.0, 2200.0, 1000.0, 2200.0), (2000.0, 3000.0, 1500.0, 2400.0))
H_M, DT, FREQ = 25.0, 0.002, 1.0
BYTES_PER_ELEM = {"rayleigh": 64,       # 32 B ids + 32 B coefficients
                  "bkt": 1624}          # 32 ids + 16 c1,c2 + 40 BKT coefs + 768 state read + 768 state write
BYTES_PER_NODE = 248     # 72 B (tm1, tm2, force write) + 176 B update


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.nv = self.h = None
        self.lock = threading.Lock()
        try:                                        # NVML is brought up before the timed region starts
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[index]) if vis and vis.split(",")[0].isdigit() else index
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.nv = nv
        except Exception as e:                      # clocks are evidence, not a reason to fail
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def sample_now(self):
        """One sample from the calling thread (the bench calls it while the timed steps are in flight)."""
        nv = self.nv
        if nv is None:
            return
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            with self.lock:
                self.sm.append(sm)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
        except Exception as e:
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def run(self):
        while not self.stop_flag.is_set():
            self.sample_now()
            time.sleep(0.01)

    def result(self) -> dict:
        self.stop_flag.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---- the reference arm / cpu baseline -----------------------------------------------------------

REF_SAMPLE_EDGE = 64          # elements per edge of the reference arm's sample mesh


def reference_sample(steps: int, fill: int = 350, cores: int | None = None, repeats: int = 3) -> dict:
    """Run the unmodified reference on the host cores on a bounded sample of the workload.

    The reference skips elements whose nodes have not moved yet (vector_is_zero / the 1e-20
    early-outs, quake_util.c:36-96), so its speed depends on how far the wave has spread.  To
    time it in the state the GPU arm is timed in (every element active) the case is run for
    `fill` steps and for `fill + steps` steps (the wave from the central source crosses the 64^3
    sample in ~300 steps), and the rate is taken over the difference of the two TOTAL SOLVER
    timers.  Ranks are pinned to cpus (HMPI_PIN, oracle/mpistub); every run is repeated `repeats`
    times and the FASTEST run of each length is used (other tenants of the box only ever add time);
    the spread over the repeats is reported."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import refcase
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if refcase.have_ref("psolve_ref_O3") and refcase.have_ref("mkcvm"):
        # ranks = power of two <= min(cores, 32); the mini-MPI forks one process per rank
        want = min(cores or ncpu, 32)
        np_ = 1
        while np_ * 2 <= want:
            np_ *= 2
        n = REF_SAMPLE_EDGE                          # elements per edge of the sample mesh
        size = n * H_M
        # vs rule (quake_util.c:215-226): split while edge > Vs/(f ppw); f chosen so that both
        # layers stop at edge = H_M exactly (2000/(8 f) in (H_M, 2 H_M) and 3464/(8 f) < 2 H_M)
        f = 0.99 * 2000.0 / (8.0 * H_M)
        wall0 = time.time()
        runs = {fill: [], fill + steps: []}
        for rep in range(repeats):
            for nst in (fill, fill + steps):
                c = refcase.Case(cvm_level=4, cvm_n=(16, 16, 16), east_m=size, layers=[list(l) for l in LAYERS],
                                 freq_hz=f, ppw=8.0, vs_min=1900.0, dt=DT, end_t=DT * (nst + 0.5),
                                 damping="rayleigh", stiffness="effective", src_risetime=0.1,
                                 src_xyz=(size / 2 + H_M / 3, size / 2 + H_M / 3, size / 2 + H_M / 3),
                                 stations=[(size / 2, size / 2, 0.0)], station_rate=1)
                with tempfile.TemporaryDirectory() as td:
                    d = refcase.write_case(c, td)
                    out = refcase.run("psolve_ref_O3", d, nranks=np_, timeout=1500, env_extra={"HMPI_PIN": "1"})
                t = refcase.parse_timing(out)
                if not {"elements", "steps", "solver_s"} <= set(t):
                    raise RuntimeError("could not parse the reference's timing report:\n" + out[-2000:])
                runs[nst].append(t)
        wall = time.time() - wall0
        lo = min(runs[fill], key=lambda t: t["solver_s"])
        hi = min(runs[fill + steps], key=lambda t: t["solver_s"])
        dsteps = hi["steps"] - lo["steps"]
        dt_s = hi["solver_s"] - lo["solver_s"]
        if dsteps <= 0 or dt_s <= 0:
            raise RuntimeError(f"reference timing difference is not positive: {runs}")
        E = hi["elements"]
        # the same difference repeat by repeat: how far apart identical runs on this box are
        per_rep = [E * (b["steps"] - a["steps"]) / (b["solver_s"] - a["solver_s"])
                   for a, b in zip(runs[fill], runs[fill + steps]) if b["solver_s"] > a["solver_s"]]
        return {"value": E * dsteps / dt_s, "unit": UNIT, "cores": np_, "kind": "reference",
                "sample": f"psolve_ref_O3 (unmodified reference, gcc -O3 -march=x86-64-v3, {np_} pinned mini-MPI ranks on "
                          f"{ncpu} host cpus), uniform {n}^3 = {int(E)} elements, same layers/rayleigh/effective; "
                          f"{int(dsteps)} steps timed as the difference of its TOTAL SOLVER timer between a "
                          f"{int(lo['steps'])}-step run ({lo['solver_s']:.2f} s) and a {int(hi['steps'])}-step run "
                          f"({hi['solver_s']:.2f} s), i.e. after the wave has reached every element; fastest of "
                          f"{repeats} repeats of each (per-repeat rates {min(per_rep) / 1e6:.1f}-{max(per_rep) / 1e6:.1f} M/s; "
                          f"wall incl. meshing {wall:.0f} s)",
                "ms_per_step": 1e3 * dt_s / dsteps, "elements": int(E), "steps": int(dsteps),
                "min": min(per_rep) if per_rep else None, "max": max(per_rep) if per_rep else None}
    # fallback: the oracle's C restatement, one core
    import hercules_oracle as ho
    from hercules_b200 import meshgen
    n = 48
    mesh, info = meshgen.uniform_halfspace(n, n, n, h=H_M, dt=DT, freq=FREQ, layers=LAYERS)
    m = ho.Mesh(mesh.elem_lnid, mesh.eTable, mesh.nTable)
    st = ho.State(m)
    st.tm1[:] = np.random.default_rng(0).standard_normal(st.tm1.shape)
    t0 = time.time()
    for _ in range(steps):
        ho.step(m, st, ho.RAYLEIGH, ho.EFFECTIVE, FREQ, DT)
    el = time.time() - t0
    return {"value": m.E * steps / el, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle C port (gcc -O2), 1 core, uniform {n}^3 elements, {steps} steps",
            "ms_per_step": 1e3 * el / steps, "elements": m.E, "steps": steps}


def reference_config(r: dict, gpus: int) -> dict:
    """What the reference arm actually ran: NOT the B200 arm's 256^3 per GPU -- a bounded sample of it."""
    E = r.get("elements")
    return {"workload": f"SAMPLE of configs[1] (not the B200 arm's mesh): layered half-space (LOH.1 values), uniform octree mesh "
                        f"of {E} elements in total (h={H_M:g} m), rayleigh damping, effective stiffness, point source, 1 station; "
                        f"the unmodified CPU reference on this box's host cores.  The same sample is timed whatever --gpus says: "
                        f"a ratio against the N-GPU line compares N x 256^3 elements on N GPUs with {E} elements on the host",
            "elements_per_gpu": None, "global_elements": E, "dt": DT, "sample_of": "configs[1]",
            "same_config_as_b200_arm": False, "n_gpus_ignored": gpus}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = reference_sample(max(args.steps, 300))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": r.get("steps", args.steps), "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": reference_config(r, args.gpus),
            "cpu_baseline": {k: r.get(k) for k in ("value", "unit", "cores", "kind", "sample", "min", "max")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- the B200 arm -----------------------------------------------------------------------------------

def block_grid(world: int) -> tuple[int, int, int]:
    """Blocks per axis (x, y, z) of the weak-scaling domain: Morton order fills x, then y, then z."""
    bx = by = bz = 1
    k = 0
    while bx * by * bz < world:
        if k % 3 == 0:
            bz *= 2          # the most significant Morton bit is z
        elif k % 3 == 1:
            by *= 2
        else:
            bx *= 2
        k += 1
    return bx, by, bz


def adaptive_bands(n: int):
    """configs[2]: three octree levels by depth on an n x n h-grid, n h deep: 11/16 of the depth in
    elements of edge h, 4/16 in 2h, 1/16 in 4h (n = 512: 92.3 M + 4.2 M + 0.13 M = 96.6 M elements)."""
    return ((n * 11 // 16, 1), (n // 8, 2), (n // 64, 4))


def column_grid(world: int) -> tuple[int, int]:
    """Columns per axis (x, y) of the adaptive weak-scaling domain: Morton order fills x, then y."""
    cx = cy = 1
    while cx * cy < world:
        if cx == cy:
            cx *= 2
        else:
            cy *= 2
    if cx * cy != world:
        raise SystemExit("--workload adaptive needs a power-of-two number of GPUs")
    return cx, cy


BASIN_MATS = ((1500.0, 500.0, 2000.0), (3000.0, 1000.0, 2200.0), (5000.0, 2000.0, 2500.0), (6928.0, 4000.0, 2800.0))


def basin_workload(n: int, damping: int, part=None, local: bool = False, allgather=None, threads: int = 1):
    """configs[4] at single-GPU scale: the terashake box (300 x 600 x 84.375 km, tick ratio 32:64:9,
    examples/terashake/physics.in) with a synthetic CVM-like model -- Vs 500 / 1000 / 2000 / 4000 m/s:
    sediment basins (Gaussian blobs, fixed seed 20240901) over a depth-layered crust, piecewise constant
    on cells of 4 h like a material etree -- meshed by hercules_b200.octree exactly as octor would
    (vs rule, 2:1 balance, hanging nodes in every orientation): four octree levels.
    n = h-cells along the 300 km edge.  part = (rank, world): that rank's Morton block of the ONE mesh;
    local: built from the rank's own neighbourhood by hercules_b200.octree_local (native primitives, leaf counts
    per coarse cell all-gathered through `allgather`) instead of meshing the whole domain on every rank and
    cutting it -- the two give identical tables (tests/test_octree_local.py).  Returns (mesh, info, dt, fmax, h)."""
    from hercules_b200 import octree
    dims = (n, 2 * n, 9 * n // 32)
    if 9 * n % 32 or dims[2] % 4:
        raise SystemExit("--workload basin needs --edge to be a multiple of 128")
    g = int(np.gcd.reduce(dims))
    smax = min(g & -g, 8)
    h = 300000.0 / n
    ppw = 8.0
    fmax = 499.0 / (ppw * h)                       # an octant of edge s h is split while s * 499 > Vs
    dt = 0.2 * h / 1500.0
    rng = np.random.default_rng(20240901)
    blobs = [(rng.uniform(0.1, 0.9) * dims[0], rng.uniform(0.1, 0.9) * dims[1], rng.uniform(0.08, 0.25) * dims[0],
              rng.uniform(0.05, 0.2) * dims[2]) for _ in range(7)]

    def mat_of(x, y, z):
        x, y, z = (np.floor(np.asarray(v) / 4) * 4 + 2 for v in (x, y, z))      # the model's own grid: cells of 4 h
        depth = np.zeros(np.shape(z))
        for bx_, by_, r, dz in blobs:
            depth = np.maximum(depth, dz * np.exp(-((x - bx_) ** 2 + (y - by_) ** 2) / r ** 2))
        return np.where(z < depth, 0, np.where(z < 2 * depth + 0.06 * dims[2], 1,
                        np.where(z < 0.45 * dims[2], 2, 3))).astype(np.int64)
    if local and part is not None:
        from hercules_b200 import octree_local
        mesh, info = octree_local.octree_halfspace_local(dims, smax, h, dt, list(BASIN_MATS), mat_of, ppw, fmax, part[0], part[1],
                                                         damping=damping, allgather=allgather, threads=threads, model_cell=4)
        info.pop("model", None)
    elif part is None or part[1] == 1:
        mesh, info = octree.octree_halfspace(dims, smax, h, dt, list(BASIN_MATS), mat_of, ppw, fmax, damping=damping)
    else:
        # every rank builds the whole mesh and takes its Morton block (octree.partition); fine up to a few 10 M elements
        mesh, info = octree.octree_halfspace_part(dims, smax, h, dt, list(BASIN_MATS), mat_of, ppw, fmax, part[0], part[1],
                                                  damping=damping)
    info["mat_of"] = mat_of
    return mesh, info, dt, fmax, h


def basin_config(n, info, damping, dt, fmax, h) -> dict:
    sizes, counts = np.unique(info["elem_size"], return_counts=True)
    return {"workload": f"configs[4] at single-GPU scale: terashake box 300 x 600 x 84.375 km (32:64:9), synthetic CVM-like model "
                        f"(sediment basins, seed 20240901, Vs 500/1000/2000/4000 m/s on cells of 4 h, h = {h:g} m), meshed as octor "
                        f"would (vs rule at {fmax:.4g} Hz, 8 points per wavelength, 2:1 balance): {info['E']} elements on "
                        f"{len(sizes)} octree levels, {info['N']} nodes, {info['D']} hanging nodes on faces and edges of every "
                        f"orientation; {damping} damping, effective stiffness, point source, 5 stations",
            "elements_per_gpu": info["E"], "global_elements": info.get("etotal", info["E"]), "hanging_nodes": info["D"],
            "elements_by_size": {int(a): int(b) for a, b in zip(sizes, counts)}, "global_grid": list(info["dims"]), "dt": dt,
            "partition": ("single rank (meshed by hercules_b200.octree_local, native primitives)" if info.get("nranks", 1) == 1 else
                          f"{info['nranks']} blocks of the Morton-ordered leaf list (as octor_partitiontree cuts it), every rank "
                          f"meshing only its block and a one-cell ring (octree_local: {info.get('local_region_elements', 0)} "
                          "leaves built on this rank); ONE mesh over all GPUs: strong scaling"),
            "l2": "inputs larger than L2 for --edge >= 768; no explicit flush"}


def containing_element(info: dict, x: float, y: float, z: float, missing_ok: bool = False) -> int:
    """Index of the (local) leaf that holds the point (units of h); -1 if it is on another rank and missing_ok."""
    ex, ey, ez = info["elem_xyz"]
    es = info["elem_size"]
    hit = np.nonzero((ex <= x) & (x < ex + es) & (ey <= y) & (y < ey + es) & (ez <= z) & (z < ez + es))[0]
    if hit.size == 0 and missing_ok:
        return -1
    if hit.size != 1:
        raise ValueError(f"point ({x},{y},{z}) is not in exactly one element")
    return int(hit[0])


def adaptive_layers(n: int):
    """Vs doubles from band to band, which is what makes octor refine by one level per band."""
    z1, z2 = (n * 11 // 16) * H_M, (n * 11 // 16 + n // 4) * H_M
    return ((0.0, 1800.0, 1000.0, 2200.0), (z1, 3600.0, 2000.0, 2500.0), (z2, 6000.0, 3464.0, 2700.0))


def workload_config(world: int, n: int, halo: str = "p2p", damping: str = "rayleigh", info: dict | None = None) -> dict:
    if info is not None and "bands" in info:
        cx, cy = column_grid(info.get("columns", world))
        return {"workload": f"configs[{3 if world > 1 or info.get('strong') else 2}]: adaptive octree mesh, 3 refinement levels (element edge "
                            f"{H_M:g}/{2*H_M:g}/{4*H_M:g} m by depth band, "
                            f"{'Vs 1000/2000/3464 m/s' if damping == 'rayleigh' else 'soft sedimentary column Vs 500-1500 m/s (configs[4] damping model)'}"
                            f"), {info['E']} elements, {info['N']} "
                            f"nodes, {info['D']} owned hanging (dangling) nodes per GPU on the two 2:1 interfaces, {damping} damping, "
                            "effective stiffness, point source, 5 stations per GPU; mesh tables in octor's layout from "
                            "meshgen.graded_halfspace (bit-exact with the reference's mesher and partition on "
                            "tests/golden/graded{2,3}_*.npz)",
                "elements_per_gpu": info["E"], "global_elements": info["etotal"], "hanging_nodes": info["D"],
                "global_grid": [n * cx, n * cy, n], "bands": [list(b) for b in info["bands"]], "dt": DT,
                "partition": ((f"{world} columns" if not info.get("strong") else f"{info.get('columns')} columns cut into {world} blocks") +
                              f" of {n}^3 h-cells = octor's equal blocks of the Morton-ordered leaf list, halo "
                              f"exchange over {halo} overlapped with interior tiles" if world > 1 else "single rank"),
                "l2": "inputs larger than L2; no explicit flush"}
    bx, by, bz = block_grid(world)
    what = ("rayleigh damping, effective stiffness" if damping == "rayleigh" else
            "BKT damping (variant of configs[1] with configs[4]'s damping model on a soft sedimentary column, "
            "Vs 500-1500 m/s; Qs, Qk from Vs as mesh_correct_properties derives them)")
    return {"workload": f"configs[1]: layered half-space{' (LOH.1 values)' if damping == 'rayleigh' else ''}, uniform octree mesh {n}^3 elements "
                        f"per GPU (h={H_M:g} m), {what}, point source, 5 stations",
            "elements_per_gpu": n ** 3, "global_elements": n ** 3 * world,
            "global_grid": [n * bx, n * by, n * bz], "dt": DT,
            "partition": f"{world} Morton-contiguous block(s) (octor_partitiontree rule), halo exchange over "
                         f"{halo} overlapped with interior tiles" if world > 1 else "single rank",
            "l2": "inputs larger than L2 (node arrays >= 400 MB each step); no explicit flush"}


def run_graded(args) -> None:
    """--workload graded: configs[2] at reduced size.  The unmodified reference (integration/_bin/
    psolve_gpu: its own main, octor mesher, solver_init, source and station writers) meshes a
    three-layer model whose Vs bands give three consecutive octree levels with hanging nodes on
    both interfaces, and libhercules_gpu.so executes its time loop.  The size is bounded by the
    reference's single-rank host mesher (~20 us per element), not by the GPU."""
    import re
    import subprocess
    sys.path.insert(0, str(ROOT / "oracle"))
    import refcase
    gpu_bin = ROOT / "integration" / "_bin" / "psolve_gpu"
    if not (gpu_bin.exists() and refcase.have_ref("mkcvm")):
        raise SystemExit("--workload graded needs integration/_bin/psolve_gpu and oracle/_ref/mkcvm "
                         "(built where /root/reference exists)")
    steps = args.steps + args.warmup
    f = args.graded_freq
    dt = 0.0025 / f                      # h_min / Vp_max with margin: h = 1000 / 2^ceil(log2(f 8 1000 / 866))
    c = refcase.Case(cvm_level=4, cvm_n=(16, 16, 8), vs_min=800, freq_hz=f,
                     layers=[(0, 1800, 866, 1800), (62.5, 3000, 1732, 2000), (250, 6000, 3464, 2700)],
                     src_xyz=(437.5, 562.5, 140.0), src_strike_dip_rake=(30.0, 70.0, 20.0), src_risetime=20 * dt,
                     stations=[(500.0, 500.0, 0.0), (700.0, 300.0, 50.0), (120.0, 880.0, 300.0)], station_rate=10,
                     damping="rayleigh", stiffness="effective", dt=dt, end_t=dt * (steps + 0.5))
    t0 = time.time()
    with tempfile.TemporaryDirectory() as td:
        d = refcase.write_case(c, td)
        p = subprocess.run([str(gpu_bin), "parameters.in"], cwd=d,
                           env=dict(os.environ, HMPI_NP="1", PSOLVE_GPU_WARMUP=str(args.warmup)),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
        if p.returncode != 0:
            raise SystemExit("psolve_gpu failed:\n" + p.stdout[-3000:])
        mon = (d / "out" / "monitor.txt").read_text()
        st0 = (d / "out" / "stations" / "station.0").read_text().splitlines()
    wall = time.time() - t0
    t = refcase.parse_timing(p.stdout)
    m = re.search(r"gpu_solver_run\(\) done: (\d+) steps, (\d+) kernel launches, loop wall ([0-9.eE+-]+) s; device time: "
                  r"step kernels ([0-9.eE+-]+) s, new displacement ([0-9.eE+-]+) s, adjust ([0-9.eE+-]+) s, "
                  r"exchanges ([0-9.eE+-]+) s", mon)
    if not m or "elements" not in t:
        raise SystemExit("could not parse psolve_gpu's report:\n" + mon[-2000:])
    hm = re.search(r"host time: taps ([0-9.eE+-]+) s, reference I/O block ([0-9.eE+-]+) s, hgpu calls ([0-9.eE+-]+) s", mon)
    nst, launches = int(m.group(1)), int(m.group(2))
    loop_s, k_s, nd_s, adj_s = float(m.group(3)), float(m.group(4)), float(m.group(5)), float(m.group(6))
    E, N = int(t["elements"]), int(t["nodes"])
    dangling = int(re.search(r"Total dangling nodes:\s*(\d+)", p.stdout).group(1))
    last = [float(x) for x in st0[-1].split()]
    if not all(np.isfinite(last)):
        raise SystemExit("non-finite station values")
    peak, peak_src = peaks()
    alg = BYTES_PER_ELEM["rayleigh"] * E + BYTES_PER_NODE * N
    line = {"metric": METRIC, "value": E * nst / loop_s, "unit": UNIT, "n_gpus": 1, "steps": nst, "warmup": args.warmup,
            "ms_per_step": 1e3 * loop_s / nst, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[2] at reduced size: adaptive octree mesh, 3 refinement levels, {E} elements, "
                                   f"{N} nodes, {dangling} hanging nodes (three-layer model meshed for {f:g} Hz by the "
                                   "reference's own octor on the host), rayleigh damping, effective stiffness, point "
                                   "source, 3 stations every 10 steps; the reference's main() with its time loop on "
                                   "libhercules_gpu.so (integration/psolve_gpu.c)",
                       "elements_per_gpu": E, "dt": dt,
                       "l2": "node arrays smaller than L2 at this size: not an HBM-roofline run" if N * 24 < 60e6 else
                             "inputs larger than L2"},
            "e2e": {"value": E * nst / loop_s, "unit": UNIT, "h2d_bytes_per_step": 192, "d2h_bytes_per_step": 58,
                    "what": "value IS end to end here: host source rows in every step, station rows out every 10 steps"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "step_kernel<1,false,256>", "achieved": alg * nst / k_s / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg * nst / k_s / 1e9 / peak, "peak_source": peak_src,
                         "kernel_ms": 1e3 * k_s / nst, "kernel_share_of_step": k_s / loop_s, "traffic": None},
            "phases_ms_per_step": {"step_kernels": 1e3 * k_s / nst, "new_disp": 1e3 * nd_s / nst, "adjust": 1e3 * adj_s / nst},
            "host_ms_per_step": ({"station_taps": 1e3 * float(hm.group(1)) / nst, "reference_io_block": 1e3 * float(hm.group(2)) / nst,
                                  "hgpu_calls": 1e3 * float(hm.group(3)) / nst} if hm else None),
            "cpu_baseline": None, "setup_s": {"reference_host_code_and_meshing": round(wall - loop_s, 1)}}
    print(json.dumps(line), flush=True)


PARITY_GOLDEN = {2: "graded3_rayleigh_eff_np2", 3: "uniform_rayleigh_eff_np3", 4: "graded3_rayleigh_eff_np4",
                 8: "graded3_rayleigh_eff_np8"}


def parity_check(hb, dist, rank: int, world: int, local: int, halo: str, flags: int) -> dict | None:
    """N > 1: before anything is timed, replay a multi-rank run of the UNMODIFIED reference (tests/golden:
    tables and tm1 snapshots of every MPI rank, 3-level mesh with hanging nodes on rank boundaries) on the N
    real GPUs through the same halo transport and the same flags as the timed run, one reference rank per
    GPU; every rank's displacement field must stay within 1e-10 relative L2 of what the reference rank held
    (schedule_senddata semantics, psolve.c:4945-5079).  Returns the worst error over ranks and snapshots."""
    name = PARITY_GOLDEN.get(world)
    path = ROOT / "tests" / "golden" / f"{name}.npz" if name else None
    if not name or not path.exists():
        return None
    sys.path.insert(0, str(ROOT / "oracle"))
    import refdump
    z = np.load(path)
    pre = f"r{rank}_"
    v = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    P = refdump.params(v)
    s = hb.Solver(hb.HostMesh.from_dump(v), dt=P["dt"], dt2=P["dt2"], damping=P["damping"], stiffness=P["stiffness"],
                  freq=P["freq"], loaded_lnid=v["loaded_lnid"], rank=rank, nranks=world, device=local, flags=flags)
    if halo == "nccl":
        uid = [hb.Solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s.comm_init(uid[0])
    else:
        blobs = [None] * world
        dist.all_gather_object(blobs, s.p2p_export())
        s.p2p_connect(blobs)
    dist.barrier()
    snaps = {int(k[len("tm1_step"):]): a for k, a in v.items() if k.startswith("tm1_step")}
    worst, nsnap = 0.0, 0
    F = v["forces"]
    for k in range(P["steps"]):
        s.step_begin(k)
        if k in snaps and np.abs(snaps[k]).max() > 0:
            got = s.fetch_all(hb.TM1)
            worst = max(worst, float(np.linalg.norm(got - snaps[k]) / np.linalg.norm(snaps[k])))
            nsnap += 1
        s.compute_force_source(F[k] if v["loaded_lnid"].size else None)
        s.compute_force_stiffness()
        s.compute_force_damping()
        s.send_force_and_adjust()
        s.compute_displacement()
        s.send_displacement_and_adjust()
    s.sync()
    s.close()
    res = [None] * world
    dist.all_gather_object(res, (worst, nsnap))
    out = {"case": name, "what": "every rank's tm1 vs the snapshots of the reference's own MPI ranks (unmodified reference, "
                                 f"{world} ranks), through the timed run's transport and flags, one rank per GPU",
           "rel_l2": max(w for w, _ in res), "ranks": world, "snapshots_compared": int(sum(n for _, n in res)),
           "steps": int(P["steps"]), "tolerance": 1e-10, "transport": halo}
    if not (0 < out["rel_l2"] < 1e-10):
        raise SystemExit(f"parity check failed before timing: {out}")
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--edge", dest="n", type=int, default=None,
                    help="elements per edge per GPU (default 256; basin: h-cells along the 300 km side, default 768)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="N > 1: skip the replay of the reference's N-rank golden on the real GPUs before timing")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: exchange after all tiles, one stream")
    ap.add_argument("--tail-overlap", action="store_true",
                    help="multi-GPU, opt-in: shared-node update + displacement exchange beside the late tiles "
                         "(HGPU_FLAG_TAIL_OVERLAP)")
    ap.add_argument("--strong", action="store_true",
                    help="--workload adaptive: cut ONE mesh of 8 columns (4 x 2; --edge 384: 326 M elements, configs[3]'s ~400 M) "
                         "over the GPUs instead of one column per GPU")
    ap.add_argument("--columns", type=int, default=8, choices=[2, 4, 8],
                    help="--strong: columns of the one mesh (--edge 512: 4 columns = 386 M elements, configs[3]'s ~400 M; "
                         "8 columns = 773 M); must be a multiple of --gpus")
    ap.add_argument("--wpass", action="store_true",
                    help="opt-in step-kernel variant (HGPU_FLAG_WPASS): damped displacement formed once per staged node")
    ap.add_argument("--tile-nodes", type=int, default=0)
    ap.add_argument("--damping", default="rayleigh", choices=["rayleigh", "bkt"],
                    help="rayleigh = the headline workload (configs[1]); bkt = the same mesh with BKT damping")
    ap.add_argument("--stiffness", default="effective", choices=["effective", "conventional"],
                    help="conventional: a solver created with HGPU_STIFFNESS_CONVENTIONAL (stiffness.c:121-176); it applies "
                         "the same operator in factored form unless --dense-k is given (Rayleigh damping only)")
    ap.add_argument("--dense-k", action="store_true",
                    help="with --stiffness conventional: HGPU_FLAG_DENSE_K, the literal dense 24 x 24 K1 / K2 products "
                         "(the DENSE variant of the step kernel; profiles/r02k_bench_conventional_160.json)")
    ap.add_argument("--workload", default="uniform", choices=["uniform", "adaptive", "graded", "basin"],
                    help="uniform = configs[1] (the headline); adaptive = configs[2]: 3-level octree mesh with hanging "
                         "nodes, ~100 M elements at --edge 512 (meshgen.graded_halfspace, single GPU); graded = configs[2] "
                         "at reduced size through the reference's own main and mesher (integration/psolve_gpu); basin = configs[4] at "
                         "single-GPU scale: terashake box, synthetic CVM-like basins, four octree levels, BKT by default "
                         "(hercules_b200.octree; --edge = h-cells along the 300 km side, multiple of 128)")
    ap.add_argument("--graded-freq", type=float, default=20.0, help="--workload graded: meshing frequency (Hz)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"],
                    help="halo transport: peer-memory mailboxes over NVLink (default) or NCCL send/recv")
    args = ap.parse_args()
    if args.n is None:
        args.n = 768 if args.workload == "basin" else 256
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.workload == "graded":
        run_graded(args)
        return

    import torch
    import torch.distributed as dist
    import hercules_b200 as hb
    from hercules_b200 import meshgen

    wd = float(os.environ.get("BENCH_WATCHDOG_S", "0"))
    if wd > 0:                                  # a hung run says where it hangs, then ends
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True, file=sys.stderr)

    def note(msg: str) -> None:
        print(f"[bench rank {os.environ.get('RANK', '0')} +{time.time() - T_START:.1f}s] {msg}", file=sys.stderr, flush=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version
    # banner, for one) is sent to stderr by pointing fd 1 there until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not hb.SO.exists():
        hb.build()
    if args.dense_k and args.stiffness != "conventional":
        raise SystemExit("--dense-k goes with --stiffness conventional")
    run_flags = (hb.FLAG_NO_OVERLAP if args.no_overlap else 0) | (hb.FLAG_TAIL_OVERLAP if args.tail_overlap else 0) | \
                (hb.FLAG_WPASS if args.wpass else 0) | (hb.FLAG_DENSE_K if args.dense_k else 0)
    note("process group up")
    parity = parity_check(hb, dist, rank, world, local, args.halo, run_flags) if world > 1 and not args.no_parity_check else None
    note(f"parity check done: {parity and parity['rel_l2']}")

    # ---- workload -------------------------------------------------------------------------------
    n = args.n
    bx, by, bz = block_grid(world)
    t0 = time.time()
    damp = hb.BKT if args.damping == "bkt" else hb.RAYLEIGH
    layers = LAYERS_BKT if args.damping == "bkt" else LAYERS
    adaptive = args.workload == "adaptive"
    basin = args.workload == "basin"
    dt_run, freq_run, h_run = DT, FREQ, H_M
    if basin:
        def gather_counts(mine):                                  # one int per coarse cell of this rank's share
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            return parts
        mesh, info, dt_run, freq_run, h_run = basin_workload(n, damp, (rank, world), local=True,
                                                             allgather=gather_counts if world > 1 else None,
                                                             threads=max(1, (os.cpu_count() or 1) // world))
    elif adaptive:
        if n % 64:
            raise SystemExit("--workload adaptive needs --edge to be a multiple of 64")
        bands = adaptive_bands(n)
        if args.strong and (args.columns % world or n & (n - 1)):
            raise SystemExit("--strong: --columns must be a multiple of --gpus and --edge a power of two (octor's Morton "
                             "blocks are whole columns only then)")
        cols = column_grid(args.columns if args.strong else world)
        try:                                        # host side: ~350 B per element while the mesh tables and the
            import psutil                           # tile plan are built, on every rank of this box at once
            need = 350 * sum(nl * (n // sz) ** 2 for nl, sz in bands) * (args.columns if args.strong else world)
            if psutil.virtual_memory().available < need:
                raise SystemExit(f"--workload adaptive --edge {n} --gpus {world}: needs ~{need >> 30} GiB of host memory for "
                                 f"the mesh tables of {world} rank(s); use a smaller --edge")
        except ImportError:
            pass
        layers = adaptive_layers(n) if args.damping == "rayleigh" else LAYERS_BKT
        # configs[3] (N > 1): the same column repeated side by side, one per GPU -- octor's equal-count blocks
        # of the Morton-ordered leaf list are then exactly the columns (meshgen.column_regions)
        mesh, info = meshgen.graded_halfspace(n * cols[0], n * cols[1], bands, h=H_M, dt=DT, freq=FREQ, layers=layers,
                                              damping=damp, part=(rank, world) if world > 1 else None)
        info["strong"] = bool(args.strong)
        info["columns"] = args.columns if args.strong else world
    elif world == 1:
        mesh, info = meshgen.uniform_halfspace(n, n, n, h=H_M, dt=DT, freq=FREQ, layers=layers, damping=damp)
    else:
        mesh, info = meshgen.uniform_halfspace(n * bx, n * by, n * bz, h=H_M, dt=DT, freq=FREQ,
                                               layers=layers, part=(rank, world), damping=damp)
    E, N = info["E"], info["N"]
    strong = (basin and world > 1) or (adaptive and args.strong)       # ONE mesh cut over the GPUs
    e_global = info["etotal"] if strong else E * world
    # point source: the 8 nodes of the element at the centre of this rank's block, 2000 m deep on
    # rank 0 (other ranks carry no source, as in a real run where one rank holds the hypocentre)
    steps_hist = max(args.steps, args.warmup)
    if rank == 0 or basin:                  # basin: the rank that holds the hypocentre's element
        zsrc = min(n - 1, int(2000 / H_M)) if not adaptive else min(int(2000 / H_M), adaptive_bands(n)[0][0] - 1)
        if basin:
            ce = containing_element(info, info["dims"][0] * 0.5, info["dims"][1] * 0.5, info["dims"][2] * 0.2, missing_ok=world > 1)
        else:
            ce = meshgen.element_index(info, n // 2, n // 2, zsrc)   # adaptive: inside the band of finest elements
        loaded = np.sort(mesh.elem_lnid[ce]).astype(np.int32) if ce >= 0 else np.zeros(0, np.int32)
        tt = (np.arange(steps_hist) + 1) * dt_run
        ramp = np.minimum(1.0, (tt / (50 * dt_run)) ** 2)[:, None, None]
        rng = np.random.default_rng(11)
        F_all = np.ascontiguousarray(ramp * 1e9 * rng.standard_normal((1, 8, 3)))
        if not loaded.size:
            F_all = np.zeros((steps_hist, 0, 3))
    else:
        loaded, F_all = np.zeros(0, np.int32), np.zeros((steps_hist, 0, 3))
    # 5 stations x 8 nodes on this rank's surface
    st_elems = [containing_element(info, info["dims"][0] * fx, info["dims"][1] * fy, 0.0, missing_ok=world > 1) if basin else
                meshgen.element_index(info, int(n * fx), int(n * fy), 0)
                for fx, fy in ((.5, .5), (.6, .6), (.7, .7), (.8, .8), (.9, .9))]
    st_elems = [e for e in st_elems if e >= 0] or [0]
    st_nodes = np.ascontiguousarray(mesh.elem_lnid[st_elems].reshape(-1), np.int32)
    t_mesh = time.time() - t0

    t0 = time.time()
    if args.stiffness == "conventional" and args.damping != "rayleigh":
        raise SystemExit("--stiffness conventional is measured with Rayleigh damping only")
    s = hb.Solver(mesh, dt=dt_run, damping=damp, stiffness=hb.CONVENTIONAL if args.stiffness == "conventional" else hb.EFFECTIVE,
                  freq=freq_run,
                  loaded_lnid=loaded, rank=rank, nranks=world, device=local,
                  tile_nodes=args.tile_nodes, flags=hb.FLAG_TIMERS | run_flags)
    if world > 1:
        if args.halo == "nccl":
            uid = [hb.Solver.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            s.comm_init(uid[0])
        else:
            blobs = [None] * world
            dist.all_gather_object(blobs, s.p2p_export())
            s.p2p_connect(blobs)
    t_init = time.time() - t0
    note(f"solver ready: mesh {t_mesh:.1f} s, init {t_init:.1f} s")
    layout = s.layout()
    stream = torch.cuda.ExternalStream(s.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s.sync()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident run: `value` ------------------------------------------------------------
    s.source_preload(0, F_all[:steps_hist])
    s.run(0, args.warmup)
    note("warm-up enqueued")
    barrier()
    note("warm-up done")
    tm0 = s.timers()
    clk = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clk.start()
    ev0.record(stream)
    s.run(0, args.steps)
    ev1.record(stream)
    clk.sample_now()                       # the steps are still in flight here
    barrier()
    note("timed run done")
    dev_s = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    clocks = clk.result()
    tm1 = s.timers()
    launches = int(tm1["launches"] - tm0["launches"])
    fused_s = tm1["fused_step"] - tm0["fused_step"]
    # device time of every phase per step (CUDA events, rank 0), under the reference's timer names
    phases_ms = {k: round(1e3 * (tm1[k] - tm0[k]) / args.steps, 4) for k in tm1
                 if k not in ("launches", "steps") and tm1[k] != tm0[k]}
    tile_launches = args.steps
    probe = s.fetch_nodes(hb.TM2, st_nodes)
    if not np.isfinite(probe).all():
        raise SystemExit("non-finite displacements after the timed run")

    # ---- end-to-end run through the per-step ABI with host buffers: `e2e` ------------------------
    e2e = None
    if not args.no_e2e:
        # host buffers from the library's page-locked allocator (what a psolve.c integration would
        # use for tm1 and the station rows): source rows in, station rows out every step, the whole
        # displacement field out once at the end -- all inside the timed region
        F_host = hb.PinnedArray(F_all.shape); F_host.a[...] = F_all
        n_st = st_nodes.size // 8
        RING = 50                                   # station rows are read back every RING steps
        st_out = hb.PinnedArray((RING, n_st, 9))
        final = hb.PinnedArray((N, 3))
        # stations interpolated on the device (station_kernel = interpolate_station_displacements,
        # psolve.c:6680-6795) into a ring the host drains every RING steps: the per-step result still
        # crosses to the host inside the timed region, but a station step no longer drains the pipeline
        s.stations_attach(st_nodes, np.zeros((n_st, 3)), capacity=RING)

        def e2e_step(k):
            s.step_begin(k)
            s.stations_record(k)                    # solver_output_stations' place in the step (psolve.c:4281)
            s.compute_force_source(F_host.a[k] if loaded.size else None)
            s.compute_force_stiffness()
            s.compute_force_damping()
            s.send_force_and_adjust()
            s.compute_displacement()
            s.send_displacement_and_adjust()
            if s.stations_pending() == RING:
                s.stations_drain(out=st_out.a)
        for k in range(args.warmup):
            e2e_step(k)
        s.stations_drain(out=st_out.a)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_step(k)
        s.stations_drain(out=st_out.a)
        s.fetch_all(hb.TM1, out=final.a)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        if not (np.isfinite(final.a).all() and np.isfinite(st_out.a).all()):
            raise SystemExit("non-finite displacements after the end-to-end run")
        e2e = {"value": e_global * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(24 * loaded.size),
               "d2h_bytes_per_step": int(72 * n_st + final.a.nbytes / args.steps),
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "what": "per-step C-ABI sequence (hgpu_step_begin, hgpu_stations_record, hgpu_force_source(host F), "
                       "hgpu_force_stiffness/_damping/_exchange, hgpu_update, hgpu_disp_exchange), station rows read back "
                       f"every {RING} steps (hgpu_stations_drain) + one final hgpu_fetch_all(tm1); host buffers page-locked "
                       "(hgpu_host_alloc)"}
        for b in (F_host, st_out, final):
            b.close()
    s.close()

    # ---- cpu baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.damping == "rayleigh" and not adaptive and not basin:
        try:
            r = reference_sample(300)
            cpu = {k: r.get(k) for k in ("value", "unit", "cores", "kind", "sample", "min", "max")}
        except Exception as e:                                   # reported baseline only
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"[:300]}

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = BYTES_PER_ELEM[args.damping] * E + BYTES_PER_NODE * N
        k_s = fused_s / max(tile_launches, 1)
        achieved = alg_bytes / k_s / 1e9 if k_s > 0 else None
        line = {
            "metric": METRIC, "value": e_global * args.steps / dev_s, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": (basin_config(n, info, args.damping, dt_run, freq_run, h_run) if basin else
                       workload_config(world, n, "peer-memory mailboxes (NVLink P2P)" if args.halo == "p2p" else "NCCL send/recv",
                                       args.damping, info if adaptive else None)),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity_check": parity,
            "roofline": {"bound": "hbm",
                         "kernel": "step_kernel<1,true,256> (dense 24 x 24 K1 / K2 stiffness + Rayleigh damping + update, fused)"
                                   if args.dense_k else
                                   (("step_kernel<1,false,256,true> (WPASS variant; " if args.wpass else "step_kernel<1,false,256> (") +
                                    "stiffness + Rayleigh damping + update, fused)" if args.damping == "rayleigh"
                                    else "step_kernel<3,false,256> (BKT memory variables + constant-Q force + update, fused)"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "bytes_per_element_update": alg_bytes / E,
                         "kernel_ms": 1e3 * k_s, "kernel_share_of_step": fused_s / dev_s if dev_s > 0 else None,
                         "traffic": None},
            "cpu_baseline": cpu,
            "layout": {k: layout[k] for k in ("tile_nodes", "ntiles", "max_tile_nodes", "tile_elems_total",
                                              "tile_halo_total", "n_regular", "n_special", "device_bytes",
                                              "smem_bytes", "block_threads", "grid_ctas", "ctas_per_sm",
                                              "early_tiles", "struct_tiles")},
            "phases_ms_per_step": phases_ms,
            "setup_s": {"mesh": round(t_mesh, 1), "hgpu_init": round(t_init, 1)},
        }
        # roofline.traffic: DRAM bytes per launch from an `ncu --set full` capture OF THIS WORKLOAD AND KERNEL
        # (profiles/traffic.json, keyed by workload / damping / kernel variant / elements), else null
        traffic_file = ROOT / "profiles" / "traffic.json"
        if traffic_file.exists():
            try:
                variant = "dense-k" if args.dense_k else "wpass" if args.wpass else "default"
                key = f"{args.workload}:{args.damping}:{variant}:{E}"
                line["roofline"]["traffic"] = json.loads(traffic_file.read_text()).get("by_workload", {}).get(key)
            except Exception:
                pass
        if args.stiffness == "conventional":
            line["config"]["workload"] = line["config"]["workload"].replace(
                "effective stiffness", "CONVENTIONAL stiffness (compute_addforce_conventional: " +
                ("literal dense K1 / K2 products, HGPU_FLAG_DENSE_K)" if args.dense_k else
                 "the same operator, applied in the factored form of the effective method)"))
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
