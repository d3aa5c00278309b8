"""hercules_oracle.py -- ctypes front end of oracle/hercules_oracle.c (liboracle.so) and of the
unmodified reference kernels (libref_kernels.so, present only where /root/reference was compiled).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg.  Never imported from hercules_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFDIR = HERE / "_ref"

RAYLEIGH, MASS, NONE, BKT = 0, 1, 2, 3          # damping.h:28
CONVENTIONAL, EFFECTIVE = 0, 1                  # stiffness.h:24
DISTRIBUTION, ASSIGNMENT = 0, 1

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build() -> None:
    """Compile the C restatement (always) -- building the checker is not using it."""
    subprocess.run(["make", "-s", "oracle"], cwd=HERE, check=True)


def build_ref() -> bool:
    """Compile the unmodified reference when /root/reference is present; False otherwise."""
    if not Path("/root/reference/quake/forward/psolve.c").exists():
        return False
    subprocess.run(["make", "-s", "-j8", "ref"], cwd=HERE, check=True)
    return True


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        so = REFDIR / "liboracle.so"
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        i32, f64 = C.c_int32, C.c_double
        L.ho_addforce_effective.argtypes = [i32, _i32p, _f64p, _f64p, _f64p]
        L.ho_addforce_conventional.argtypes = [i32, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.ho_damping_addforce.argtypes = [i32, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.ho_calc_conv.argtypes = [i32, _i32p, _f32p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, f64, f64]
        L.ho_constant_Q_addforce.argtypes = [i32, _i32p, _f64p, _f32p, _f64p, _f64p, _f64p, _f64p,
                                             _f64p, _f64p, _f64p, f64, f64]
        L.ho_addforce_s.argtypes = [i32, _i32p, _f64p, f64, _f64p]
        L.ho_compute_displacement.argtypes = [i32, _f64p, _f64p, _f64p, C.c_void_p, _f64p]
        L.ho_compute_adjust.argtypes = [i32, _i32p, _f64p, i32, i32]
        L.ho_compute_K.argtypes = [_f64p, _f64p]
        L.ho_compute_setab.argtypes = [i32, f64, C.POINTER(f64), C.POINTER(f64)]
        L.ho_solver_init_tables.argtypes = [i32, i32, _i32p, _i8p, _f32p, _i64p, _i64p, f64, f64,
                                            f64, f64, f64, f64, _f64p, _f64p]
        L.ho_solver_init_tables.restype = i32
        for f in ("ho_addforce_effective", "ho_addforce_conventional", "ho_damping_addforce",
                  "ho_calc_conv", "ho_constant_Q_addforce", "ho_addforce_s",
                  "ho_compute_displacement", "ho_compute_adjust", "ho_compute_K",
                  "ho_compute_setab"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def ref_kernels():
    """The reference's own stiffness.c/damping.c behind flat wrappers, or None if not built."""
    global _ref
    if _ref is None:
        so = REFDIR / "libref_kernels.so"
        if not so.exists():
            return None
        L = C.CDLL(str(so))
        i32, f64 = C.c_int32, C.c_double
        L.refk_addforce_effective.argtypes = [i32, i32, _i32p, _f64p, _f64p, _f64p]
        L.refk_addforce_conventional.argtypes = [i32, i32, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.refk_damping_addforce.argtypes = [i32, i32, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.refk_calc_conv.argtypes = [i32, i32, _i32p, _f32p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, f64, f64]
        L.refk_constant_Q_addforce.argtypes = [i32, i32, _i32p, _f64p, _f32p, _f64p, _f64p, _f64p,
                                               _f64p, _f64p, _f64p, _f64p, f64, f64]
        for f in ("refk_addforce_effective", "refk_addforce_conventional", "refk_damping_addforce",
                  "refk_calc_conv", "refk_constant_Q_addforce"):
            getattr(L, f).restype = None
        _ref = L
    return _ref


def compute_K():
    K1 = np.zeros((8, 8, 3, 3)); K2 = np.zeros((8, 8, 3, 3))
    lib().ho_compute_K(K1.reshape(-1), K2.reshape(-1))
    return K1, K2


def compute_setab(damping: int, freq: float):
    a, b = C.c_double(), C.c_double()
    lib().ho_compute_setab(damping, freq, C.byref(a), C.byref(b))
    return a.value, b.value


class Mesh:
    """Flat view of one rank's mesh_t + solver tables, as dumped by ref_dump or generated."""

    def __init__(self, lnid, eTable, nTable, dnode=None, edata=None, K1=None, K2=None):
        self.lnid = np.ascontiguousarray(lnid, np.int32)
        self.eTable = np.ascontiguousarray(eTable, np.float64)
        self.nTable = np.ascontiguousarray(nTable, np.float64)
        self.dnode = np.ascontiguousarray(dnode if dnode is not None else np.zeros((0, 6)), np.int32)
        self.E, self.N, self.D = self.lnid.shape[0], self.nTable.shape[0], self.dnode.shape[0]
        self.edata = (np.ascontiguousarray(edata, np.float32) if edata is not None
                      else np.zeros((self.E, 14), np.float32))
        if K1 is None:
            K1, K2 = compute_K()
        self.K1 = np.ascontiguousarray(K1, np.float64).reshape(8, 8, 3, 3)
        self.K2 = np.ascontiguousarray(K2, np.float64).reshape(8, 8, 3, 3)

    @classmethod
    def from_dump(cls, d):
        return cls(d["elem_lnid"], d["eTable"], d["nTable"], d["dnode"], d["elem_edata"],
                   d["K1"], d["K2"])


class State:
    def __init__(self, mesh: Mesh, bkt: bool = False, accel: bool = False):
        N = mesh.N
        self.tm1 = np.zeros((N, 3)); self.tm2 = np.zeros((N, 3)); self.force = np.zeros((N, 3))
        self.tm3 = np.zeros((N, 3)) if accel else None
        self.conv = np.zeros((4, 8 * mesh.E, 3)) if bkt else None


def add_forces(m: Mesh, s: State, damping: int, stiffness: int, freq: float, dt: float) -> None:
    """solver_compute_force_stiffness + solver_compute_force_damping (psolve.c:3962-4006)."""
    L = lib()
    K1, K2 = m.K1.reshape(-1), m.K2.reshape(-1)
    et, ln = m.eTable.reshape(-1), m.lnid.reshape(-1)
    t1, t2, f = s.tm1.reshape(-1), s.tm2.reshape(-1), s.force.reshape(-1)
    if damping != BKT:
        if stiffness == EFFECTIVE:
            L.ho_addforce_effective(m.E, ln, et, t1, f)
        else:
            L.ho_addforce_conventional(m.E, ln, et, K1, K2, t1, f)
    if damping in (RAYLEIGH, MASS):
        L.ho_damping_addforce(m.E, ln, et, K1, K2, t1, t2, f)
    elif damping == BKT:
        c = s.conv
        L.ho_calc_conv(m.E, ln, m.edata.reshape(-1), t1, t2, c[0].reshape(-1), c[1].reshape(-1),
                       c[2].reshape(-1), c[3].reshape(-1), freq, dt)
        L.ho_constant_Q_addforce(m.E, ln, et, m.edata.reshape(-1), t1, t2, c[0].reshape(-1),
                                 c[1].reshape(-1), c[2].reshape(-1), c[3].reshape(-1), f, freq, dt)


def step(m: Mesh, s: State, damping: int, stiffness: int, freq: float, dt: float,
         loaded_lnid=None, F=None) -> None:
    """One pass of the solver_run loop body on one rank without neighbours (psolve.c:4265-4319):
    swap, source, forces, adjust(DISTRIBUTION), update, adjust(ASSIGNMENT)."""
    L = lib()
    s.tm1, s.tm2 = s.tm2, s.tm1
    if loaded_lnid is not None and len(loaded_lnid):
        L.ho_addforce_s(len(loaded_lnid), np.ascontiguousarray(loaded_lnid, np.int32),
                        np.ascontiguousarray(F, np.float64).reshape(-1), dt * dt,
                        s.force.reshape(-1))
    add_forces(m, s, damping, stiffness, freq, dt)
    L.ho_compute_adjust(m.D, m.dnode.reshape(-1), s.force.reshape(-1), 3, DISTRIBUTION)
    tm3 = s.tm3.ctypes.data if s.tm3 is not None else None
    L.ho_compute_displacement(m.N, m.nTable.reshape(-1), s.tm1.reshape(-1), s.tm2.reshape(-1),
                              tm3, s.force.reshape(-1))
    L.ho_compute_adjust(m.D, m.dnode.reshape(-1), s.tm2.reshape(-1), 3, ASSIGNMENT)


def run(m: Mesh, nsteps: int, damping: int, stiffness: int, freq: float, dt: float,
        loaded_lnid, forces, snapshot_steps=(), accel=False):
    """Time loop; returns (state, {step: tm1 copy taken at the top of that step after the swap}),
    the same instant ref_dump taps (psolve.c:4271-4275)."""
    s = State(m, bkt=(damping == BKT), accel=accel)
    snaps = {}
    L = lib()
    for k in range(nsteps):
        s.tm1, s.tm2 = s.tm2, s.tm1
        if k in snapshot_steps:
            snaps[k] = s.tm1.copy()
        s.tm1, s.tm2 = s.tm2, s.tm1          # undo: step() swaps itself
        step(m, s, damping, stiffness, freq, dt, loaded_lnid, forces[k] if len(loaded_lnid) else None)
    return s, snaps
