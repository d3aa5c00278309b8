"""Host-side mirror of the reference's solver_* call sequence over the C ABI.

``Solver`` keeps the names and the order of the static wrappers solver_run calls once per time
step (quake/forward/psolve.c:3953-4163, 4265-4319), so a parity test reads like the reference's
own loop body::

    s.step_begin(step)                 # swap tm1/tm2                 psolve.c:4271-4273
    s.compute_force_source(F)          # solver_compute_force_source  psolve.c:3953
    s.compute_force_stiffness()        #                              psolve.c:3962
    s.compute_force_damping()          #                              psolve.c:3983
    s.send_force_and_adjust()          # phases 8-10                  psolve.c:4036-4058
    s.compute_displacement()           #                              psolve.c:4072
    s.send_displacement_and_adjust()   # phases 13-15                 psolve.c:4130-4154

All arithmetic happens in libhercules_gpu.so on the GPU; numpy arrays here are only the host
buffers handed across the ABI.  Errors raise ``HerculesGpuError`` (the reference aborts the job).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

RAYLEIGH, MASS, NONE, BKT = 0, 1, 2, 3      # damping_type_t, damping.h:28
CONVENTIONAL, EFFECTIVE = 0, 1              # stiffness_type_t, stiffness.h:24
TM1, TM2, TM3, FORCE = 1, 2, 3, 4
CONV_SHEAR_1, CONV_SHEAR_2, CONV_KAPPA_1, CONV_KAPPA_2 = 5, 6, 7, 8   # psolve.h:308-311, [8 E][3] each
FLAG_NO_FUSE = 1
FLAG_TIMERS = 2
FLAG_NO_OVERLAP = 4
FLAG_TAIL_OVERLAP = 8
FLAG_WPASS = 16
FLAG_NO_STRUCT = 32
FLAG_STRUCT = 64
FLAG_DENSE_K = 128


class HerculesGpuError(RuntimeError):
    pass


def _chk(rc: int) -> None:
    if rc != 0:
        msg = _lib.lib().hgpu_last_error()
        raise HerculesGpuError(f"hgpu error {rc}: {msg.decode() if msg else '?'}")


@dataclass
class MsgList:
    """One side (c-list or s-list) of a schedule_t (psolve.h:255-272), flattened."""
    peer: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    mapping: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))

    @classmethod
    def from_dump(cls, hdr, mapping):
        hdr = np.asarray(hdr, np.int32).reshape(-1, 2)
        return cls(np.ascontiguousarray(hdr[:, 0]), np.ascontiguousarray(hdr[:, 1]),
                   np.ascontiguousarray(mapping, np.int32))


@dataclass
class HostMesh:
    """One rank's mesh_t + solver tables as flat numpy arrays (see include/hercules_gpu.h)."""
    elem_lnid: np.ndarray            # [E][8] int32
    eTable: np.ndarray               # [E][4] f64
    nTable: np.ndarray               # [N][7] f64
    dnode: np.ndarray | None = None  # [D][6] int32
    edata: np.ndarray | None = None  # [E][14] f32
    K1: np.ndarray | None = None     # [8][8][3][3]
    K2: np.ndarray | None = None
    dn_c: MsgList = field(default_factory=MsgList)
    dn_s: MsgList = field(default_factory=MsgList)
    an_c: MsgList = field(default_factory=MsgList)
    an_s: MsgList = field(default_factory=MsgList)

    @classmethod
    def from_dump(cls, d: dict) -> "HostMesh":
        return cls(d["elem_lnid"], d["eTable"], d["nTable"], d["dnode"], d["elem_edata"],
                   d["K1"], d["K2"],
                   MsgList.from_dump(d["dn_c_hdr"], d["dn_c_map"]),
                   MsgList.from_dump(d["dn_s_hdr"], d["dn_s_map"]),
                   MsgList.from_dump(d["an_c_hdr"], d["an_c_map"]),
                   MsgList.from_dump(d["an_s_hdr"], d["an_s_map"]))


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None and a.size else None


class Solver:
    def __init__(self, mesh: HostMesh, dt: float, damping: int = RAYLEIGH,
                 stiffness: int = EFFECTIVE, freq: float = 0.0, loaded_lnid=None,
                 rank: int = 0, nranks: int = 1, device: int = -1, tile_nodes: int = 0,
                 flags: int = 0, print_accel: bool = False, dt2: float | None = None):
        L = _lib.lib()
        self._L = L
        self._h = C.c_void_p()
        keep = []

        def arr(a, dt_, shape=None):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dt_)
            if shape is not None:
                a = a.reshape(shape)
            keep.append(a)
            return a

        lnid = arr(mesh.elem_lnid, np.int32, (-1, 8))
        et = arr(mesh.eTable, np.float64, (-1, 4))
        nt = arr(mesh.nTable, np.float64, (-1, 7))
        dn = arr(mesh.dnode, np.int32, (-1, 6)) if mesh.dnode is not None else None
        ed = arr(mesh.edata, np.float32, (-1, 14)) if mesh.edata is not None else None
        K1 = arr(mesh.K1, np.float64) if mesh.K1 is not None else None
        K2 = arr(mesh.K2, np.float64) if mesh.K2 is not None else None
        self.E, self.N = lnid.shape[0], nt.shape[0]
        self.D = 0 if dn is None else dn.shape[0]
        if et.shape[0] != self.E:
            raise ValueError("eTable and elem_lnid disagree on the element count")

        m = _lib.Mesh()
        m.lenum, m.nharbored, m.ldnnum = self.E, self.N, self.D
        m.elem_lnid, m.eTable, m.nTable = _p(lnid, C.c_int32), _p(et, C.c_double), _p(nt, C.c_double)
        m.edata, m.dnode = _p(ed, C.c_float), _p(dn, C.c_int32)
        m.K1, m.K2 = _p(K1, C.c_double), _p(K2, C.c_double)
        for name in ("dn_c", "dn_s", "an_c", "an_s"):
            ml = getattr(mesh, name)
            peer, nodes, mp = arr(ml.peer, np.int32), arr(ml.nodes, np.int32), arr(ml.mapping, np.int32)
            c = getattr(m, name)
            c.count = peer.size
            c.peer, c.nodes, c.mapping = _p(peer, C.c_int32), _p(nodes, C.c_int32), _p(mp, C.c_int32)

        ll = arr(loaded_lnid if loaded_lnid is not None else np.zeros(0), np.int32)
        self.nloaded = int(ll.size)
        p = _lib.Params()
        p.dt, p.dt2, p.freq = dt, (dt * dt if dt2 is None else dt2), freq
        p.damping, p.stiffness, p.print_accel = damping, stiffness, int(print_accel)
        p.rank, p.nranks, p.nloaded, p.loaded_lnid = rank, nranks, self.nloaded, _p(ll, C.c_int32)
        p.device, p.tile_nodes, p.flags = device, tile_nodes, flags
        _chk(L.hgpu_init(C.byref(self._h), C.byref(m), C.byref(p)))
        self.damping, self.stiffness = damping, stiffness

    # -- life cycle -----------------------------------------------------------------------------
    def close(self) -> None:
        if self._h:
            self._L.hgpu_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, unique_id: bytes) -> None:
        buf = C.create_string_buffer(unique_id, 128)
        _chk(self._L.hgpu_comm_init(self._h, buf))

    def p2p_export(self) -> bytes:
        """This rank's mailbox description for the peer-memory halo transport."""
        n = C.c_int32()
        _chk(self._L.hgpu_comm_p2p_export(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _chk(self._L.hgpu_comm_p2p_export(self._h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def p2p_connect(self, blobs) -> None:
        """blobs[r] = p2p_export() of rank r (all-gathered by the host)."""
        bufs = [C.create_string_buffer(b, len(b)) for b in blobs]
        ptrs = (C.c_void_p * len(bufs))(*[C.cast(b, C.c_void_p) for b in bufs])
        sizes = (C.c_int32 * len(bufs))(*[len(b) for b in blobs])
        _chk(self._L.hgpu_comm_p2p_connect(self._h, ptrs, sizes))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _chk(_lib.lib().hgpu_comm_unique_id(buf))
        return buf.raw

    # -- the solver_run loop body ------------------------------------------------------------------
    def step_begin(self, step: int) -> None:
        _chk(self._L.hgpu_step_begin(self._h, step))

    def compute_force_source(self, F) -> None:
        if self.nloaded == 0:
            return
        F = np.ascontiguousarray(F, np.float64)
        if F.size != 3 * self.nloaded:
            raise ValueError("F must be [nloaded][3]")
        _chk(self._L.hgpu_force_source(self._h, F.ctypes.data))

    def compute_force_source_resident(self, step: int) -> None:
        """compute_addforce_s from the rows source_preload left in HBM (no per-step host traffic)."""
        _chk(self._L.hgpu_force_source_resident(self._h, step))

    def compute_force_stiffness(self) -> None:
        _chk(self._L.hgpu_force_stiffness(self._h))

    def compute_force_damping(self) -> None:
        _chk(self._L.hgpu_force_damping(self._h))

    def send_force_and_adjust(self) -> None:
        _chk(self._L.hgpu_force_exchange(self._h))

    def compute_displacement(self) -> None:
        _chk(self._L.hgpu_update(self._h))

    def send_displacement_and_adjust(self) -> None:
        _chk(self._L.hgpu_disp_exchange(self._h))

    def step(self, step: int, F=None) -> None:
        if self.nloaded:
            F = np.ascontiguousarray(F, np.float64)
            _chk(self._L.hgpu_step(self._h, step, F.ctypes.data))
        else:
            _chk(self._L.hgpu_step(self._h, step, None))

    def source_preload(self, step0: int, F_all) -> int:
        """Make source rows for steps step0.. ([nsteps][nloaded][3]) resident in HBM; returns nsteps."""
        F_all = np.ascontiguousarray(F_all, np.float64).reshape(-1, max(self.nloaded, 1), 3)
        if self.nloaded:
            _chk(self._L.hgpu_source_preload(self._h, step0, F_all.shape[0], F_all.ctypes.data))
        return F_all.shape[0]

    def run(self, step0: int, nsteps: int, F_all=None) -> None:
        if self.nloaded and F_all is not None:
            F_all = np.ascontiguousarray(F_all, np.float64)
            if F_all.size < 3 * self.nloaded * nsteps:
                raise ValueError("F_all must be [nsteps][nloaded][3]")
            _chk(self._L.hgpu_run(self._h, step0, nsteps, F_all.ctypes.data))
        else:
            _chk(self._L.hgpu_run(self._h, step0, nsteps, None))

    # -- taps ------------------------------------------------------------------------------------------
    def _rows(self, which: int) -> int:
        return 8 * self.E if CONV_SHEAR_1 <= which <= CONV_KAPPA_2 else self.N

    def fetch_all(self, which: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self._rows(which), 3), np.float64)
        elif out.dtype != np.float64 or out.size != 3 * self._rows(which) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array of the selected array's size")
        _chk(self._L.hgpu_fetch_all(self._h, which, out.ctypes.data))
        return out

    def fetch_all_async(self, which: int, out: np.ndarray) -> None:
        """Snapshot now (stream order), copy to `out` (ideally a PinnedArray's .a) in the background."""
        if out.shape != (self._rows(which), 3) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 [rows][3] array")
        _chk(self._L.hgpu_fetch_all_async(self._h, which, out.ctypes.data))

    def fetch_wait(self) -> None:
        _chk(self._L.hgpu_fetch_wait(self._h))

    def store_all(self, which: int, a) -> None:
        a = np.ascontiguousarray(a, np.float64)
        if a.size != 3 * self._rows(which):
            raise ValueError("array must be [nharbored][3] ([8 lenum][3] for a conv array)")
        _chk(self._L.hgpu_store_all(self._h, which, a.ctypes.data))

    def fetch_nodes(self, which: int, lnid, out: np.ndarray | None = None) -> np.ndarray:
        lnid = np.ascontiguousarray(lnid, np.int32)
        if out is None:
            out = np.empty((lnid.size, 3), np.float64)
        _chk(self._L.hgpu_fetch_nodes(self._h, which, lnid.ctypes.data, lnid.size, out.ctypes.data))
        return out

    # -- stations on the device (interpolate_station_displacements, psolve.c:6680-6795) ---------------
    def stations_attach(self, nodes, localcoords, vel: bool = False, acc: bool = False, rate: int = 0,
                        capacity: int = 1024) -> None:
        nodes = np.ascontiguousarray(nodes, np.int32).reshape(-1, 8)
        loc = np.ascontiguousarray(localcoords, np.float64).reshape(-1, 3)
        if nodes.shape[0] != loc.shape[0]:
            raise ValueError("nodes [n][8] and localcoords [n][3] disagree on the station count")
        self._nst, self._st_cap = nodes.shape[0], capacity
        _chk(self._L.hgpu_stations_attach(self._h, self._nst, nodes.ctypes.data, loc.ctypes.data, int(vel), int(acc),
                                          rate, capacity))

    def stations_record(self, step: int) -> None:
        _chk(self._L.hgpu_stations_record(self._h, step))

    def stations_pending(self) -> int:
        n = self._L.hgpu_stations_pending(self._h)
        if n < 0:
            _chk(n)
        return n

    def stations_drain(self, out: np.ndarray | None = None):
        """(steps [r], rows [r][nstations][9] = dis, vel, acc) recorded since the last drain."""
        cap = self._st_cap
        if out is None:
            out = np.empty((cap, self._nst, 9), np.float64)
        steps = np.empty(cap, np.int32)
        n = C.c_int32()
        _chk(self._L.hgpu_stations_drain(self._h, out.ctypes.data, steps.ctypes.data, cap, C.byref(n)))
        return steps[:n.value].copy(), out[:n.value]

    # -- planes on the device (Old_planes_print's interpolation, io_planes.c:168-191) ----------------
    def planes_attach(self, nodes, localcoords) -> None:
        nodes = np.ascontiguousarray(nodes, np.int32).reshape(-1, 8)
        loc = np.ascontiguousarray(localcoords, np.float64).reshape(-1, 3)
        if nodes.shape[0] != loc.shape[0]:
            raise ValueError("nodes [n][8] and localcoords [n][3] disagree on the point count")
        self._npl = nodes.shape[0]
        _chk(self._L.hgpu_planes_attach(self._h, self._npl, nodes.ctypes.data, loc.ctypes.data))

    def planes_record(self, out: np.ndarray | None = None) -> np.ndarray:
        """Starts the interpolation of every plane point from tm1 and its copy to `out` ([npoints][3]);
        `out` is valid after planes_wait()."""
        if out is None:
            out = np.empty((self._npl, 3), np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == 3 * self._npl
        _chk(self._L.hgpu_planes_record(self._h, out.ctypes.data))
        return out

    def planes_wait(self) -> None:
        _chk(self._L.hgpu_planes_wait(self._h))

    def sync(self) -> None:
        _chk(self._L.hgpu_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self._L.hgpu_stream(self._h) or 0)

    def timers(self) -> dict:
        t = _lib.Timers()
        _chk(self._L.hgpu_get_timers(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in t._fields_}

    def layout(self) -> dict:
        t = _lib.Layout()
        _chk(self._L.hgpu_get_layout(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in t._fields_}


class PinnedArray:
    """A float64 numpy array over page-locked host memory from hgpu_host_alloc (freed with the
    object): hand ``.a`` to fetch_all / fetch_nodes / step for full-speed, bounce-free copies."""

    def __init__(self, shape):
        self._L = _lib.lib()
        n = int(np.prod(shape))
        self._p = self._L.hgpu_host_alloc(max(1, 8 * n))
        if not self._p:
            raise HerculesGpuError("hgpu_host_alloc failed: " + (self._L.hgpu_last_error() or b"?").decode())
        buf = (C.c_double * n).from_address(self._p)
        self.a = np.frombuffer(buf, np.float64, n).reshape(shape)

    def close(self) -> None:
        if self._p:
            self.a = None
            self._L.hgpu_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def plan_build(elem_lnid, nharbored: int, tile_nodes: int = 0, mesh: "HostMesh | None" = None) -> dict:
    """Host-only: build and self-check the tile plan for a mesh (no GPU needed); returns its sizes.
    With ``mesh`` (a rank's HostMesh) the nodes of its halo schedules and hanging-node lists make
    their tiles "self" tiles, as hgpu_init does on a multi-rank mesh."""
    L = _lib.lib()
    lnid = np.ascontiguousarray(elem_lnid, np.int32).reshape(-1, 8)
    m = _lib.Mesh()
    m.lenum, m.nharbored, m.ldnnum = lnid.shape[0], int(nharbored), 0
    m.elem_lnid = _p(lnid, C.c_int32)
    keep = []
    if mesh is not None:
        if mesh.dnode is not None and len(mesh.dnode):
            dn = np.ascontiguousarray(mesh.dnode, np.int32).reshape(-1, 6)
            keep.append(dn)
            m.ldnnum, m.dnode = dn.shape[0], _p(dn, C.c_int32)
        for name in ("dn_c", "dn_s", "an_c", "an_s"):
            ml = getattr(mesh, name)
            arrs = [np.ascontiguousarray(a, np.int32) for a in (ml.peer, ml.nodes, ml.mapping)]
            keep += arrs
            c = getattr(m, name)
            c.count = arrs[0].size
            c.peer, c.nodes, c.mapping = (_p(a, C.c_int32) for a in arrs)
    out = _lib.Layout()
    _chk(L.hgpu_plan_build(C.byref(m), tile_nodes, C.byref(out)))
    return {n: getattr(out, n) for n, _ in out._fields_}
