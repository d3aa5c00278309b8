"""Loader for libhercules_gpu.so (C ABI: include/hercules_gpu.h).

The library is built in-tree by ``hercules_b200.build()`` (nvcc, sm_100a only).  There is no
Python or CPU implementation behind it: if the shared object is missing, or no CUDA device is
present when a solver is created, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
SO = PKG / "libhercules_gpu.so"
CSRC = PKG / "csrc"

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
p_i32, p_f64, p_f32 = C.POINTER(i32), C.POINTER(f64), C.POINTER(C.c_float)


class MsgList(C.Structure):
    _fields_ = [("count", i32), ("peer", p_i32), ("nodes", p_i32), ("mapping", p_i32)]


class Mesh(C.Structure):
    _fields_ = [("lenum", i32), ("nharbored", i32), ("ldnnum", i32),
                ("elem_lnid", p_i32), ("eTable", p_f64), ("nTable", p_f64), ("edata", p_f32),
                ("dnode", p_i32), ("K1", p_f64), ("K2", p_f64),
                ("dn_c", MsgList), ("dn_s", MsgList), ("an_c", MsgList), ("an_s", MsgList)]


class Params(C.Structure):
    _fields_ = [("dt", f64), ("dt2", f64), ("freq", f64), ("damping", i32), ("stiffness", i32),
                ("print_accel", i32), ("rank", i32), ("nranks", i32), ("nloaded", i32),
                ("loaded_lnid", p_i32), ("device", i32), ("tile_nodes", i32), ("flags", i32)]


class Timers(C.Structure):
    _fields_ = [(n, f64) for n in ("addforce_s", "addforce_e", "damping", "send_dn_force",
                                   "adjust_force", "send_an_force", "new_disp", "send_an_disp",
                                   "adjust_disp", "send_dn_disp", "fused_step")] + \
               [("launches", i64), ("steps", i64)]


class Layout(C.Structure):
    _fields_ = [("tile_nodes", i32), ("ntiles", i32), ("max_tile_nodes", i32),
                ("max_tile_elems", i32), ("tile_elems_total", i64), ("tile_halo_total", i64),
                ("n_regular", i64), ("n_special", i64), ("device_bytes", i64),
                ("smem_bytes", i32), ("block_threads", i32), ("grid_ctas", i32), ("ctas_per_sm", i32), ("early_tiles", i32),
                ("est_gather_wavefronts", f64), ("est_scatter_wavefronts", f64),
                ("max_tile_acc", i32), ("max_tile_recs", i32), ("max_tile_srcs", i32), ("struct_tiles", i32),
                ("partial_slots", i64), ("deps_total", i64)]


# every symbol include/hercules_gpu.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "hgpu_last_error": (C.c_char_p, []),
    "hgpu_abi_version": (C.c_int, []),
    "hgpu_device_count": (C.c_int, []),
    "hgpu_init": (C.c_int, [C.POINTER(_H), C.POINTER(Mesh), C.POINTER(Params)]),
    "hgpu_comm_unique_id": (C.c_int, [C.c_void_p]),
    "hgpu_comm_init": (C.c_int, [_H, C.c_void_p]),
    "hgpu_comm_p2p_export": (C.c_int, [_H, C.c_void_p, i32, C.POINTER(i32)]),
    "hgpu_comm_p2p_connect": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(i32)]),
    "hgpu_finalize": (C.c_int, [_H]),
    "hgpu_step_begin": (C.c_int, [_H, i32]),
    "hgpu_force_source": (C.c_int, [_H, C.c_void_p]),
    "hgpu_force_stiffness": (C.c_int, [_H]),
    "hgpu_force_damping": (C.c_int, [_H]),
    "hgpu_force_exchange": (C.c_int, [_H]),
    "hgpu_update": (C.c_int, [_H]),
    "hgpu_disp_exchange": (C.c_int, [_H]),
    "hgpu_step": (C.c_int, [_H, i32, C.c_void_p]),
    "hgpu_source_preload": (C.c_int, [_H, i32, i32, C.c_void_p]),
    "hgpu_force_source_resident": (C.c_int, [_H, i32]),
    "hgpu_fetch_all_async": (C.c_int, [_H, i32, C.c_void_p]),
    "hgpu_fetch_wait": (C.c_int, [_H]),
    "hgpu_run": (C.c_int, [_H, i32, i32, C.c_void_p]),
    "hgpu_fetch_nodes": (C.c_int, [_H, i32, C.c_void_p, i32, C.c_void_p]),
    "hgpu_fetch_all": (C.c_int, [_H, i32, C.c_void_p]),
    "hgpu_store_all": (C.c_int, [_H, i32, C.c_void_p]),
    "hgpu_stations_attach": (C.c_int, [_H, i32, C.c_void_p, C.c_void_p, i32, i32, i32, i32]),
    "hgpu_stations_record": (C.c_int, [_H, i32]),
    "hgpu_stations_pending": (C.c_int, [_H]),
    "hgpu_stations_drain": (C.c_int, [_H, C.c_void_p, C.c_void_p, i32, C.POINTER(i32)]),
    "hgpu_planes_attach": (C.c_int, [_H, i64, C.c_void_p, C.c_void_p]),
    "hgpu_planes_record": (C.c_int, [_H, C.c_void_p]),
    "hgpu_planes_wait": (C.c_int, [_H]),
    "hgpu_host_alloc": (C.c_void_p, [C.c_size_t]),
    "hgpu_host_free": (None, [C.c_void_p]),
    "hgpu_sync": (C.c_int, [_H]),
    "hgpu_get_timers": (C.c_int, [_H, C.POINTER(Timers)]),
    "hgpu_stream": (C.c_void_p, [_H]),
    "hgpu_get_layout": (C.c_int, [_H, C.POINTER(Layout)]),
    "hgpu_plan_build": (C.c_int, [C.POINTER(Mesh), i32, C.POINTER(Layout)]),
}


def build(verbose: bool = False) -> Path:
    """Compile libhercules_gpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(CSRC)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("building libhercules_gpu.so failed")
    return SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not SO.exists():
            raise RuntimeError(f"{SO} is missing: run hercules_b200.build() (nvcc, sm_100a). "
                               "There is no CPU fallback.")
        L = C.CDLL(str(SO))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the header and the .so disagree
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib
