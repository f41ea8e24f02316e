"""ctypes binding of libtaub200.so (C ABI declared in include/taub200.h).

There is NO fallback: if the CUDA library is missing or fails to load, importing a solver raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TAUB200_LIB: another build of the same ABI (A/B timing of two revisions on one box, tools/perf_quick.py)
LIB_PATH = os.environ.get("TAUB200_LIB") or os.path.join(_HERE, "libtaub200.so")

c_int, c_i64, c_vp, c_float = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float

GHOST = 2
COL0 = 4
MAX_LABELS = 64
BINARY, MULTIPHASE, ANISOTROPIC, MULTIPHASE_CLASS = 0, 1, 2, 3
OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3


class Geom(ctypes.Structure):
    _fields_ = [("bs", ctypes.c_int32), ("Nx", ctypes.c_int32), ("Ny", ctypes.c_int32), ("Nz", ctypes.c_int32),
                ("Nx_global", ctypes.c_int32), ("i_offset", ctypes.c_int32), ("periodic", ctypes.c_int32),
                ("planes", ctypes.c_int32), ("rows", ctypes.c_int32), ("pitch", ctypes.c_int32),
                ("plane_stride", ctypes.c_int64), ("image_stride", ctypes.c_int64)]


class Problem(ctypes.Structure):
    _fields_ = [("g", Geom), ("kind", ctypes.c_int32), ("L", ctypes.c_int32),
                ("field", c_vp * 2), ("codes", c_vp), ("labels", c_vp), ("lut", c_vp),
                ("omega", ctypes.c_float), ("cur", ctypes.c_int32), ("stop", c_vp),
                ("peer_lo", c_vp * 2), ("peer_hi", c_vp * 2), ("sync_ws", c_vp), ("sync_epoch", ctypes.c_int32),
                ("redo_ws", c_vp)]


# name -> (restype, argtypes); must list every symbol include/taub200.h declares
SIGNATURES = {
    "taub_abi_version": (c_int, []),
    "taub_last_error": (ctypes.c_char_p, []),
    "taub_device_info": (c_int, [ctypes.POINTER(c_int)] * 4),
    "taub_launch_count": (ctypes.c_ulonglong, []),
    "taub_set_device": (c_int, [c_int]),
    "taub_geom_init": (c_int, [ctypes.POINTER(Geom)] + [c_int] * 7),
    "taub_field_elems": (ctypes.c_size_t, [ctypes.POINTER(Geom)]),
    "taub_codes_elems": (ctypes.c_size_t, [ctypes.POINTER(Geom)]),
    "taub_sums_ws_bytes": (ctypes.c_size_t, [ctypes.POINTER(Geom)]),
    "taub_init_binary": (c_int, [ctypes.POINTER(Problem), c_vp, c_int, c_int, c_vp, c_vp]),
    "taub_init_multiphase": (c_int, [ctypes.POINTER(Problem), c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "taub_multiphase_keys": (c_int, [ctypes.POINTER(Problem), c_int, c_int, c_vp, c_vp]),
    "taub_class_ws_bytes": (ctypes.c_size_t, []),
    "taub_class_count": (c_int, [ctypes.POINTER(Problem), c_int, c_int, c_vp, c_vp]),
    "taub_class_assign": (c_int, [ctypes.POINTER(Problem), c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp]),
    "taub_init_anisotropic": (c_int, [ctypes.POINTER(Problem), c_vp, c_int, c_int, c_vp, c_vp]),
    "taub_init_electrode": (c_int, [ctypes.POINTER(Problem), c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "taub_plane_counts": (c_int, [ctypes.POINTER(Geom), c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "taub_refresh_ghosts": (c_int, [ctypes.POINTER(Geom), c_vp, c_int, c_int, c_vp]),
    "taub_half_sweep": (c_int, [ctypes.POINTER(Problem), c_i64, c_int, c_int, c_vp]),
    "taub_fused_sweep2": (c_int, [ctypes.POINTER(Problem), c_i64, c_int, c_int, c_vp]),
    "taub_inexact_events": (ctypes.c_ulonglong, []),
    "taub_redo_ws_ints": (ctypes.c_size_t, []),
    "taub_can_fuse": (c_int, [ctypes.POINTER(Problem)]),
    "taub_iterate": (c_int, [ctypes.POINTER(Problem), c_i64, c_int, c_int, c_vp]),
    "taub_can_reside": (c_int, [ctypes.POINTER(Problem)]),
    "taub_fused_plan": (c_int, [ctypes.POINTER(Problem), c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32)]),
    "taub_resident_pairs": (c_int, [ctypes.POINTER(Problem), c_i64, c_int, c_vp]),
    "taub_sync_ws_ints": (ctypes.c_size_t, []),
    "taub_resident_timeouts": (ctypes.c_ulonglong, []),
    "taub_resident_profile": (c_int, [ctypes.POINTER(ctypes.c_ulonglong), c_int]),
    "taub_plane_means": (c_int, [ctypes.POINTER(Problem), c_vp, c_vp, c_vp, c_vp]),
    "taub_check_async": (c_int, [ctypes.POINTER(Problem), c_vp, c_vp, c_vp, c_vp, c_vp, c_float, c_vp, c_vp]),
    "taub_stop_rule_async": (c_int, [c_int, c_int, c_vp, c_vp, c_vp, c_float, c_vp, c_vp, c_vp]),
    "taub_flood_round": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "taub_unpackbits": (c_i64, [c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t]),
    "taub_unlzw": (c_i64, [c_vp, ctypes.c_size_t, c_vp, ctypes.c_size_t]),
}

_lib = None


class TaubError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m taufactor_b200.build` "
                "(nvcc, sm_100a). taufactor_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("TAUB200_LIB") and not hasattr(lib, name):
                if name.endswith("_ws_ints"):      # an older build under A/B timing: a workspace it never touches
                    setattr(lib, name, lambda: 4096)
                continue            # ... and it may lack the newest diagnostics
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.taub_abi_version() != 13 and not os.environ.get("TAUB200_LIB"):
            raise ImportError("libtaub200.so ABI version mismatch; rebuild it")
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().taub_last_error().decode(errors="replace")
        raise TaubError(f"{what or 'libtaub200'} failed ({rc}): {msg}")
