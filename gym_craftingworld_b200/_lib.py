"""ctypes binding of libcw_b200.so (the C ABI declared in include/cw_b200.h).

There is NO fallback: if the library cannot be built or loaded the import of the package fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 4
STATS_LEN = 24
STATS_REPLICAS = 16
MAX_SIDE = 64
CHAIN_MAX_POS = 1024
F_AUTO_RESET = 1
F_DELTA_TRANSPORT = 2

# every symbol include/cw_b200.h declares (tests check the library exports exactly these)
SYMBOLS = ["cw_abi_version", "cw_error_string", "cw_reset", "cw_step", "cw_render", "cw_step_render", "cw_step_render_chained", "cw_step_chained", "cw_step_render_edit", "cw_step_delta", "cw_rollout", "cw_prefill_resets",
           "cw_imagine", "cw_frame_policy", "cw_onehot", "cw_render_alt", "cw_host_create", "cw_host_reset", "cw_host_bind_actions", "cw_host_step",
           "cw_host_step_many", "cw_host_load_state", "cw_host_stats", "cw_host_device_state", "cw_host_stream", "cw_host_fetch_frames", "cw_host_sync",
           "cw_host_destroy"]


class CwConfig(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("cell_stride", C.c_int32), ("max_steps", C.c_int32),
                ("subset_reward", C.c_int32), ("stacking", C.c_int32), ("n_selected", C.c_int32),
                ("number_of_tasks", C.c_int32), ("selected", C.c_uint8 * 16)]


class CwState(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("init_grid", C.c_void_p), ("agent", C.c_void_p), ("goal", C.c_void_p),
                ("t", C.c_void_p), ("episode", C.c_void_p), ("n", C.c_int64), ("seed", C.c_uint64),
                ("env_id_base", C.c_uint64), ("fixed_grid", C.c_void_p), ("fixed_agent", C.c_void_p),
                ("n_fixed", C.c_int64), ("goal_grid", C.c_void_p), ("goal_agent", C.c_void_p), ("init_agent", C.c_void_p),
                ("reset_rec", C.c_void_p), ("reset_list", C.c_void_p)]


class CwError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    vp, i64, u64, ci = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
    cfgp, stp = C.POINTER(CwConfig), C.POINTER(CwState)
    lib.cw_abi_version.restype = ci
    lib.cw_abi_version.argtypes = []
    lib.cw_error_string.restype = C.c_char_p
    lib.cw_error_string.argtypes = [ci]
    protos = {
        "cw_reset": [cfgp, stp, vp, vp, vp, vp, vp],
        "cw_step": [cfgp, stp, vp, vp, vp, vp, ci, vp],
        "cw_render": [cfgp, vp, vp, vp, i64, vp],
        "cw_step_render": [cfgp, stp, vp, vp, vp, vp, vp, vp, vp, ci, vp],
        "cw_step_render_edit": [cfgp, stp, vp, vp, vp, vp, vp, vp, vp, ci, vp, vp],
        "cw_step_render_chained": [cfgp, stp, vp, vp, vp, vp, vp, vp, vp, ci, vp, ci, ci, vp],
        "cw_rollout": [cfgp, stp, vp, vp, vp, vp, ci, ci, vp],
        "cw_step_chained": [cfgp, stp, vp, vp, vp, vp, ci, vp, ci, vp],
        "cw_step_delta": [cfgp, stp, vp, vp, vp, vp, ci, ci, vp],
        "cw_imagine": [cfgp, stp, vp, vp],
        "cw_prefill_resets": [cfgp, stp, vp],
        "cw_onehot": [cfgp, vp, vp, vp, i64, vp],
        "cw_frame_policy": [cfgp, vp, i64, vp, vp],
        "cw_render_alt": [cfgp, vp, vp, vp, i64, vp],
        "cw_host_create": [cfgp, i64, ci, u64, u64, ci, C.POINTER(vp)],
        "cw_host_reset": [vp, vp, vp],
        "cw_host_step": [vp, vp, vp, vp, vp],
        "cw_host_bind_actions": [vp, vp],
        "cw_host_step_many": [vp, vp, ci, vp, vp, vp],
        "cw_host_load_state": [vp, vp, vp, vp, vp, vp],
        "cw_host_stream": [vp, C.POINTER(vp)],
        "cw_host_sync": [vp],
        "cw_host_fetch_frames": [vp, vp, vp],
        "cw_host_stats": [vp, vp],
        "cw_host_device_state": [vp, stp, C.POINTER(vp)],
        "cw_host_destroy": [vp],
    }
    for name, args in protos.items():
        if os.environ.get("CW_LIB_PATH") and not hasattr(lib, name):
            continue                                               # A/B against an older build: newer entry points are absent
        fn = getattr(lib, name)
        fn.restype = ci
        fn.argtypes = args


def load():
    """Build (if stale) and load libcw_b200.so; raises if that is impossible -- there is no CPU path."""
    global _lib
    if _lib is None:
        path = os.environ.get("CW_LIB_PATH") or _build.build()     # CW_LIB_PATH: A/B experiments with another build
        if not os.path.exists(path):
            raise ImportError(f"gym_craftingworld_b200: CUDA library missing at {path}")
        lib = C.CDLL(path)
        _declare(lib)
        if lib.cw_abi_version() != ABI_VERSION and not os.environ.get("CW_LIB_PATH"):
            raise ImportError(f"gym_craftingworld_b200: ABI mismatch ({lib.cw_abi_version()} != {ABI_VERSION})")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().cw_error_string(rc).decode()
        raise CwError(f"{what or 'libcw_b200'} failed: {msg} (code {rc})")


def cell_stride(H: int, W: int) -> int:
    return (H * W + 15) // 16 * 16
