"""ctypes binding of ``libvasp_hemo.so`` (declared in ``include/vasp_hemo.h``).

There is deliberately no fallback: if the shared library has not been built, or no B200 is visible, every entry
point raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``python -m vasp_b200.build``).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libvasp_hemo.so"

PUSH_GLOBAL_FIRST = 1
PUSH_HALO_FIRST = 2


class VaspHemoError(RuntimeError):
    """A libvasp_hemo.so call failed; the message is ``vh_last_error()``."""


_lib = None

_P = C.c_void_p
_I64 = C.c_int64
_DBL = C.c_double

_SIGNATURES = {
    "vh_create": [C.c_int, C.POINTER(_P)],
    "vh_destroy": [_P],
    "vh_device_count": [C.POINTER(C.c_int)],
    "vh_set_mesh": [_P, _P, _I64, _P, _I64],
    "vh_set_velocity_layout": [_P, C.c_int, _P, _I64, _DBL, _P, C.POINTER(_I64), _I64],
    "vh_get_sizes": [_P, C.POINTER(_I64)],
    "vh_get_maps": [_P, _P, _P, _P, _P, _P, _P, _P, _P],
    "vh_get_geometry": [_P, _P, _P, _P],
    "vh_begin": [_P, _DBL, _DBL],
    "vh_set_tuning": [_P, _I64, _I64],
    "vh_set_wss_layout": [_P, _I64, _I64],
    "vh_push_snapshots": [_P, _P, _I64, _I64, C.c_int, _P],
    "vh_push_snapshots_device": [_P, _P, _I64, _I64, C.c_int, _P],
    "vh_set_host_compaction": [_P, C.c_int, C.c_int],
    "vh_get_compact_info": [_P, C.POINTER(_I64), C.POINTER(C.c_int)],
    "vh_get_wall_slots": [_P, _P],
    "vh_compact_snapshots": [_P, _P, _I64, _I64, _P],
    "vh_compact_rows": [_P, _P, _I64, _P],
    "vh_host_gather": [_P, _P, _I64, _I64, _P, _I64, C.POINTER(_I64), _P, _I64, C.c_int],
    "vh_push_compact": [_P, _P, _I64, _I64, C.c_int, _P],
    "vh_push_compact_device": [_P, _P, _I64, _I64, C.c_int, _P],
    "vh_get_io_stats": [_P, C.POINTER(_DBL), C.POINTER(_I64)],
    "vh_get_sums": [_P, _P, C.POINTER(_I64)],
    "vh_set_sums": [_P, _P, _I64],
    "vh_sums_device_ptr": [_P, C.POINTER(_P)],
    "vh_get_tau_last": [_P, _P],
    "vh_finalize": [_P, _I64, _P, _P, _P, _P, _P],
    "vh_sync": [_P],
    "vh_get_timers": [_P, C.POINTER(_DBL), C.POINTER(_DBL), C.POINTER(_I64)],
    "vh_set_profile": [_P, C.c_int],
    "vh_get_kernel_profile": [_P, C.POINTER(_DBL), C.POINTER(_DBL), C.POINTER(_I64)],
    "vh_timer_start": [_P],
    "vh_timer_stop": [_P, C.POINTER(_DBL)],
    "vh_alloc_pinned": [C.POINTER(_P), _I64],
    "vh_free_pinned": [_P],
    "vh_alloc_device": [_P, C.POINTER(_P), _I64],
    "vh_free_device": [_P, _P],
    "vh_memcpy_h2d": [_P, _P, _P, _I64],
    "vh_memcpy_d2h": [_P, _P, _P, _I64],
    "vh_flush_l2": [_P],
    "vh_mem_info": [_P, C.POINTER(_I64), C.POINTER(_I64)],
    "vh_nccl_unique_id": [C.c_char_p],
    "vh_nccl_init": [_P, C.c_char_p, C.c_int, C.c_int],
    "vh_nccl_allreduce_sums": [_P],
    "vh_nccl_allreduce_max": [_P, C.POINTER(_DBL)],
    "vh_nccl_barrier": [_P],
    "vh_nccl_destroy": [_P],
    "vh_peer_init": [_P],
    "vh_peer_reduce_finalize": [_P, _I64, _P, _P, _P, _P, _P],
}

EXPORTED_SYMBOLS = tuple(sorted(list(_SIGNATURES) + ["vh_last_error"]))


def load() -> C.CDLL:
    """Load the CUDA library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise VaspHemoError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built and vasp_b200 has no CPU path. "
                "Run `python -c \"import __graft_entry__ as g; g.build()\"` from the repository root.")
        lib = C.CDLL(str(LIB_PATH))
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        lib.vh_last_error.argtypes = []
        lib.vh_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().vh_last_error().decode(errors="replace")
        raise VaspHemoError(f"libvasp_hemo error {rc}: {msg}")
