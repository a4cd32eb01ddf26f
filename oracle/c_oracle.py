"""ctypes face of ``oracle/hemo_oracle.c`` (the threaded C twin of the numpy restatement).

TEST INFRASTRUCTURE ONLY -- see ``hemo_oracle.py``.  Used to time the CPU baseline and, in the tests, to check the
C twin against the numpy restatement.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from . import hemo_oracle as ho

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libhemo_oracle.so"


class _Maps(C.Structure):
    _fields_ = [("order", C.c_int32), ("ndof", C.c_int32), ("nF", C.c_int64), ("nW", C.c_int64),
                ("facet_wall", C.c_void_p), ("wall_nodes", C.c_void_p), ("glam", C.c_void_p),
                ("normal", C.c_void_p), ("area", C.c_void_p), ("facet_lv", C.c_void_p),
                ("bcell_local", C.c_void_p), ("Ainv", C.c_void_p)]


def build(force: bool = False) -> Path:
    src = HERE / "hemo_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        res = subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"building the C oracle failed:\n{res.stdout}\n{res.stderr}")
    return LIB


_lib = None


def _load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        lib = C.CDLL(str(LIB))
        lib.hemo_oracle_run.restype = C.c_int
        lib.hemo_oracle_run.argtypes = [C.POINTER(_Maps), C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64),
                                        C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.hemo_oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().hemo_oracle_max_threads())


class COracle:
    """Time loop of ``hemo_oracle.c`` on the maps of a numpy :class:`hemo_oracle.SurfaceStress`."""

    def __init__(self, stress: ho.SurfaceStress):
        m = stress.maps
        self.nF = stress.nF
        self._keep = {
            "facet_wall": np.ascontiguousarray(stress.facet_wall, dtype=np.int64),
            "wall_nodes": np.ascontiguousarray(m.cell_nodes[m.wall_cells], dtype=np.int64),
            "glam": np.ascontiguousarray(stress.glam, dtype=np.float64),
            "normal": np.ascontiguousarray(stress.normal, dtype=np.float64),
            "area": np.ascontiguousarray(stress.area, dtype=np.float64),
            "facet_lv": np.ascontiguousarray(stress.facet_lv, dtype=np.int8),
            "bcell_local": np.ascontiguousarray(m.bcell_local, dtype=np.int8),
            "Ainv": np.ascontiguousarray(np.linalg.inv(stress.A), dtype=np.float64),  # LUSolver: factor once
        }
        k = self._keep
        self._maps = _Maps(stress.order, k["wall_nodes"].shape[1], self.nF, len(m.wall_cells),
                           *[k[n].ctypes.data for n in ("facet_wall", "wall_nodes", "glam", "normal", "area",
                                                        "facet_lv", "bcell_local", "Ainv")])
        self.mu = stress.mu

    def run(self, u: np.ndarray, dt: float, comp_offset, node_stride: int = 1,
            tau_prev: Optional[np.ndarray] = None, keep_wss: bool = False, threads: int = 1) -> Dict:
        u = np.asarray(u)
        assert u.dtype == np.float64 and u.ndim == 2 and u.strides[1] == 8
        n_snap = u.shape[0]
        nF = self.nF
        wss_sum, tawss_sum, twssg_sum = np.zeros((nF, 3, 3)), np.zeros((nF, 3)), np.zeros((nF, 3))
        tau_last = np.zeros((nF, 3, 3))
        series = np.empty((n_snap, nF, 3, 3)) if keep_wss else None
        off = (C.c_int64 * 3)(*[int(x) for x in comp_offset])
        prev = None if tau_prev is None else np.ascontiguousarray(tau_prev, dtype=np.float64)
        stride = u.strides[0] // 8 if n_snap > 1 else u.shape[1]
        rc = _load().hemo_oracle_run(C.byref(self._maps), u.ctypes.data, n_snap, stride, off, int(node_stride),
                                     float(self.mu), float(dt), None if prev is None else prev.ctypes.data,
                                     wss_sum.ctypes.data, tawss_sum.ctypes.data, twssg_sum.ctypes.data,
                                     tau_last.ctypes.data, None if series is None else series.ctypes.data,
                                     int(threads))
        if rc != 0:
            raise MemoryError("hemo_oracle_run failed")
        out = {"wss_sum": wss_sum, "tawss_sum": tawss_sum, "twssg_sum": twssg_sum, "count": n_snap,
               "tau_last": tau_last}
        if keep_wss:
            out["wss"] = series
        return out
