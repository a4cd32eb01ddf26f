"""Python face of the CUDA engine: a thin, torch-free wrapper over the C ABI in ``include/vasp_hemo.h``.

:class:`HemoEngine` stands where the reference builds its dolfin objects inside ``compute_hemodyanamics``
(``src/vasp/postprocessing/postprocessing_fenics/compute_hemodynamics.py``):

=========================================  ============================================================
reference                                  here
=========================================  ============================================================
``Mesh`` read + ``BoundaryMesh`` (:187-191)   :meth:`HemoEngine.set_mesh`        (K0 on the device)
function spaces + transfer matrix (:204-223)  :meth:`HemoEngine.set_velocity_layout`
``Stress(...)`` (:247)                        :meth:`HemoEngine.begin`
loop body (:272-318)                          :meth:`HemoEngine.push`            (K2/K3)
final formulas (:326-346)                     :meth:`HemoEngine.finalize`        (K4)
=========================================  ============================================================
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import PUSH_GLOBAL_FIRST, PUSH_HALO_FIRST, VaspHemoError, check  # noqa: F401

SUM_ROWS = 15  # rows 0-8: sum tau (3*j + c), 9-11: sum |tau| (j), 12-14: sum P(|dtau/dt|) (j); each row nF long


def _ptr(a: Optional[np.ndarray]) -> C.c_void_p:
    return C.c_void_p(None) if a is None else C.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """Page-locked host array (freed when the array is garbage collected)."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib.vh_alloc_pinned(C.byref(p), max(nbytes, 1)))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.vh_free_pinned, p)
    return arr


def device_count() -> int:
    """CUDA devices visible to this process (raises if the library is missing: there is no CPU path)."""
    n = C.c_int(0)
    check(_lib.load().vh_device_count(C.byref(n)))
    return int(n.value)


class HemoEngine:
    """One engine per GPU (one process per GPU under a launcher)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.vh_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.nF = 0
        self.order = 0
        self.vec_len = 0
        self._finalizer = weakref.finalize(self, self._lib.vh_destroy, h)

    def close(self) -> None:
        if self._finalizer.alive:
            self._finalizer()

    # ---- K0 ------------------------------------------------------------------------------------------------
    def set_mesh(self, xyz: np.ndarray, tets: np.ndarray) -> None:
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        tets = np.ascontiguousarray(tets, dtype=np.int64)
        if xyz.ndim != 2 or xyz.shape[1] != 3 or tets.ndim != 2 or tets.shape[1] != 4:
            raise ValueError("xyz must be (nv,3) and tets (nc,4)")
        check(self._lib.vh_set_mesh(self._h, _ptr(xyz), xyz.shape[0], _ptr(tets), tets.shape[0]))
        self.nv = xyz.shape[0]
        self._refresh_sizes()

    def set_velocity_layout(self, order: int, refined_xyz: Optional[np.ndarray] = None,
                            n_nodes: Optional[int] = None, tol: Optional[float] = None,
                            node_perm: Optional[np.ndarray] = None,
                            comp_offset: Optional[Tuple[int, int, int]] = None, node_stride: int = 1) -> None:
        """Velocity vector layout: component ``c`` of node ``v`` sits at
        ``comp_offset[c] + node_stride * node_perm[v]`` (default: the blocked layout ``create_hdf5.py:158-174``
        writes)."""
        if order == 2:
            if refined_xyz is None:
                raise ValueError("order=2 needs the refined-mesh vertex coordinates")
            refined_xyz = np.ascontiguousarray(refined_xyz, dtype=np.float64)
            n_nodes = refined_xyz.shape[0]
            if tol is None:
                # largest extent of the bounding box, column by column (an axis-0 reduction of an (N, 3) array is 3x slower:
                # 0.4 s at 7 M nodes)
                tol = 1e-8 * max(float(refined_xyz[:, c].max() - refined_xyz[:, c].min()) for c in range(3))
        else:
            n_nodes = self.nv if n_nodes is None else int(n_nodes)
            refined_xyz = None
            tol = 0.0 if tol is None else tol
        if comp_offset is None:
            comp_offset = (0, n_nodes, 2 * n_nodes)
        if node_perm is not None:
            node_perm = np.ascontiguousarray(node_perm, dtype=np.int64)
            if node_perm.shape != (n_nodes,) or (n_nodes and node_perm.min() < 0):
                raise ValueError("node_perm must have one non-negative entry per velocity node")
        off = (C.c_int64 * 3)(*[int(x) for x in comp_offset])
        check(self._lib.vh_set_velocity_layout(self._h, int(order), _ptr(refined_xyz), int(n_nodes), float(tol),
                                               _ptr(node_perm), off, int(node_stride)))
        self.order = int(order)
        n_slots = n_nodes if node_perm is None else int(node_perm.max()) + 1  # node_perm may point into a longer vector
        self.vec_len = max(comp_offset) + (n_slots - 1) * node_stride + 1
        self._refresh_sizes()

    def _refresh_sizes(self) -> None:
        n = (C.c_int64 * 7)()
        check(self._lib.vh_get_sizes(self._h, n))
        (self.nF, self.nBV, self.n_wall_cells, self.n_multi, self.ndof, self.n_nodes,
         self.n_wall_nodes) = (int(x) for x in n)

    def maps(self) -> Dict[str, np.ndarray]:
        nF = self.nF
        out = {
            "facet_cell": np.empty(nF, np.int32), "facet_local": np.empty(nF, np.int8),
            "facets": np.empty((nF, 3), np.int32), "bcell_parent": np.empty((nF, 3), np.int32),
            "btopology": np.empty((nF, 3), np.int32), "bvert_parent": np.empty(self.nBV, np.int32),
            "bcell_local": np.empty((nF, 3), np.int8),
        }
        fn = np.empty((nF, self.ndof), np.int32) if self.order else None
        check(self._lib.vh_get_maps(self._h, _ptr(out["facet_cell"]), _ptr(out["facet_local"]), _ptr(out["facets"]),
                                    _ptr(out["bcell_parent"]), _ptr(out["btopology"]), _ptr(out["bvert_parent"]),
                                    _ptr(out["bcell_local"]), _ptr(fn)))
        if fn is not None:
            out["facet_nodes"] = fn
        return out

    def geometry(self) -> Dict[str, np.ndarray]:
        """normal (nF,3), area (nF,), glam (nF,4,3) in facet-canonical labels (boundary dofs 0,1,2 then the
        opposite vertex)."""
        nF = self.nF
        normal, area, glam = np.empty((nF, 3)), np.empty(nF), np.empty((nF, 4, 3))
        check(self._lib.vh_get_geometry(self._h, _ptr(normal), _ptr(area), _ptr(glam)))
        return {"normal": normal, "area": area, "glam": glam}

    # ---- time loop -------------------------------------------------------------------------------------------
    def begin(self, mu: float, dt: float) -> None:
        check(self._lib.vh_begin(self._h, float(mu), float(dt)))

    def set_tuning(self, batch_snapshots: int = 0, chunk_snapshots: int = 0) -> None:
        check(self._lib.vh_set_tuning(self._h, int(batch_snapshots), int(chunk_snapshots)))

    def set_wss_matrix(self, matrix: Optional[np.ndarray], first_column: int = 0) -> None:
        """Select the layout of the WSS output of :meth:`push`.

        ``matrix`` = C-contiguous float64 (9 nF, n_cols) array (ideally pinned): tau of the k-th non-halo snapshot
        pushed from now on becomes column ``first_column + k``, row ``9 f + 3 j + c`` -- the (dof x time) matrix
        VaSP's spectral tools build from WSS.h5 (``create_transformed_matrix``,
        postprocessing_h5py_common.py:226-246,337-343).  ``None`` restores one vector per snapshot."""
        if matrix is None:
            self._wss_matrix = None
            check(self._lib.vh_set_wss_layout(self._h, 0, 0))
            return
        if (matrix.dtype != np.float64 or matrix.ndim != 2 or not matrix.flags.c_contiguous
                or matrix.shape[0] != 9 * self.nF):
            raise ValueError("WSS matrix must be C-contiguous float64 with 9*nF rows")
        check(self._lib.vh_set_wss_layout(self._h, int(matrix.shape[1]), int(first_column)))
        self._wss_matrix = matrix

    def set_wss_layout(self, ld: int, first_column: int = 0) -> None:
        """Raw form of :meth:`set_wss_matrix` for device-resident output (:meth:`push_device`)."""
        check(self._lib.vh_set_wss_layout(self._h, int(ld), int(first_column)))

    def push(self, u: np.ndarray, flags: int = 0, keep_wss: bool = False,
             wss_out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        """Process ``u`` = (n_snap, >= vec_len) float64 rows (ideally a :func:`pinned_empty` buffer).

        Returns tau of every non-halo snapshot as (n, nF, 3 dofs, 3 comps) when ``keep_wss``; after
        :meth:`set_wss_matrix` the snapshots become columns of that matrix instead (and it is returned)."""
        if getattr(self, "_wss_matrix", None) is not None:
            if wss_out is not None or keep_wss:
                raise ValueError("set_wss_matrix() is active: the WSS output goes to that matrix")
            if u.dtype != np.float64 or u.ndim != 2 or u.strides[1] != 8 or u.shape[1] < self.vec_len:
                raise ValueError("u must be a 2-D float64 array with contiguous rows of at least vec_len entries")
            stride = u.strides[0] if u.shape[0] > 1 else u.shape[1] * 8
            check(self._lib.vh_push_snapshots(self._h, _ptr(u), u.shape[0], stride, int(flags),
                                              _ptr(self._wss_matrix)))
            return self._wss_matrix
        if u.dtype != np.float64 or u.ndim != 2 or u.strides[1] != 8:
            raise ValueError("u must be a 2-D float64 array with contiguous rows")
        if u.shape[1] < self.vec_len:
            raise ValueError(f"snapshot vectors have {u.shape[1]} entries, layout needs {self.vec_len}")
        n_real = u.shape[0] - (1 if flags & PUSH_HALO_FIRST else 0)
        if keep_wss and wss_out is None:
            wss_out = pinned_empty((n_real, self.nF, 3, 3))
        if wss_out is not None and (wss_out.dtype != np.float64 or not wss_out.flags.c_contiguous
                                    or wss_out.size < n_real * self.nF * 9):
            raise ValueError("wss_out must be C-contiguous float64 with n*nF*9 entries")
        stride = u.strides[0] if u.shape[0] > 1 else u.shape[1] * 8  # numpy reports stride 0 for a single row view
        check(self._lib.vh_push_snapshots(self._h, _ptr(u), u.shape[0], stride, int(flags), _ptr(wss_out)))
        return wss_out

    # ---- wall-layer compaction in front of the bus (include/vasp_hemo.h) ---------------------------------------------
    def set_host_compaction(self, mode: str = "auto", threads: int = 0) -> None:
        """``mode``: "auto" (gather the wall layer on the host when it is a small share of the vector), "off", "on"."""
        check(self._lib.vh_set_host_compaction(self._h, {"auto": 0, "off": 1, "on": 2}[mode], int(threads)))

    @property
    def compact_len(self) -> int:
        """Doubles per compact block: 3 components x the wall-layer node count rounded up to 32."""
        n = C.c_int64()
        check(self._lib.vh_get_compact_info(self._h, C.byref(n), None))
        return int(n.value)

    @property
    def compaction_active(self) -> bool:
        """Would :meth:`push` gather the wall layer on the host for the current layout and mode?"""
        a = C.c_int()
        check(self._lib.vh_get_compact_info(self._h, None, C.byref(a)))
        return bool(a.value)

    def wall_slots(self) -> np.ndarray:
        """Element offset inside a snapshot vector of every wall-layer node (ascending), ``(n_wall_nodes,)`` int64."""
        out = np.empty(self.n_wall_nodes, np.int64)
        check(self._lib.vh_get_wall_slots(self._h, _ptr(out)))
        return out

    def compact(self, u: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host gather of the wall layer: ``(n, >= vec_len)`` rows -> ``(n, compact_len)`` compact blocks."""
        if u.dtype != np.float64 or u.ndim != 2 or u.strides[1] != 8 or u.shape[1] < self.vec_len:
            raise ValueError("u must be a 2-D float64 array with contiguous rows of at least vec_len entries")
        n, m = u.shape[0], self.compact_len
        if out is None:
            out = pinned_empty((n, m))
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size < n * m:
            raise ValueError("out must be C-contiguous float64 with n * compact_len entries")
        stride = u.strides[0] if n > 1 else u.shape[1] * 8
        check(self._lib.vh_compact_snapshots(self._h, _ptr(u), n, stride, _ptr(out)))
        return out

    def compact_rows(self, addresses: np.ndarray, out: np.ndarray) -> np.ndarray:
        """Same for rows given by their host addresses (``uint64``), e.g. datasets inside an ``mmap`` of ``u.h5``."""
        addresses = np.ascontiguousarray(addresses, dtype=np.uint64)
        n, m = len(addresses), self.compact_len
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size < n * m:
            raise ValueError("out must be C-contiguous float64 with n * compact_len entries")
        check(self._lib.vh_compact_rows(self._h, _ptr(addresses), n, _ptr(out)))
        return out

    def push_compact(self, c: np.ndarray, flags: int = 0, wss_out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        """:meth:`push` for ``(n, compact_len)`` compact blocks (ideally pinned)."""
        m = self.compact_len
        if c.dtype != np.float64 or c.ndim != 2 or c.strides[1] != 8 or c.shape[1] < m:
            raise ValueError("c must be a 2-D float64 array with contiguous rows of at least compact_len entries")
        n_real = c.shape[0] - (1 if flags & PUSH_HALO_FIRST else 0)
        if wss_out is not None and getattr(self, "_wss_matrix", None) is None and (
                wss_out.dtype != np.float64 or not wss_out.flags.c_contiguous or wss_out.size < n_real * self.nF * 9):
            raise ValueError("wss_out must be C-contiguous float64 with n*nF*9 entries")
        if getattr(self, "_wss_matrix", None) is not None:
            wss_out = self._wss_matrix
        stride = c.strides[0] if c.shape[0] > 1 else c.shape[1] * 8
        check(self._lib.vh_push_compact(self._h, _ptr(c), c.shape[0], stride, int(flags), _ptr(wss_out)))
        return wss_out

    def push_compact_device(self, d_c: int, n_snap: int, stride_bytes: int, flags: int = 0, d_wss: int = 0) -> None:
        """Asynchronous launch over compact blocks already resident in device memory."""
        check(self._lib.vh_push_compact_device(self._h, C.c_void_p(d_c), int(n_snap), int(stride_bytes), int(flags),
                                               C.c_void_p(d_wss or None)))

    def io_stats(self) -> Dict[str, float]:
        g, b = C.c_double(), C.c_int64()
        check(self._lib.vh_get_io_stats(self._h, C.byref(g), C.byref(b)))
        return {"gather_ms": g.value, "h2d_bytes": int(b.value)}

    def push_device(self, d_u: int, n_snap: int, stride_bytes: int, flags: int = 0, d_wss: int = 0) -> None:
        """Asynchronous launch over snapshots already resident in device memory (raw device addresses)."""
        check(self._lib.vh_push_snapshots_device(self._h, C.c_void_p(d_u), int(n_snap), int(stride_bytes), int(flags),
                                                 C.c_void_p(d_wss or None)))

    def sums(self) -> Tuple[np.ndarray, int]:
        s = np.empty((SUM_ROWS, self.nF))
        cnt = C.c_int64()
        check(self._lib.vh_get_sums(self._h, _ptr(s), C.byref(cnt)))
        return s, int(cnt.value)

    def set_sums(self, sums: np.ndarray, count: int) -> None:
        s = np.ascontiguousarray(sums, dtype=np.float64)
        if s.shape != (SUM_ROWS, self.nF):
            raise ValueError(f"sums must be ({SUM_ROWS}, nF)")
        check(self._lib.vh_set_sums(self._h, _ptr(s), int(count)))

    def tau_last(self) -> np.ndarray:
        t = np.empty((self.nF, 3, 3))
        check(self._lib.vh_get_tau_last(self._h, _ptr(t)))
        return t

    def finalize(self, n_total: Optional[int] = None) -> Dict[str, np.ndarray]:
        if n_total is None:
            n_total = self.sums()[1]
        out = {k: np.empty((self.nF, 3)) for k in ("TAWSS", "OSI", "RRT", "ECAP", "TWSSG")}
        check(self._lib.vh_finalize(self._h, int(n_total), _ptr(out["TAWSS"]), _ptr(out["OSI"]), _ptr(out["RRT"]),
                                    _ptr(out["ECAP"]), _ptr(out["TWSSG"])))
        return out

    def sync(self) -> None:
        check(self._lib.vh_sync(self._h))

    def timers(self) -> Dict[str, float]:
        k, c, n = C.c_double(), C.c_double(), C.c_int64()
        check(self._lib.vh_get_timers(self._h, C.byref(k), C.byref(c), C.byref(n)))
        return {"kernel_ms": k.value, "h2d_ms": c.value, "launches": int(n.value)}

    def set_profile(self, on: bool) -> None:
        check(self._lib.vh_set_profile(self._h, int(bool(on))))

    def kernel_profile(self) -> Tuple[float, float, int]:
        """(summed k1_stage ms, summed k2_wall ms, launches of each) since the last call."""
        m1, m2, n = C.c_double(), C.c_double(), C.c_int64()
        check(self._lib.vh_get_kernel_profile(self._h, C.byref(m1), C.byref(m2), C.byref(n)))
        return m1.value, m2.value, int(n.value)

    def finalize_async(self, n_total: int) -> None:
        """Enqueue the final formulas only; results stay on the device (bench: device-resident timing)."""
        check(self._lib.vh_finalize(self._h, int(n_total), None, None, None, None, None))

    def timer_start(self) -> None:
        check(self._lib.vh_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        check(self._lib.vh_timer_stop(self._h, C.byref(ms)))
        return ms.value

    # ---- raw device memory (bench: inputs resident in HBM) -----------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self._lib.vh_alloc_device(self._h, C.byref(p), int(nbytes)))
        return int(p.value)

    def device_free(self, d_ptr: int) -> None:
        check(self._lib.vh_free_device(self._h, C.c_void_p(d_ptr)))

    def h2d(self, d_ptr: int, src: np.ndarray) -> None:
        src = np.ascontiguousarray(src)
        check(self._lib.vh_memcpy_h2d(self._h, C.c_void_p(d_ptr), _ptr(src), src.nbytes))

    def d2h(self, dst: np.ndarray, d_ptr: int) -> None:
        check(self._lib.vh_memcpy_d2h(self._h, _ptr(dst), C.c_void_p(d_ptr), dst.nbytes))

    def flush_l2(self) -> None:
        check(self._lib.vh_flush_l2(self._h))

    def mem_info(self) -> Tuple[int, int]:
        f, t = C.c_int64(), C.c_int64()
        check(self._lib.vh_mem_info(self._h, C.byref(f), C.byref(t)))
        return int(f.value), int(t.value)

    # ---- multi-GPU ------------------------------------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().vh_nccl_unique_id(buf))
        return buf.raw

    def nccl_init(self, uid: bytes, rank: int, world: int) -> None:
        if len(uid) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        check(self._lib.vh_nccl_init(self._h, uid, int(rank), int(world)))

    def allreduce_sums(self) -> None:
        check(self._lib.vh_nccl_allreduce_sums(self._h))

    def allreduce_max(self, value: float) -> float:
        v = C.c_double(float(value))
        check(self._lib.vh_nccl_allreduce_max(self._h, C.byref(v)))
        return v.value

    def barrier(self) -> None:
        check(self._lib.vh_nccl_barrier(self._h))

    def nccl_destroy(self) -> None:
        """Unmap the peers' memory and destroy the communicator (collective in practice: see NcclComm.close)."""
        check(self._lib.vh_nccl_destroy(self._h))

    def peer_init(self) -> None:
        """Map every rank's running sums over NVLink (CUDA IPC); collective, after ``nccl_init``."""
        check(self._lib.vh_peer_init(self._h))

    def peer_reduce_finalize(self, n_total: int, host: bool = True) -> Optional[Dict[str, np.ndarray]]:
        """Fused cross-GPU reduction + final formulas (one kernel per rank, peer loads over NVLink)."""
        if not host:
            check(self._lib.vh_peer_reduce_finalize(self._h, int(n_total), None, None, None, None, None))
            return None
        out = {k: np.empty((self.nF, 3)) for k in ("TAWSS", "OSI", "RRT", "ECAP", "TWSSG")}
        check(self._lib.vh_peer_reduce_finalize(self._h, int(n_total), _ptr(out["TAWSS"]), _ptr(out["OSI"]),
                                                _ptr(out["RRT"]), _ptr(out["ECAP"]), _ptr(out["TWSSG"])))
        return out
