"""Oracle-backed stand-in for ``vasp_b200.engine.HemoEngine`` (TESTS ONLY).

The entry point's host logic -- file formats, flag handling, block reader, WSS series, stride, raw turtleFSI input,
WSS matrix -- is worth running in the CPU suite too, where no GPU exists.  This class offers the part of
``HemoEngine``'s public face that ``vasp_b200/compute_hemodynamics.py`` uses and computes with the numpy restatement
of the reference (``oracle/hemo_oracle.py``).  Tests monkeypatch it in; the product never imports it."""
from __future__ import annotations

import numpy as np

from oracle import hemo_oracle as ho

PUSH_GLOBAL_FIRST, PUSH_HALO_FIRST = 1, 2


class OracleHemoEngine:
    #: tests flip this to exercise the entry point's compact route (rows gathered out of the file mapping); the gather
    #: itself is the library's handle-free ``vh_host_gather`` -- data movement, no GPU needed
    compact_route = False

    def __init__(self, device: int = 0):
        self.device = device
        self._S = None
        self._matrix = None

    # ---- wall-layer compaction stand-ins (include/vasp_hemo.h)
    @property
    def compaction_active(self):
        return bool(self.compact_route)

    def _slots(self):
        m = self._probe.maps
        if self._order == 2:
            cn = self._node[ho.p2_cell_nodes(self._tets)[0][m.wall_cells]]
        else:
            cn = self._node[ho.order_cells(self._tets)[m.wall_cells]]
        return (np.unique(cn) * self._stride).astype(np.int32)

    @property
    def compact_len(self):
        return 3 * ((len(self._slots()) + 31) // 32 * 32)

    def compact_rows(self, addresses, out):
        import ctypes
        from vasp_b200 import _lib
        sl = self._slots()
        nwp = self.compact_len // 3
        slp = np.concatenate([sl, np.full(nwp - len(sl), sl[-1], np.int32)])
        addresses = np.ascontiguousarray(addresses, dtype=np.uint64)
        off = (ctypes.c_int64 * 3)(*self._off)
        assert _lib.load().vh_host_gather(addresses.ctypes.data, None, 0, len(addresses), slp.ctypes.data, nwp, off,
                                          out.ctypes.data, out.strides[0], 2) == 0
        return out

    def push_compact(self, c, flags=0, wss_out=None):
        sl, nwp = self._slots().astype(np.int64), self.compact_len // 3
        top = max(self._off) + int(sl.max()) + 1
        u = np.full((len(c), top), np.nan)  # anything outside the wall layer must never be read
        for k, o in enumerate(self._off):
            u[:, o + sl] = np.asarray(c)[:, k * nwp:k * nwp + len(sl)]
        return self.push(u, flags=flags, wss_out=wss_out)

    # ---- K0 stand-ins
    def set_mesh(self, xyz, tets):
        self._xyz, self._tets = np.asarray(xyz, dtype=np.float64), np.asarray(tets, dtype=np.int64)
        self._probe = ho.SurfaceStress(self._xyz, self._tets, 1.0, 1)
        self.nF = self._probe.nF

    def set_velocity_layout(self, order, refined_xyz=None, n_nodes=None, tol=None, node_perm=None, comp_offset=None,
                            node_stride=1):
        self._order = order
        if order == 2:
            pts = np.asarray(refined_xyz, dtype=np.float64)
            n = len(pts)
            p2 = ho.p2_node_coordinates(self._xyz, ho.p2_cell_nodes(self._tets)[1])
            tol = 1e-8 * float(np.ptp(pts, axis=0).max()) if tol is None else tol
            node = ho.match_points(p2, pts, tol)
        else:
            n = len(self._xyz) if n_nodes is None else n_nodes
            node = np.arange(len(self._xyz))
        if node_perm is not None:
            node = np.asarray(node_perm, dtype=np.int64)[node]
        self._node = node
        self._off = tuple(comp_offset) if comp_offset is not None else (0, n, 2 * n)
        self._stride = int(node_stride)

    def maps(self):
        m = self._probe.maps
        return {"facets": m.facets, "facet_cell": m.facet_cell, "facet_local": m.facet_local,
                "bcell_parent": m.bcell_parent, "btopology": m.btopology, "bvert_parent": m.bvert_parent,
                "bcell_local": m.bcell_local}

    # ---- time loop
    def begin(self, mu, dt):
        if self._order == 2:
            self._S = ho.SurfaceStress(self._xyz, self._tets, mu, 2, self._node)
        else:
            self._S = ho.SurfaceStress(self._xyz, self._tets, mu, 1)
            self._S.maps.cell_nodes = self._node[self._S.maps.cell_nodes]
        self._dt = float(dt)
        nF = self.nF
        self._wss = np.zeros((nF, 3, 3))
        self._tawss = np.zeros((nF, 3))
        self._twssg = np.zeros((nF, 3))
        self._count, self._prev = 0, None

    def set_tuning(self, batch_snapshots=0, chunk_snapshots=0):
        pass

    def set_wss_matrix(self, matrix, first_column=0):
        self._matrix, self._col = matrix, first_column

    def push(self, u, flags=0, keep_wss=False, wss_out=None):
        u = np.asarray(u)
        if flags & PUSH_HALO_FIRST:
            self._prev = self._S(u[0], self._off, self._stride)
            u = u[1:]
        elif flags & PUSH_GLOBAL_FIRST:
            self._prev = None
        elif self._prev is None:
            raise RuntimeError("first push of a time loop needs a flag")
        r = ho.run_time_loop(self._S, u, self._dt, self._off, self._stride, tau_prev=self._prev, keep_wss=True)
        self._prev = r["tau_last"]
        self._wss += r["wss_sum"]
        self._tawss += r["tawss_sum"]
        self._twssg += r["twssg_sum"]
        self._count += r["count"]
        if self._matrix is not None:
            k = len(u)
            self._matrix[:, self._col:self._col + k] = r["wss"].reshape(k, -1).T
            self._col += k
            return self._matrix
        if wss_out is not None:
            wss_out.reshape(-1)[:r["wss"].size] = r["wss"].reshape(-1)
            return wss_out
        return r["wss"] if keep_wss else None

    def sums(self):
        s = np.concatenate([self._wss.reshape(self.nF, 9).T, self._tawss.T, self._twssg.T])
        return s, self._count

    def set_sums(self, sums, count):
        s = np.asarray(sums)
        self._wss, self._tawss, self._twssg = s[:9].T.reshape(self.nF, 3, 3).copy(), s[9:12].T.copy(), s[12:].T.copy()
        self._count = int(count)

    def finalize(self, n_total=None):
        return ho.finalize(self._wss, self._tawss, self._twssg, self._count if n_total is None else n_total)

    def timers(self):
        return {"kernel_ms": 0.0, "h2d_ms": 0.0, "launches": 0}

    def close(self):
        pass
