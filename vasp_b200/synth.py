"""Seeded synthetic inputs: vessel-like tetrahedral meshes, their 1->8 refinement and velocity time series.

Used by ``bench.py``, ``__graft_entry__.smoke()`` and the tests (there is no network for datasets, and the
reference's own ``u.h5`` / ``mesh_refined_fluid.h5`` blobs are not shipped).  Everything is plain
numpy; nothing here is on the hot path.

Mesh recipe: ``n x n x m`` hexahedra on a rounded-square cross-section swept along a gently curved
centreline, Kuhn 6-tet split (conforming), vertex and cell numbering randomly permuted so that nothing depends on
structured order, cell rows ascending as dolfin stores them.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

# Kuhn / Freudenthal split of the unit cube into 6 tets sharing the 0-7 diagonal (corner = x + 2y + 4z)
_KUHN = np.array([[0, 1, 3, 7], [0, 1, 5, 7], [0, 2, 3, 7], [0, 2, 6, 7], [0, 4, 5, 7], [0, 4, 6, 7]])


def vessel_mesh(n: int, m: int, radius: float = 1.0, length: Optional[float] = None, bend: float = 0.15,
                stenosis: float = 0.0, bulge: float = 0.0, seed: Optional[int] = 1234) -> Dict[str, np.ndarray]:
    """``6*n*n*m`` tets.  Returns ``xyz (nv,3)``, ``tets (nc,4)`` and ``param (nv,3)`` = (a, b, s) with a, b in
    [-1,1] across the section (the wall is max(|a|,|b|) = 1) and s in [0,1] along the vessel."""
    if length is None:
        length = 2.0 * radius * m / n
    a = np.linspace(-1.0, 1.0, n + 1)
    s = np.linspace(0.0, 1.0, m + 1)
    A, B, S = np.meshgrid(a, a, s, indexing="ij")
    A, B, S = A.ravel(), B.ravel(), S.ravel()
    w = 0.85  # blend square -> disc (w = 1 would make the four corner cells degenerate)
    y = (1 - w) * A + w * A * np.sqrt(1.0 - 0.5 * B * B)
    z = (1 - w) * B + w * B * np.sqrt(1.0 - 0.5 * A * A)
    r = radius * (1.0 - stenosis * np.exp(-((S - 0.5) / 0.08) ** 2))
    r_z = r * (1.0 + bulge * np.exp(-((S - 0.6) / 0.1) ** 2) * (B > 0) * B)
    x = length * S
    yc = bend * length * np.sin(np.pi * S)  # curved centreline
    xyz = np.stack([x, yc + r * y + stenosis * 0.5 * radius * np.exp(-((S - 0.5) / 0.08) ** 2), r_z * z], axis=1)

    def vid(i, j, k):
        return (i * (n + 1) + j) * (m + 1) + k

    I, J, K = np.meshgrid(np.arange(n), np.arange(n), np.arange(m), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    corners = np.stack([vid(I + (c & 1), J + ((c >> 1) & 1), K + ((c >> 2) & 1)) for c in range(8)], axis=1)
    tets = corners[:, _KUHN].reshape(-1, 4)
    param = np.stack([A, B, S], axis=1)
    if seed is not None:
        rng = np.random.default_rng(seed)
        perm = rng.permutation(len(xyz))  # new id of old vertex
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        xyz, param = xyz[inv], param[inv]
        tets = perm[tets][rng.permutation(len(tets))]
    return {"xyz": np.ascontiguousarray(xyz), "tets": np.sort(tets, axis=1).astype(np.int64), "param": param}


def mesh_edges(tets: np.ndarray) -> np.ndarray:
    """Unique edges as (lo, hi) rows in lexicographic order."""
    t = np.sort(np.asarray(tets, dtype=np.int64), axis=1)
    pairs = np.concatenate([t[:, [i, j]] for i in range(4) for j in range(i + 1, 4)], axis=0)
    nv = int(t.max()) + 1
    key = np.unique(pairs[:, 0] * nv + pairs[:, 1])
    return np.stack([key // nv, key % nv], axis=1)


def p2_points(xyz: np.ndarray, tets: np.ndarray, seed: Optional[int] = None
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """P2 node set = refined-mesh vertex set: coarse vertices then edge midpoints (optionally shuffled).

    Returns ``(points (N2,3), edges (Ne,2), new_id (N2,))`` where ``new_id[i]`` is the position in ``points`` of
    un-shuffled node ``i`` (vertex ``i`` for ``i < nv``, else midpoint of ``edges[i - nv]``)."""
    edges = mesh_edges(tets)
    pts = np.concatenate([xyz, 0.5 * (xyz[edges[:, 0]] + xyz[edges[:, 1]])], axis=0)
    new_id = np.arange(len(pts))
    if seed is not None:
        new_id = np.random.default_rng(seed).permutation(len(pts))
        out = np.empty_like(pts)
        out[new_id] = pts
        pts = out
    return pts, edges, new_id


def refine_uniform(xyz: np.ndarray, tets: np.ndarray, seed: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Regular 1->8 split (what dolfin ``refine`` produces up to numbering; reference
    ``create_refined_mesh.py:50``).  Vertices = :func:`p2_points`; returns ``(xyz_refined, tets_refined)``."""
    t = np.sort(np.asarray(tets, dtype=np.int64), axis=1)
    pts, edges, new_id = p2_points(xyz, t, seed)
    nv = xyz.shape[0]
    ekey = edges[:, 0] * nv + edges[:, 1]

    def mid(a, b):
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        return nv + np.searchsorted(ekey, lo * nv + hi)

    v0, v1, v2, v3 = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    e01, e02, e03, e12, e13, e23 = mid(v0, v1), mid(v0, v2), mid(v0, v3), mid(v1, v2), mid(v1, v3), mid(v2, v3)
    kids = [(v0, e01, e02, e03), (v1, e01, e12, e13), (v2, e02, e12, e23), (v3, e03, e13, e23),
            (e01, e02, e03, e13), (e01, e02, e12, e13), (e02, e03, e13, e23), (e02, e12, e13, e23)]
    fine = np.stack([np.stack(k, axis=1) for k in kids], axis=1).reshape(-1, 4)
    return pts, np.sort(new_id[fine], axis=1)


# --------------------------------------------------------------------------------------------------------------
# velocity series: u(x, t) = sum_k coef_k(t) * basis_k(x)  (few spatial modes -> cheap to regenerate per batch)
# --------------------------------------------------------------------------------------------------------------
N_MODES = 4


def velocity_basis(points: np.ndarray, seed: int = 2024) -> np.ndarray:
    """``(N_MODES, 3, N)`` smooth vector fields: mode 0 is a Poiseuille-like axial profile, the others are
    seeded quadratic perturbations.  ``points`` are node coordinates (any units)."""
    rng = np.random.default_rng(seed)
    lo, hi = points.min(axis=0), points.max(axis=0)
    q = (points - 0.5 * (lo + hi)) / (0.5 * np.max(hi - lo))  # scaled to roughly [-1,1]
    x, y, z = q[:, 0], q[:, 1], q[:, 2]
    rho2 = (y * y + z * z) / max(float(np.max(y * y + z * z)), 1e-300)
    out = np.empty((N_MODES, 3, len(points)))
    out[0, 0], out[0, 1], out[0, 2] = 1.0 - rho2, 0.05 * y * x, 0.05 * z * x
    for k in range(1, N_MODES):
        c = rng.normal(size=(3, 10))
        mono = np.stack([np.ones_like(x), x, y, z, x * x, y * y, z * z, x * y, y * z, z * x])
        out[k] = (c @ mono) * (1.0 - 0.8 * rho2)
    return out


def velocity_coefficients(n_snap: int, period: float = 1.0, eps: float = 0.2, seed: int = 2024,
                          t0: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
    """Times ``(n_snap,)`` and mode coefficients ``(n_snap, N_MODES)`` (Womersley-like flow-rate waveform)."""
    rng = np.random.default_rng(seed + 1)
    dt = period / n_snap
    t = t0 + dt * np.arange(1, n_snap + 1)
    phase = rng.uniform(0, 2 * np.pi, size=N_MODES)
    coef = np.empty((n_snap, N_MODES))
    coef[:, 0] = 1.0 + 0.6 * np.sin(2 * np.pi * t / period) + 0.3 * np.sin(4 * np.pi * t / period)
    for k in range(1, N_MODES):
        coef[:, k] = eps * np.sin(2 * np.pi * k * t / period + phase[k])
    return t, coef


def velocity_series(basis: np.ndarray, coef: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Snapshot vectors in the blocked layout ``[x..., y..., z...]`` of ``create_hdf5.py:158-163``:
    ``(n_snap, 3*N)``."""
    k, _, n = basis.shape
    flat = basis.reshape(k, 3 * n)
    if out is None:
        return coef @ flat
    np.matmul(coef, flat, out=out[:, :3 * n])
    return out
