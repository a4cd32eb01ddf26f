"""CPU restatement (numpy, fp64) of VaSP's wall-shear-stress post-processing path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vasp_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and
there only as the checker or as the timed CPU baseline.

PARITY UNPINNED at the 1e-10 level: the reference evaluates this path inside legacy FEniCS
(dolfin/FFC/FIAT/PETSc), none of which is installable here, and the reference's own input blobs for its
single test are not shipped (``.MISSING_LARGE_BLOBS``).  What *is* pinned: the reference test's known answer
(``tests/test_compute_hemodynamics.py:68-73``: wall-averaged TAWSS of Poiseuille flow in (1.95, 2.05)) and its
OSI range assertion (``:84-88``, ``compute_hemodynamics.py:366-372``); see ``tests/test_oracle.py``.  Also pinned, by
RUNNING the reference's own Python in the build container (``tests/golden/make_reference_goldens.py`` ->
``tests/test_reference_goldens.py``): the traction formula against the UFL expression of ``Stress.__init__``
(``:142-150``) evaluated with numeric operands, the dof copy map against ``InterpolateDG.__call__`` (``:65-89``) and the whole
time-loop bookkeeping against ``compute_hemodyanamics`` (``:160-372``) executed on emulated dolfin objects, with only
``Stress`` and ``project_dg`` standing on this file's restatements.  Still unpinned: the finite-element assembly
inside dolfin/FFC (facet quadrature, the 7-point rule of ``project_dg``) and dolfin's boundary-mesh numbering.

The restatement is deliberately *literal*: it assembles the same facet integrals dolfin would assemble, with the
quadrature rules FFC would pick, solves the same block systems and re-does the reference's coordinate-matching
DOF copy.  The CUDA path uses closed forms instead, so agreement between the two is a real check.

Reference map (``src/vasp/postprocessing/postprocessing_fenics/compute_hemodynamics.py`` unless noted):

=====================  =========================================================================
function here          reference lines
=====================  =========================================================================
``order_cells``        dolfin ``Mesh.order()`` on read (``:187-189``) [dolfin-recall]
``exterior_facets``    ``BoundaryMesh(mesh, "exterior")`` ``:191`` + facet->cell ``:59-61``
``boundary_mesh``      dolfin ``BoundaryComputation`` vertex numbering + ``Mesh.order()`` (``order=True``) [dolfin-recall]
``p2_cell_nodes``      ``VectorFunctionSpace(mesh, "CG", 2)`` ``:206`` (UFC P2 local order)
``match_points``       ``PETScDMCollection.create_transfer_matrix`` ``:223`` applied at ``:275``
``SurfaceStress``      ``Stress`` ``:120-157``, ``SurfaceProjector`` ``:92-117``, ``InterpolateDG`` ``:32-89``
``project_dg_norm``    ``project_dg`` (``postprocessing_fenics_common.py:31-54``) called at ``:311``
``run_time_loop``      snapshot loop ``:271-318``
``finalize``           ``:326-346`` and the OSI assertion ``:366-372``
=====================  =========================================================================
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

# UFC reference tetrahedron, P2: dofs 0-3 vertices, 4-9 edges (local vertex pairs) [dolfin-recall]
P2_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))

# FFC default facet rule for polynomial degree 2 on a triangle (FIAT "default" scheme, 3 points, exact to
# degree 2); barycentric points, weights as fractions of the facet area.  [dolfin-recall]
_Q2_PTS = np.array([[2 / 3, 1 / 6, 1 / 6], [1 / 6, 1 / 6, 2 / 3], [1 / 6, 2 / 3, 1 / 6]])
_Q2_WTS = np.array([1 / 3, 1 / 3, 1 / 3])

# FIAT "default" triangle rule for degree 5 (Strang-Fix / Radon, 7 points) used for
# project_dg(inner(w, w) ** (1/2)): UFL estimates (1+1)+2 = 4 for the non-integer power, +1 test function.
_a, _b = 0.10128650732345633, 0.79742698535308720
_c, _d = 0.47014206410511505, 0.05971587178976981
Q5_PTS = np.array([
    [1 / 3, 1 / 3, 1 / 3],
    [_a, _b, _a], [_a, _a, _b], [_b, _a, _a],
    [_c, _d, _c], [_c, _c, _d], [_d, _c, _c],
])
Q5_WTS = np.array([0.225] + [0.12593918054482717] * 3 + [0.13239415278850616] * 3)


# --------------------------------------------------------------------------------------------------
# R1: mesh topology
# --------------------------------------------------------------------------------------------------
def order_cells(tets: np.ndarray) -> np.ndarray:
    """Cell vertices ascending, as dolfin's ordered meshes store them."""
    return np.sort(np.asarray(tets, dtype=np.int64), axis=1)


def exterior_facets(tets: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Exterior facets in dolfin's facet order (lexicographic in the sorted vertex triple).

    Returns ``(facets (nF,3) ascending vertex ids, facet_cell (nF,), facet_local (nF,))``; local facet k of a
    cell is the face opposite local vertex k.
    """
    tets = order_cells(tets)
    nc = tets.shape[0]
    keep = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    faces = tets[:, keep].reshape(-1, 3)  # row 4*cell + k, already ascending
    nv = int(tets.max()) + 1
    if nv < (1 << 21):  # the triple fits one int64 key: same lexicographic order, one sort instead of three
        order = np.argsort((faces[:, 0] * nv + faces[:, 1]) * nv + faces[:, 2], kind="stable")
    else:
        order = np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))
    fs = faces[order]
    same_next = np.zeros(len(fs), dtype=bool)
    same_next[:-1] = (fs[1:] == fs[:-1]).all(axis=1)
    same_prev = np.zeros(len(fs), dtype=bool)
    same_prev[1:] = same_next[:-1]
    ext = ~(same_next | same_prev)
    idx = order[ext]
    return fs[ext], idx // 4, (idx % 4).astype(np.int8)


def boundary_mesh(xyz: np.ndarray, tets: np.ndarray, facets: np.ndarray, facet_cell: np.ndarray,
                  facet_local: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Boundary triangle mesh the way ``BoundaryMesh(mesh, "exterior")`` (``compute_hemodynamics.py:191``) builds it.

    Returns ``(bvert_parent (nBV,), btopology (nF,3) boundary vertex numbers, bcell_parent (nF,3) parent vertex
    ids in boundary-cell vertex order)``.  dolfin's ``BoundaryComputation`` numbers the boundary vertices by first
    encounter over the exterior facets in facet order; the reference leaves ``order`` at its default ``True``, so
    ``Mesh.order()`` runs afterwards and every boundary cell lists its vertices ascending in *boundary* vertex
    number (the right-oriented variant, first two vertices swapped towards the outward normal, is what
    ``order=False`` would give and is not what the reference uses).
    """
    flat = facets.reshape(-1)
    _, first = np.unique(flat, return_index=True)
    first.sort()
    bvert_parent = flat[first]
    number = np.full(int(xyz.shape[0]), -1, dtype=np.int64)
    number[bvert_parent] = np.arange(len(bvert_parent))
    bnum = number[facets]
    perm = np.argsort(bnum, axis=1, kind="stable")
    return bvert_parent, np.take_along_axis(bnum, perm, axis=1), np.take_along_axis(facets, perm, axis=1)


def mesh_edges(tets: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Unique edges (lexicographic (lo, hi)) and the (Nc, 6) cell->edge map in UFC P2 edge order."""
    tets = order_cells(tets)
    pairs = np.stack([tets[:, [a for a, _ in P2_EDGES]], tets[:, [b for _, b in P2_EDGES]]], axis=-1)
    flat = pairs.reshape(-1, 2)
    nv = int(tets.max()) + 1
    key = flat[:, 0] * nv + flat[:, 1]
    ukey, inv = np.unique(key, return_inverse=True)
    edges = np.stack([ukey // nv, ukey % nv], axis=1)
    return edges, inv.reshape(-1, 6)


def p2_cell_nodes(tets: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """P2 node ids per cell: 4 vertices then 6 edge nodes (id = Nv + edge index).  Returns (cell_nodes, edges)."""
    tets = order_cells(tets)
    edges, cell_edges = mesh_edges(tets)
    nv = int(tets.max()) + 1
    return np.concatenate([tets, nv + cell_edges], axis=1), edges


def p2_node_coordinates(xyz: np.ndarray, edges: np.ndarray) -> np.ndarray:
    return np.concatenate([xyz, 0.5 * (xyz[edges[:, 0]] + xyz[edges[:, 1]])], axis=0)


# --------------------------------------------------------------------------------------------------
# R2: CG1(refined) -> CG2(coarse) transfer == point match
# --------------------------------------------------------------------------------------------------
def match_points(query: np.ndarray, cloud: np.ndarray, tol: float) -> np.ndarray:
    """Index into ``cloud`` of the point coinciding with each ``query`` point (within ``tol``)."""
    from scipy.spatial import cKDTree

    dist, idx = cKDTree(cloud).query(query, k=1)
    if np.any(dist > tol):
        bad = int(np.argmax(dist))
        raise ValueError(f"P2 node {bad} has no refined-mesh vertex within {tol} (nearest {dist[bad]:.3e})")
    return idx.astype(np.int64)


# --------------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------------
def barycentric_gradients(xyz: np.ndarray, tets: np.ndarray) -> np.ndarray:
    """grad(lambda_a) for the 4 vertices of each tet: (n, 4, 3)."""
    p = xyz[tets]  # (n,4,3)
    m = np.concatenate([np.ones(p.shape[:2] + (1,)), p], axis=2)  # rows [1 x y z]
    inv = np.linalg.inv(m)  # columns are coefficients of lambda_a
    return np.transpose(inv[:, 1:, :], (0, 2, 1))


def _p2_basis_gradients(lam: np.ndarray, glam: np.ndarray) -> np.ndarray:
    """Gradients of the 10 P2 basis functions at barycentric point(s).

    ``lam`` (..., 4), ``glam`` (..., 4, 3) -> (..., 10, 3).
    """
    out = np.empty(lam.shape[:-1] + (10, 3))
    for a in range(4):
        out[..., a, :] = (4.0 * lam[..., a, None] - 1.0) * glam[..., a, :]
    for e, (a, b) in enumerate(P2_EDGES):
        out[..., 4 + e, :] = 4.0 * (lam[..., a, None] * glam[..., b, :] + lam[..., b, None] * glam[..., a, :])
    return out


@dataclass
class WallMaps:
    """Everything that is constant over the time loop (the index maps are what the CUDA precompute must
    reproduce bit-exactly)."""
    order: int
    facets: np.ndarray        # (nF,3) parent vertex ids ascending
    facet_cell: np.ndarray    # (nF,)
    facet_local: np.ndarray   # (nF,) int8
    bvert_parent: np.ndarray  # (nBV,)
    btopology: np.ndarray     # (nF,3)
    bcell_parent: np.ndarray  # (nF,3) parent vertex ids in boundary-cell order
    bcell_local: np.ndarray   # (nF,3) local cell vertex (0..3) matched to boundary dof j  (R5)
    cell_nodes: np.ndarray    # (Nc, 4|10) velocity-vector node index per cell dof (R2 composed)
    wall_cells: np.ndarray    # unique cells owning exterior facets
    n_ext: np.ndarray         # exterior facets per wall cell


class SurfaceStress:
    """``Stress`` + ``SurfaceProjector`` + ``InterpolateDG`` for one mesh (literal assembly)."""

    def __init__(self, xyz: np.ndarray, tets: np.ndarray, mu: float, order: int = 2,
                 node_of_p2: Optional[np.ndarray] = None):
        """``node_of_p2``: velocity-vector node index of every P2 node (vertices then edges); identity when
        ``None``.  For ``order=1`` the velocity lives on the mesh vertices."""
        self.xyz = np.asarray(xyz, dtype=np.float64)
        self.tets = order_cells(tets)
        self.mu = float(mu)
        self.order = order
        facets, fcell, flocal = exterior_facets(self.tets)
        bvp, btopo, bcp = boundary_mesh(self.xyz, self.tets, facets, fcell, flocal)
        if order == 2:
            cn, self.edges = p2_cell_nodes(self.tets)
            if node_of_p2 is not None:
                cn = np.asarray(node_of_p2, dtype=np.int64)[cn]
        elif order == 1:
            cn, self.edges = self.tets.copy(), None
        else:
            raise ValueError("order must be 1 or 2")
        self.nF = len(facets)
        wall_cells, inv, counts = np.unique(fcell, return_inverse=True, return_counts=True)
        self.facet_wall = inv  # facet -> row in wall_cells
        self.glam = barycentric_gradients(self.xyz, self.tets[wall_cells])  # (nW,4,3)
        # outward unit normal and area per exterior facet
        g = self.glam[inv, flocal.astype(np.int64)]
        self.normal = -g / np.linalg.norm(g, axis=1, keepdims=True)
        p0, p1, p2 = (self.xyz[facets[:, i]] for i in range(3))
        self.area = 0.5 * np.linalg.norm(np.cross(p1 - p0, p2 - p0), axis=1)
        # facet vertices as local cell vertices (ascending = the three != facet_local)
        keep = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
        self.facet_lv = keep[flocal.astype(np.int64)]  # (nF,3)
        # ---- SurfaceProjector.__init__: A = assemble(inner(u,v)*ds, keep_diagonal) ; ident_zeros ------
        nW = len(wall_cells)
        A = np.zeros((nW, 4, 4))
        for q in range(3):
            phi = np.zeros((self.nF, 4))
            np.put_along_axis(phi, self.facet_lv, np.broadcast_to(_Q2_PTS[q], (self.nF, 3)), axis=1)
            np.add.at(A, inv, (_Q2_WTS[q] * self.area)[:, None, None] * phi[:, :, None] * phi[:, None, :])
        zero_rows = ~np.any(A != 0.0, axis=2)
        A[zero_rows, np.nonzero(zero_rows)[1]] = 1.0  # ident_zeros()
        self.A = A
        # ---- InterpolateDG: coordinate matching of boundary dofs against cell dofs ----------------------
        cell_xyz = self.xyz[self.tets[fcell]]  # (nF,4,3) DG1 dof coordinates of the owning cell
        sub_xyz = self.xyz[bcp]                # (nF,3,3) boundary-cell dof coordinates
        copy = np.full((self.nF, 3), -1, dtype=np.int64)
        for dof in range(4):                   # `for dof in closure_dofs`
            taken = np.zeros(self.nF, dtype=bool)
            for j in range(3):                 # `for j, sub_coord in enumerate(...)`: first match, then break
                close = np.all(np.abs(cell_xyz[:, dof] - sub_xyz[:, j]) <= 1e-8 + 1e-5 * np.abs(sub_xyz[:, j]),
                               axis=1) & ~taken
                copy[close, j] = dof
                taken |= close
        if np.any(copy < 0):
            raise RuntimeError("InterpolateDG: unmatched boundary dof")
        self.maps = WallMaps(order, facets, fcell, flocal, bvp, btopo, bcp, copy.astype(np.int8), cn,
                             wall_cells, counts)

    # ----------------------------------------------------------------------------------------------
    def _traction_at(self, u_cell: np.ndarray, lam_pts: np.ndarray) -> np.ndarray:
        """Ft at barycentric facet points.  ``u_cell`` (nF, ndof, 3), ``lam_pts`` (nF, nq, 4) -> (nF,nq,3)."""
        glam = self.glam[self.facet_wall]  # (nF,4,3)
        if self.order == 2:
            gphi = _p2_basis_gradients(lam_pts, glam[:, None, :, :])  # (nF,nq,10,3)
        else:
            gphi = np.broadcast_to(glam[:, None, :, :], lam_pts.shape[:2] + (4, 3))
        G = np.einsum("fdi,fqdj->fqij", u_cell, gphi)              # grad(u)_ij = d u_i / d x_j
        sigma = self.mu * (G + np.swapaxes(G, 2, 3))               # 2*mu*sym(grad(u))
        n = self.normal[:, None, :]
        F = -np.einsum("fqij,fqj->fqi", sigma, np.broadcast_to(n, G.shape[:2] + (3,)))
        Fn = np.einsum("fqi,fqi->fq", F, np.broadcast_to(n, F.shape))
        return F - Fn[..., None] * n

    def __call__(self, u_vec: np.ndarray, comp_offset: Tuple[int, int, int], node_stride: int = 1) -> np.ndarray:
        """tau (nF, 3 boundary dofs, 3 components) for one snapshot vector."""
        m = self.maps
        cn = m.cell_nodes[m.facet_cell]  # (nF, ndof)
        u_cell = np.stack([u_vec[comp_offset[c] + node_stride * cn] for c in range(3)], axis=-1)
        # b = assemble(inner(Ft, v)*ds)
        lam = np.zeros((self.nF, 3, 4))
        for q in range(3):
            np.put_along_axis(lam[:, q, :], self.facet_lv, np.broadcast_to(_Q2_PTS[q], (self.nF, 3)), axis=1)
        Ft = self._traction_at(u_cell, lam)  # (nF,3q,3)
        contrib = np.einsum("q,f,fqa,fqc->fac", _Q2_WTS, self.area, lam, Ft)  # (nF,4,3)
        b = np.zeros((len(m.wall_cells), 4, 3))
        np.add.at(b, self.facet_wall, contrib)
        x = np.linalg.solve(self.A, b)  # LUSolver: block diagonal per cell (and per component)
        # InterpolateDG.__call__
        xc = x[self.facet_wall]  # (nF,4,3)
        return np.take_along_axis(xc, m.bcell_local.astype(np.int64)[:, :, None], axis=1)


def project_dg_norm(w: np.ndarray, area: np.ndarray) -> np.ndarray:
    """``project_dg(inner(w, w) ** (1/2), DG1)`` per boundary triangle.  ``w`` (nF,3 dofs,3 comps) -> (nF,3)."""
    wq = np.einsum("qj,fjc->fqc", Q5_PTS, w)
    mag = np.sqrt(np.einsum("fqc,fqc->fq", wq, wq))
    rhs = area[:, None] * np.einsum("q,qi,fq->fi", Q5_WTS, Q5_PTS, mag)
    M = (area / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))
    return np.linalg.solve(M, rhs[..., None])[..., 0]  # LocalSolver


# --------------------------------------------------------------------------------------------------
# time loop and final formulas
# --------------------------------------------------------------------------------------------------
def run_time_loop(stress: SurfaceStress, snapshots, dt: float, comp_offset, node_stride: int = 1,
                  tau_prev: Optional[np.ndarray] = None, keep_wss: bool = False) -> Dict[str, np.ndarray]:
    """Sequential loop of ``compute_hemodynamics.py:271-318`` over an iterable of snapshot vectors.

    Returns the *un-normalised* sums so that time shards can be added before :func:`finalize`.
    ``tau_prev`` is zero at the global first snapshot (``:244``).
    """
    nF = stress.nF
    wss_sum = np.zeros((nF, 3, 3))
    tawss_sum = np.zeros((nF, 3))
    twssg_sum = np.zeros((nF, 3))
    prev = np.zeros((nF, 3, 3)) if tau_prev is None else tau_prev.copy()
    series = []
    count = 0
    for u in snapshots:
        tau = stress(np.asarray(u, dtype=np.float64), comp_offset, node_stride)
        if keep_wss:
            series.append(tau.copy())
        tawss_sum += np.linalg.norm(tau, axis=2)
        wss_sum += tau
        twssg_sum += project_dg_norm((tau - prev) / dt, stress.area)
        prev = tau
        count += 1
    out = {"wss_sum": wss_sum, "tawss_sum": tawss_sum, "twssg_sum": twssg_sum, "count": count, "tau_last": prev}
    if keep_wss:
        out["wss"] = np.stack(series) if series else np.zeros((0, nF, 3, 3))
    return out


def finalize(wss_sum: np.ndarray, tawss_sum: np.ndarray, twssg_sum: np.ndarray, count: int
             ) -> Dict[str, np.ndarray]:
    with np.errstate(divide="ignore", invalid="ignore"):
        twssg = twssg_sum / count
        tawss = tawss_sum / count
        mean_mag = np.linalg.norm(wss_sum / count, axis=2)
        rrt = 1.0 / mean_mag
        osi = 0.5 * (1.0 - mean_mag / tawss)
        ecap = osi / tawss
    return {"TAWSS": tawss, "OSI": osi, "RRT": rrt, "ECAP": ecap, "TWSSG": twssg}


def check_osi(osi: np.ndarray, tol: float = 1e-12) -> None:
    lo, hi = float(np.min(osi)), float(np.max(osi))
    assert -tol <= lo < 0.5, "OSI min is not within 0 to 0.5"
    assert -tol < hi <= 0.5 + tol, "OSI max is not within 0 to 0.5"
