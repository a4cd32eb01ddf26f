"""Degenerate meshes: cells that own 2, 3 or all 4 of their facets on the boundary (the
SurfaceProjector block of such a cell couples its facets through the shared dofs; with four exterior facets no
row of the block is an identity row any more).  The four reference meshes only contain cells with two."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-10

UNIT = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
CUBE = np.array([[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1)], dtype=np.float64)

MESHES = {
    # one tetrahedron: 4 exterior facets in one cell
    "one_tet": (UNIT * [1.0, 1.3, 0.7] + 0.1, np.array([[0, 1, 2, 3]])),
    # two tetrahedra glued at a face: 3 exterior facets each
    "two_tets": (np.vstack([UNIT, [[0.9, 0.8, 0.7]]]), np.array([[0, 1, 2, 3], [1, 2, 3, 4]])),
    # cube cut into 5: four corner cells with 3 exterior facets, the middle one with none
    "cube5": (CUBE, np.array([[0, 1, 2, 4], [3, 1, 2, 7], [5, 1, 4, 7], [6, 2, 4, 7], [1, 2, 4, 7]])),
    # Kuhn split of the cube into 6: every cell has 2 exterior facets
    "cube6": (CUBE, np.array([[0, 1, 3, 7], [0, 1, 5, 7], [0, 2, 3, 7], [0, 2, 6, 7], [0, 4, 5, 7], [0, 4, 6, 7]])),
}


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("name", sorted(MESHES))
def test_cells_with_many_exterior_facets(engine_lib, name, order):
    xyz, tets = MESHES[name]
    tets = tets.astype(np.int64)
    case = H.make_case(xyz, tets, order, n_snap=37, seed=11)
    # the synthetic basis vanishes on a unit-radius wall; add a field with gradients everywhere
    rng = np.random.default_rng(5)
    p = case["points"]
    n = len(p)
    A = rng.normal(size=(3, 3))
    B = rng.normal(size=(3, 3, 3)) if order == 2 else np.zeros((3, 3, 3))
    t = np.arange(37) * case["dt"]
    base = p @ A.T + np.einsum("ijk,nj,nk->ni", B, p, p)                  # (n, 3), exactly representable
    u = np.stack([np.concatenate([(base[:, c] * (1 + 0.5 * np.sin(7 * tk + c))) for c in range(3)]) for tk in t])
    case["u"] = u + case["u"]
    mu = 0.9
    S, res, fin = H.oracle_run(case, mu, keep_wss=True)
    counts = np.bincount(S.maps.facet_cell, minlength=len(tets))
    assert counts.max() == {"one_tet": 4, "two_tets": 3, "cube5": 3, "cube6": 2}[name]
    eng = H.engine_for(case, mu)
    m = eng.maps()
    assert np.array_equal(m["facets"], S.maps.facets) and np.array_equal(m["facet_cell"], S.maps.facet_cell)
    wss = eng.push(case["u"], flags=1, keep_wss=True)
    out = eng.finalize()
    assert H.rel_l2(wss, res["wss"]) < TOL
    for f in H.FIELDS:
        assert H.rel_l2(out[f], fin[f]) < TOL, f
    eng.close()
