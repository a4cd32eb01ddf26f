"""Properties the continuous problem has and the restated reference algorithm must inherit (checks on the checker):
linearity and Galilean invariance of the traction, covariance under rigid motions of mesh + velocity, independence of
the vertex / cell numbering.  None of them needs a second implementation to compare with."""
import numpy as np
import pytest

from tests import helpers as H

MU = 0.8


def _case(order, n_snap=5, seed=3, mesh="cylinder"):
    src = H.load_fluid(mesh)
    return H.make_case(src["xyz"], src["tets"], order, n_snap=n_snap, seed=seed)


def _tau(case, u):
    S = H.oracle_stress(case, MU)
    n = case["n_nodes"]
    return S, np.stack([S(v, (0, n, 2 * n)) for v in np.atleast_2d(u)])


@pytest.mark.parametrize("order", [1, 2])
def test_traction_is_linear_and_blind_to_a_uniform_velocity(order):
    case = _case(order)
    u = case["u"]
    S, tau = _tau(case, u[:3])
    a, b = 1.7, -0.4
    _, mix = _tau(case, a * u[0] + b * u[1])
    scale = np.abs(tau).max()
    assert np.abs(mix[0] - (a * tau[0] + b * tau[1])).max() < 1e-13 * scale
    n = case["n_nodes"]
    shift = np.concatenate([np.full(n, 3.0), np.full(n, -2.0), np.full(n, 0.5)])
    _, moved = _tau(case, u[2] + shift)
    # a constant of size ~3 rides through differences of O(1/h) gradients: round-off relative to |shift| / h
    assert np.abs(moved[0] - tau[2]).max() < 1e-11 * scale


@pytest.mark.parametrize("order", [1, 2])
def test_rigid_motion_of_mesh_and_velocity_rotates_the_traction(order):
    case = _case(order, n_snap=4)
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1                                   # a proper rotation keeps the orientation of every cell
    shift = np.array([0.3, -1.1, 2.0])
    n = case["n_nodes"]
    moved = dict(case, xyz=case["xyz"] @ q.T + shift, points=case["points"] @ q.T + shift)
    u3 = case["u"].reshape(len(case["u"]), 3, n)
    moved["u"] = np.einsum("ij,sjn->sin", q, u3).reshape(len(u3), 3 * n)
    S0, res0, fin0 = H.oracle_run(case, MU, keep_wss=True)
    S1, res1, fin1 = H.oracle_run(moved, MU, keep_wss=True)
    assert np.array_equal(S0.maps.facets, S1.maps.facets)                     # connectivity did not change
    want = np.einsum("ij,sfkj->sfki", q, res0["wss"])
    assert H.rel_l2(res1["wss"], want) < 1e-11          # the shift costs a digit in the edge vectors
    for name in H.FIELDS:                                                     # scalars are invariant
        assert H.rel_l2(fin1[name], fin0[name]) < 1e-10, name


@pytest.mark.parametrize("order", [1, 2])
def test_fields_do_not_depend_on_the_numbering(order):
    """Permute vertex ids and cell order: facet numbers, boundary-vertex numbers and dof order all change, the field
    as a function of position must not.  Values are compared per (facet, vertex) in the original vertex ids."""
    case = _case(order, n_snap=4, mesh="stenosis")
    rng = np.random.default_rng(7)
    nv = len(case["xyz"])
    perm = rng.permutation(nv)                          # new id of old vertex v is perm[v]
    xyz2 = np.empty_like(case["xyz"])
    xyz2[perm] = case["xyz"]
    tets2 = perm[case["tets"]][rng.permutation(len(case["tets"]))]
    other = dict(case, xyz=xyz2, tets=tets2)            # the velocity nodes (points, u) keep their own numbering
    if order == 1:
        other["points"] = xyz2
        n = nv
        u3 = case["u"].reshape(len(case["u"]), 3, n)
        u2 = np.empty_like(u3)
        u2[:, :, perm] = u3
        other["u"] = u2.reshape(len(u3), 3 * n)
    S0, res0, fin0 = H.oracle_run(case, MU)
    S1, res1, fin1 = H.oracle_run(other, MU)
    assert S0.nF == S1.nF and not np.array_equal(S0.maps.facets, S1.maps.facets)

    inv = np.empty(nv, dtype=np.int64)
    inv[perm] = np.arange(nv)                           # old id of new vertex

    def keyed(S, to_old, field):
        v = to_old[S.maps.bcell_parent.astype(np.int64)]            # (nF, 3): original id of every boundary dof's vertex
        tri = np.sort(v, axis=1)[:, None, :].repeat(3, axis=1)      # the facet, as a sorted triple of original ids
        key = np.concatenate([tri, v[:, :, None]], axis=2).reshape(-1, 4)
        order_ = np.lexsort(key.T[::-1])
        return key[order_], field.reshape(-1)[order_]

    for name in ("TAWSS", "OSI", "TWSSG"):
        k0, v0 = keyed(S0, np.arange(nv), fin0[name])
        k1, v1 = keyed(S1, inv, fin1[name])
        assert np.array_equal(k0, k1)
        assert H.rel_l2(v1, v0) < 1e-10, name
