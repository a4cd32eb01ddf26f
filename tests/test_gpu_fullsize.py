"""GPU parity at BASELINE.json mesh sizes and the edge cases of the time loop.

At full mesh size the CUDA path is checked (a) directly against the threaded C twin of the oracle on a handful of
snapshots, bit-exact maps included, and (b) over the workload's whole snapshot count through size-independent
properties of the path: homogeneity in the velocity, additivity over time shards with a halo snapshot, invariance
under the batch split, and exact structure under pure flow reversal.  Tolerance: relative L2 <= 1e-10 per field
(north_star); index maps bit-exact.
"""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _synthetic_case(n, m, order, n_snap, bulge=0.0, stenosis=0.0, seed=1234):
    from vasp_b200 import synth
    mesh = synth.vessel_mesh(n, m, radius=2.0e-3, stenosis=stenosis, bulge=bulge, seed=seed)
    return H.make_case(mesh["xyz"], mesh["tets"], order, n_snap=n_snap, seed=seed, period=0.951)


def _c_oracle(case, mu, n_snap, threads=None):
    from oracle import c_oracle, hemo_oracle as ho
    S = H.oracle_stress(case, mu)
    co = c_oracle.COracle(S)
    n = case["n_nodes"]
    res = co.run(case["u"][:n_snap], case["dt"], (0, n, 2 * n), threads=threads or c_oracle.max_threads())
    return S, res, ho.finalize(res["wss_sum"], res["tawss_sum"], res["twssg_sum"], res["count"])


def test_aneurysm_size_p1_against_oracle_and_properties(engine_lib):
    """BASELINE.json configs[2] mesh size (2.0 M tets, 73 k exterior facets), P1."""
    n_snap = 186
    case = _synthetic_case(40, 208, 1, n_snap, bulge=0.6)
    mu = 3.5e-3
    S, res, fin = _c_oracle(case, mu, 8)
    eng = H.engine_for(case, mu)
    m = eng.maps()
    assert np.array_equal(m["facets"], S.maps.facets)
    assert np.array_equal(m["facet_cell"], S.maps.facet_cell)
    assert np.array_equal(m["facet_local"], S.maps.facet_local)
    assert np.array_equal(m["btopology"], S.maps.btopology)
    assert np.array_equal(m["bvert_parent"], S.maps.bvert_parent)
    assert np.array_equal(m["bcell_local"], S.maps.bcell_local)
    assert np.array_equal(m["facet_nodes"], S.maps.cell_nodes[S.maps.facet_cell])
    # facet numbering is dolfin's: lexicographic in the sorted vertex triple
    f = m["facets"].astype(np.int64)
    key = (f[:, 0] * (f.max() + 1) + f[:, 1]) * (f.max() + 1) + f[:, 2]
    assert np.all(np.diff(key) > 0)
    eng.push(case["u"][:8], flags=1)
    out = eng.finalize()
    for name in H.FIELDS:
        assert H.rel_l2(out[name], fin[name]) < TOL, name
    assert H.rel_l2(eng.tau_last(), res["tau_last"]) < TOL

    # --- the whole series: properties ---------------------------------------------------------------------------
    u = case["u"]
    eng.begin(mu, case["dt"])
    eng.push(u, flags=1)
    whole, cnt = eng.sums()
    base = eng.finalize()
    assert cnt == n_snap
    osi = base["OSI"]
    assert np.nanmin(osi) >= -1e-12 and np.nanmax(osi) <= 0.5 + 1e-12  # compute_hemodynamics.py:366-372
    # homogeneity: u -> a u scales every sum by |a| (the mean vector by a), leaves OSI, scales RRT and ECAP by 1 / |a|
    a = -2.5
    eng.begin(mu, case["dt"])
    eng.push(a * u, flags=1)
    scaled = eng.finalize()
    assert H.rel_l2(scaled["TAWSS"], abs(a) * base["TAWSS"]) < TOL
    assert H.rel_l2(scaled["TWSSG"], abs(a) * base["TWSSG"]) < TOL
    assert H.rel_l2(scaled["OSI"], base["OSI"]) < 1e-9       # 1 - ratio: cancellation amplifies round-off
    assert H.rel_l2(scaled["RRT"], base["RRT"] / abs(a)) < TOL
    # additivity over time shards: three shards, each later one seeded by a halo snapshot, add up to the whole
    cuts = [0, 61, 130, n_snap]
    total = np.zeros_like(whole)
    n_total = 0
    for k in range(3):
        lo, hi = cuts[k], cuts[k + 1]
        eng.begin(mu, case["dt"])
        if k == 0:
            eng.push(u[lo:hi], flags=1)
        else:
            eng.push(u[lo - 1:hi], flags=2)
        s, c = eng.sums()
        total += s
        n_total += c
    assert n_total == n_snap
    assert H.rel_l2(total, whole) < 1e-12
    # the batch split does not change the result beyond summation order
    eng.set_tuning(17, 0)
    eng.begin(mu, case["dt"])
    eng.push(u, flags=1)
    s17, _ = eng.sums()
    assert H.rel_l2(s17, whole) < 1e-12
    eng.close()


def test_p2_mid_size_against_oracle_and_reversal(engine_lib):
    """P2 (the reference-native order) on a 0.41 M-tet vessel: direct parity, then pure flow reversal."""
    case = _synthetic_case(24, 120, 2, 6, stenosis=0.3)
    mu = 3.5e-3
    S, res, fin = _c_oracle(case, mu, 6)
    eng = H.engine_for(case, mu)
    assert np.array_equal(eng.maps()["facet_nodes"], S.maps.cell_nodes[S.maps.facet_cell])
    eng.push(case["u"], flags=1)
    out = eng.finalize()
    for name in H.FIELDS:
        assert H.rel_l2(out[name], fin[name]) < TOL, name
    # u(t_k) = (-1)^k u_0: the mean WSS vanishes for an even count => OSI = 1/2, RRT = inf (reference :344-346),
    # TAWSS = |tau_0|, and every step but the first contributes P(|2 tau_0|) / dt to TWSSG
    u0 = case["u"][:1]
    signs = np.array([1.0, -1.0] * 4)[:, None]
    eng.begin(mu, case["dt"])
    wss = eng.push(signs * u0, flags=1, keep_wss=True)
    rev = eng.finalize()
    mag = np.linalg.norm(wss[0], axis=2)
    assert H.rel_l2(rev["TAWSS"], mag) < TOL
    assert np.all(np.abs(rev["OSI"] - 0.5) < 1e-12)
    assert np.all(np.isinf(rev["RRT"]) | (rev["RRT"] > 1e12 / np.maximum(mag, 1e-300)))
    eng.begin(mu, case["dt"])
    eng.push(u0, flags=1)
    one = eng.finalize(1)                                    # TWSSG of a single step from rest: P(|tau_0|) / dt
    assert H.rel_l2(rev["TWSSG"], one["TWSSG"] * (1 + 2 * 7) / 8) < TOL
    eng.close()


@pytest.mark.parametrize("order", [2, 1])
def test_edge_cases_of_the_time_loop(engine_lib, order):
    """Single snapshot, a count that is not a multiple of the 32 / 64-column passes, one-snapshot pushes, and the
    error paths of the C ABI (no CPU fallback: errors are loud)."""
    from vasp_b200.engine import VaspHemoError
    src = H.load_fluid("cylinder")
    case = H.make_case(src["xyz"], src["tets"], order, n_snap=67)
    mu = 1.3
    _, res, fin = H.oracle_run(case, mu)
    u = case["u"]
    eng = H.engine_for(case, mu)
    # one snapshot at a time: every push is a 1-column block continued through tau_last
    eng.push(u[:1], flags=1)
    for k in range(1, 67):
        eng.push(u[k:k + 1])
    out = eng.finalize()
    for name in H.FIELDS:
        assert H.rel_l2(out[name], fin[name]) < TOL, name
    # a single snapshot in the whole loop
    _, res1, fin1 = H.oracle_run({**case, "u": u[:1]}, mu)
    eng.begin(mu, case["dt"])
    eng.push(u[:1], flags=1)
    out = eng.finalize()
    for name in H.FIELDS:
        assert H.rel_l2(out[name], fin1[name]) < TOL, name
    # strided rows (a view into a wider buffer) take the pitched-copy path
    wide = np.zeros((67, u.shape[1] + 5))
    wide[:, :u.shape[1]] = u
    eng.begin(mu, case["dt"])
    eng.push(wide[:, :u.shape[1]], flags=1)
    out = eng.finalize()
    for name in H.FIELDS:
        assert H.rel_l2(out[name], fin[name]) < TOL, name
    # errors: continuation without a first push, halo push with nothing after the halo, short vectors
    eng.begin(mu, case["dt"])
    with pytest.raises(VaspHemoError):
        eng.push(u[:3])
    with pytest.raises(VaspHemoError):
        eng.push(u[:1], flags=2)
    with pytest.raises(ValueError):
        eng.push(u[:, :-1], flags=1)
    eng.close()
