"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.

Tolerance (north_star): relative L2 <= 1e-10 per field for TAWSS/OSI/RRT/ECAP (also checked: TWSSG, per-step WSS);
index maps bit-exact."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _check_maps(eng, S):
    m, o = eng.maps(), S.maps
    assert np.array_equal(m["facets"], o.facets)
    assert np.array_equal(m["facet_cell"], o.facet_cell)
    assert np.array_equal(m["facet_local"], o.facet_local)
    assert np.array_equal(m["bcell_parent"], o.bcell_parent)
    assert np.array_equal(m["btopology"], o.btopology)
    assert np.array_equal(m["bvert_parent"], o.bvert_parent)
    assert np.array_equal(m["bcell_local"], o.bcell_local)
    assert np.array_equal(m["facet_nodes"], o.cell_nodes[o.facet_cell])
    assert eng.n_wall_cells == len(o.wall_cells)
    assert eng.n_multi == int(np.sum(o.n_ext[S.facet_wall] >= 2))


@pytest.mark.parametrize("order", [2, 1])
@pytest.mark.parametrize("mesh", ["pipe", "cylinder", "stenosis", "aneurysm"])
def test_maps_and_fields_match_oracle(engine_lib, mesh, order):
    src = H.load_pipe() if mesh == "pipe" else H.load_fluid(mesh)
    case = H.make_case(src["xyz"], src["tets"], order, n_snap=7)
    mu = 3.5e-3
    S, res, fin = H.oracle_run(case, mu, keep_wss=True)
    eng = H.engine_for(case, mu)
    _check_maps(eng, S)
    g = eng.geometry()
    assert H.rel_l2(g["area"], S.area) < 1e-13
    assert H.rel_l2(g["normal"], S.normal) < 1e-13
    wss = eng.push(case["u"], flags=1, keep_wss=True)
    assert wss.shape == res["wss"].shape
    for k in range(wss.shape[0]):
        assert H.rel_l2(wss[k], res["wss"][k]) < TOL, f"WSS step {k}"
    out = eng.finalize()
    for name in H.FIELDS:
        assert H.rel_l2(out[name], fin[name]) < TOL, name
    sums, cnt = eng.sums()
    assert cnt == 7
    assert H.rel_l2(sums[:9].reshape(3, 3, -1).transpose(2, 0, 1), res["wss_sum"]) < TOL
    eng.close()


@pytest.mark.parametrize("order", [2, 1])
def test_chunking_batching_and_halo_are_equivalent(engine_lib, order):
    src = H.load_fluid("stenosis")
    case = H.make_case(src["xyz"], src["tets"], order, n_snap=23)
    mu = 1.0
    _, res, fin = H.oracle_run(case, mu)
    u = case["u"]
    for batch, chunk in ((0, 0), (5, 1), (4, 3), (23, 23), (7, 2)):
        eng = H.engine_for(case, mu)
        eng.set_tuning(batch, chunk)
        eng.push(u[:9], flags=1)
        eng.push(u[9:10])          # continuation through tau_last
        eng.push(u[10:])
        out = eng.finalize()
        for name in H.FIELDS:
            assert H.rel_l2(out[name], fin[name]) < TOL, (name, batch, chunk)
        assert H.rel_l2(eng.tau_last(), res["tau_last"]) < TOL
        eng.close()
    # time shards with a halo snapshot add up to the sequential result
    a, b = H.engine_for(case, mu), H.engine_for(case, mu)
    a.push(u[:11], flags=1)
    b.push(u[10:], flags=2)        # u[10] only seeds tau_prev
    sa, ca = a.sums()
    sb, cb = b.sums()
    assert (ca, cb) == (11, 12)
    a.set_sums(sa + sb, ca + cb)
    out = a.finalize()
    for name in H.FIELDS:
        assert H.rel_l2(out[name], fin[name]) < TOL, name
    a.close(), b.close()


@pytest.mark.parametrize("order", [2, 1])
def test_segments_passes_and_column_blocks(engine_lib, order):
    """100 snapshots: several 31-column lane passes per segment, several segments, several staged column blocks
    (device-resident push), per-step WSS and tau_last all agree with the sequential oracle."""
    src = H.load_fluid("cylinder")
    case = H.make_case(src["xyz"], src["tets"], order, n_snap=100)
    mu = 2.0
    _, res, fin = H.oracle_run(case, mu, keep_wss=True)
    u = np.ascontiguousarray(case["u"])
    for batch, chunk in ((0, 0), (0, 31), (0, 62), (40, 0), (33, 93)):
        eng = H.engine_for(case, mu)
        eng.set_tuning(batch, chunk)
        d_u = eng.device_alloc(u.nbytes)
        eng.h2d(d_u, u)
        nF = eng.nF
        d_w = eng.device_alloc(100 * nF * 72)
        eng.push_device(d_u, 100, u.shape[1] * 8, flags=1, d_wss=d_w)
        wss = np.empty((100, nF, 3, 3))
        eng.d2h(wss, d_w)
        assert H.rel_l2(wss, res["wss"]) < TOL, (batch, chunk)
        out = eng.finalize()
        for name in H.FIELDS:
            assert H.rel_l2(out[name], fin[name]) < TOL, (name, batch, chunk)
        assert H.rel_l2(eng.tau_last(), res["tau_last"]) < TOL
        # the same through the host path, split in two pushes with a halo start
        eng.begin(mu, case["dt"])
        eng.push(u[:37], flags=1)
        eng.push(u[37:])
        out = eng.finalize()
        for name in H.FIELDS:
            assert H.rel_l2(out[name], fin[name]) < TOL, (name, batch, chunk, "host")
        eng.device_free(d_u)
        eng.device_free(d_w)
        eng.close()


def test_poiseuille_known_answer(engine_lib):
    """The reference's own pin (tests/test_compute_hemodynamics.py:68-88): wall-averaged TAWSS in (1.95, 2.05) for
    G=4, mu=1, R=1, and OSI within [-1e-12, 0.5 + 1e-12]."""
    from oracle import hemo_oracle as ho
    src = H.load_pipe()
    xyz, tets = src["xyz"], src["tets"]
    from vasp_b200 import synth
    pts, edges, _ = synth.p2_points(xyz, tets, seed=11)
    n = len(pts)
    u0 = np.concatenate([1.0 - pts[:, 1] ** 2 - pts[:, 2] ** 2, np.zeros(2 * n)])
    from vasp_b200.engine import HemoEngine
    eng = HemoEngine(0)
    eng.set_mesh(xyz, tets)
    eng.set_velocity_layout(2, refined_xyz=pts)
    eng.begin(1.0, 0.1)
    eng.push(np.tile(u0, (9, 1)), flags=1)
    out = eng.finalize()
    m, g = eng.maps(), eng.geometry()
    fx = xyz[m["facets"]][:, :, 0]
    wall = (fx > 0.1).all(axis=1) & (fx < 4.9).all(axis=1)   # boundary cells whose vertices are all inside
    avg = float((out["TAWSS"].mean(axis=1) * g["area"])[wall].sum() / g["area"][wall].sum())
    assert 1.95 < avg < 2.05
    assert abs(avg - 1.99490423) < 1e-7
    ho.check_osi(out["OSI"])
    eng.close()


def test_rigid_rotation_gives_zero_traction(engine_lib):
    src = H.load_fluid("aneurysm")
    from vasp_b200 import synth
    pts, _, _ = synth.p2_points(src["xyz"], src["tets"], seed=3)
    w = np.array([0.3, -0.2, 0.5])
    ur = np.cross(w, pts)
    u = np.concatenate([ur[:, 0], ur[:, 1], ur[:, 2]])[None, :]
    from vasp_b200.engine import HemoEngine
    eng = HemoEngine(0)
    eng.set_mesh(src["xyz"], src["tets"])
    eng.set_velocity_layout(2, refined_xyz=pts)
    eng.begin(1.0, 1.0)
    wss = eng.push(u, flags=1, keep_wss=True)
    scale = np.abs(ur).max() / np.ptp(pts, axis=0).max()
    assert np.abs(wss).max() < 1e-9 * scale
    eng.close()
