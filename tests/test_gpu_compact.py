"""Wall-layer compaction in front of the bus (include/vasp_hemo.h, csrc/compact.cu): moving only the dofs that can
reach a wall facet (the reference integrates over ds only, compute_hemodynamics.py:113-115) must not change a bit.

The compact route hands K2 exactly the staged block W the whole-vector route builds, so everything downstream is
required to be BITWISE equal between the two routes -- and, like every route, within 1e-10 of the oracle.
"""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _run(eng, case, mu, mode, u=None, flags=1, keep_wss=False):
    eng.set_host_compaction(mode, 8)   # 8 gather threads whatever the box has: the automatic rule looks at the count
    eng.begin(mu, case["dt"])
    wss = eng.push(case["u"] if u is None else u, flags=flags, keep_wss=keep_wss)
    sums, cnt = eng.sums()
    return sums, cnt, eng.finalize(), eng.tau_last(), (None if wss is None else np.array(wss))


@pytest.mark.parametrize("order", [2, 1])
@pytest.mark.parametrize("name", ["pipe", "stenosis"])
def test_compact_route_is_bitwise_the_whole_vector_route(engine_lib, name, order):
    src = H.load_pipe() if name == "pipe" else H.load_fluid(name)
    case = H.make_case(src["xyz"], src["tets"], order, n_snap=71)
    mu = 0.7
    _, res, fin = H.oracle_run(case, mu, keep_wss=True)
    eng = H.engine_for(case, mu)
    assert not eng.compaction_active or eng.n_wall_nodes <= 0.35 * eng.n_nodes  # small meshes: mostly wall layer
    s_off, c_off, f_off, t_off, w_off = _run(eng, case, mu, "off", keep_wss=True)
    s_on, c_on, f_on, t_on, w_on = _run(eng, case, mu, "on", keep_wss=True)
    assert c_on == c_off == 71
    assert np.array_equal(s_on, s_off) and np.array_equal(t_on, t_off) and np.array_equal(w_on, w_off)
    for k in H.FIELDS:
        assert np.array_equal(f_on[k], f_off[k]), k
        assert H.rel_l2(f_on[k], fin[k]) < TOL, k
    assert H.rel_l2(w_on, res["wss"]) < TOL
    st = eng.io_stats()
    assert st["h2d_bytes"] == 71 * 8 * eng.compact_len  # only the wall layer crossed the bus
    assert eng.compact_len == 3 * ((eng.n_wall_nodes + 31) // 32 * 32)

    # the host gather itself, against numpy with the exported slot table
    slots = eng.wall_slots()
    assert np.all(np.diff(slots) > 0)
    n = case["n_nodes"]
    c = eng.compact(case["u"][:5])
    nwp = eng.compact_len // 3
    for comp in range(3):
        assert np.array_equal(c[:, comp * nwp:comp * nwp + len(slots)], case["u"][:5, comp * n + slots])
        assert np.all(c[:, comp * nwp + len(slots):(comp + 1) * nwp] == case["u"][:5, comp * n + slots[-1]][:, None])

    # compact blocks pushed by the caller, from the host and resident in device memory; halo continuation
    cu = eng.compact(case["u"])
    eng.begin(mu, case["dt"])
    eng.push_compact(cu[:40], flags=1)
    eng.push_compact(cu[40:])
    s_c, _ = eng.sums()
    eng.begin(mu, case["dt"])
    eng.push(case["u"][:40], flags=1)
    eng.push(case["u"][40:])
    s_ref, _ = eng.sums()
    assert np.array_equal(s_c, s_ref)
    d = eng.device_alloc(cu.nbytes)
    eng.h2d(d, cu)
    eng.begin(mu, case["dt"])
    eng.push_compact_device(d + 29 * cu.strides[0], 71 - 29, cu.strides[0], flags=2)  # snapshot 29 is the halo
    s_halo, c_halo = eng.sums()
    eng.begin(mu, case["dt"])
    eng.set_host_compaction("off", 8)
    eng.push(case["u"][29:], flags=2)
    s_halo_ref, c_ref = eng.sums()
    # one launch over the resident block against the batches of a host push: summation order only
    assert c_halo == c_ref == 71 - 30 and H.rel_l2(s_halo, s_halo_ref) < 1e-12
    eng.device_free(d)
    eng.close()


def test_compaction_with_interleaved_layout_and_injected_node_perm(engine_lib):
    """The raw turtleFSI route: vectors of shape (N_all, 3), fluid nodes picked by node_perm."""
    src = H.load_fluid("cylinder")
    case = H.make_case(src["xyz"], src["tets"], 2, n_snap=37)
    n = case["n_nodes"]
    rng = np.random.default_rng(3)
    n_all = n + 311
    ids = np.sort(rng.choice(n_all, n, replace=False))
    raw = np.full((37, n_all, 3), 1.0e3)
    raw[:, ids, :] = case["u"].reshape(37, 3, n).transpose(0, 2, 1)
    raw = raw.reshape(37, 3 * n_all)
    from vasp_b200.engine import HemoEngine
    mu = 1.1
    outs = {}
    for mode in ("off", "on"):
        eng = HemoEngine(0)
        eng.set_mesh(case["xyz"], case["tets"])
        eng.set_velocity_layout(2, refined_xyz=case["points"], node_perm=ids, comp_offset=(0, 1, 2), node_stride=3)
        eng.set_host_compaction(mode, 4)
        eng.begin(mu, case["dt"])
        wss = np.array(eng.push(raw, flags=1, keep_wss=True))
        outs[mode] = (eng.sums()[0], wss, eng.finalize())
        eng.close()
    assert np.array_equal(outs["on"][0], outs["off"][0]) and np.array_equal(outs["on"][1], outs["off"][1])
    _, res, fin = H.oracle_run(case, mu, keep_wss=True)
    for k in H.FIELDS:
        assert H.rel_l2(outs["on"][2][k], fin[k]) < TOL, k
    assert H.rel_l2(outs["on"][1], res["wss"]) < TOL


def test_auto_mode_compacts_large_meshes_and_batches_split(engine_lib):
    """A mesh whose wall layer is a small share of the nodes takes the compact route on its own; pieces of the pinned
    ring, batches of the device stage and the column blocks of K1/K2 all split the push without changing the sums."""
    from vasp_b200 import synth
    mesh = synth.vessel_mesh(32, 40, radius=2.0e-3, seed=11)
    case = H.make_case(mesh["xyz"], mesh["tets"], 1, n_snap=150, seed=11)
    mu = 3.5e-3
    eng = H.engine_for(case, mu)
    eng.set_host_compaction("auto", 8)
    assert eng.compaction_active and eng.n_wall_nodes < 0.35 * eng.n_nodes
    s_auto, _, f_auto, _, _ = _run(eng, case, mu, "auto")
    assert eng.io_stats()["h2d_bytes"] == 150 * 8 * eng.compact_len
    s_off, _, f_off, _, _ = _run(eng, case, mu, "off")
    assert np.array_equal(s_auto, s_off)
    eng.set_tuning(batch_snapshots=23)
    s_b, _, _, _, _ = _run(eng, case, mu, "on")
    assert H.rel_l2(s_b, s_off) < 1e-12  # other segment boundaries: summation order only
    wide = np.zeros((150, case["u"].shape[1] + 3))
    wide[:, :case["u"].shape[1]] = case["u"]
    s_w, _, _, _, _ = _run(eng, case, mu, "on", u=wide[:, :case["u"].shape[1]])
    assert np.array_equal(s_w, s_b)
    eng.close()
