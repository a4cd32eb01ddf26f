"""GPU parity on the meshes BASELINE.json quotes its targets on: configs[3] (AVF size, 5.0 M tets, P2) and configs[4]
(10.0 M tets, P2) -- the workloads ``bench.py`` measures, built by the same recipe (``bench.WORKLOADS``).

Per mesh: every index map of K0 bit-exact against the oracle; the five fields and the running sums of a handful of
whole snapshot vectors within 1e-10 of the threaded C twin of the oracle, through the whole-vector route (K1 gathers
the wall layer out of 163 / 325 MB vectors) and through the compact route (host gather in front of the bus), which
must agree with each other bitwise; and a device-resident push long enough to be cut into several column blocks of the
staged block W (K2 addresses W with 32-bit element offsets; the block is sized to stay below 2^31 elements), checked
against the oracle over the same series.
"""
import numpy as np
import pytest

import bench
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _compact_oracle(g, eng, stress, co):
    """Let the C oracle read compact blocks: same values, the container the GPU was given (test plumbing only)."""
    slots = eng.wall_slots()
    remap = np.searchsorted(slots, stress.maps.cell_nodes[stress.maps.wall_cells])
    assert np.array_equal(slots[remap], stress.maps.cell_nodes[stress.maps.wall_cells])
    co._keep["wall_nodes"] = np.ascontiguousarray(remap, dtype=np.int64)
    co._maps.wall_nodes = co._keep["wall_nodes"].ctypes.data
    nwp = eng.compact_len // 3
    return (0, nwp, 2 * nwp)


@pytest.mark.parametrize("name,n_long", [("avf_p2", 400), ("vessel10m_p2", 300)])
def test_baseline_p2_meshes_against_the_oracle(engine_lib, name, n_long):
    from oracle import hemo_oracle as ho
    from vasp_b200 import synth
    from vasp_b200.engine import HemoEngine, pinned_empty
    g = bench.build_geometry(name)
    stress, co, threads = bench.oracle_for(g)
    eng = HemoEngine(0)
    eng.set_mesh(g["xyz"], g["tets"])
    eng.set_velocity_layout(2, refined_xyz=g["points"])
    eng.set_host_compaction("auto", 8)   # the automatic rule looks at the thread count; do not depend on the box's cores
    m, S = eng.maps(), stress.maps
    assert len(g["tets"]) > (9.9e6 if name == "vessel10m_p2" else 4.9e6)
    for key, want in (("facets", S.facets), ("facet_cell", S.facet_cell), ("facet_local", S.facet_local),
                      ("bcell_parent", S.bcell_parent), ("btopology", S.btopology), ("bvert_parent", S.bvert_parent),
                      ("bcell_local", S.bcell_local), ("facet_nodes", S.cell_nodes[S.facet_cell])):
        assert np.array_equal(m[key], want), key
    assert eng.compaction_active and eng.n_wall_nodes == len(np.unique(S.cell_nodes[S.wall_cells]))

    # ---- a handful of whole vectors: oracle, whole-vector route, compact route ------------------------------------
    n, n_s = len(g["points"]), 6
    coef, dt = bench.series_coefficients(n_s, 0, 1000)
    u = pinned_empty((n_s, 3 * n))
    synth.velocity_series(g["basis"], coef, out=u)
    res = co.run(u, dt, (0, n, 2 * n), threads=threads)
    fin = ho.finalize(res["wss_sum"], res["tawss_sum"], res["twssg_sum"], res["count"])
    want_sums = np.concatenate([res["wss_sum"].reshape(-1, 9).T, res["tawss_sum"].T, res["twssg_sum"].T])
    got = {}
    for mode in ("off", "on"):
        eng.set_host_compaction(mode, 8)
        eng.begin(bench.MU, dt)
        eng.push(u, flags=1)
        sums, cnt = eng.sums()
        out = eng.finalize()
        assert cnt == n_s and H.rel_l2(sums, want_sums) < TOL
        for k in H.FIELDS:
            assert H.rel_l2(out[k], fin[k]) < TOL, (mode, k)
        assert H.rel_l2(eng.tau_last(), res["tau_last"]) < TOL
        got[mode] = sums
        assert eng.io_stats()["h2d_bytes"] == n_s * 8 * (eng.compact_len if mode == "on" else 3 * n)
    assert np.array_equal(got["on"], got["off"])
    del u

    # ---- a long device-resident push of compact blocks: several column blocks of W -----------------------------------
    slots, nwp = eng.wall_slots(), eng.compact_len // 3
    idx = np.concatenate([slots, np.full(nwp - len(slots), slots[-1])])
    flat = np.ascontiguousarray(g["basis"][:, :, idx]).reshape(synth.N_MODES, 3 * nwp)
    coef, dt = bench.series_coefficients(n_long, 0, n_long)
    d = eng.device_alloc(n_long * 3 * nwp * 8)
    off = _compact_oracle(g, eng, stress, co)
    acc, prev = None, None
    for a in range(0, n_long, 50):
        rows = coef[a:a + 50] @ flat
        eng.h2d(d + a * 3 * nwp * 8, rows)
        r = co.run(rows, dt, off, tau_prev=prev, threads=threads)
        prev = r["tau_last"]
        acc = r if acc is None else {k: (acc[k] + r[k] if k != "tau_last" else r[k]) for k in r}
    eng.set_tuning(batch_snapshots=127)   # W holds at most 128 columns: the push below needs three or four blocks
    launches0 = eng.timers()["launches"]
    eng.begin(bench.MU, dt)
    eng.push_compact_device(d, n_long, 3 * nwp * 8, flags=1)
    blocks = (eng.timers()["launches"] - launches0) // 3          # K1 + K2 + K3 per column block
    assert blocks >= 2, "the push was meant to be cut into several column blocks"
    out = eng.finalize()
    fin = ho.finalize(acc["wss_sum"], acc["tawss_sum"], acc["twssg_sum"], n_long)
    for k in H.FIELDS:
        assert H.rel_l2(out[k], fin[k]) < TOL, k
    assert H.rel_l2(eng.tau_last(), acc["tau_last"]) < TOL
    osi = out["OSI"]
    assert np.nanmin(osi) >= -1e-12 and np.nanmax(osi) <= 0.5 + 1e-12      # compute_hemodynamics.py:366-372
    eng.device_free(d)
    eng.close()
