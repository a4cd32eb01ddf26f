"""GPU tests of the time-major WSS output: K2 writes tau as rows of the (dof x time) matrix of
VaSP's spectral tools; it must carry exactly the numbers of the per-snapshot vectors (bitwise) and match the oracle."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_cli import _make_folder
from vasp_b200 import synth, wss_matrix
from vasp_b200.engine import PUSH_GLOBAL_FIRST, PUSH_HALO_FIRST, pinned_empty

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("order,mesh", [(1, "stenosis"), (2, "aneurysm")])
def test_matrix_columns_are_the_per_snapshot_vectors(engine_lib, order, mesh):
    src = H.load_fluid(mesh)
    n_snap, mu = 75, 3.5e-3
    case = H.make_case(src["xyz"], src["tets"], order, n_snap=n_snap)
    S, res, fin = H.oracle_run(case, mu, keep_wss=True)
    eng = H.engine_for(case, mu)
    nF = eng.nF
    steps = eng.push(case["u"], flags=PUSH_GLOBAL_FIRST, keep_wss=True)          # (n, nF, 3, 3)
    ref_out = eng.finalize()
    # same series in ragged host batches (7 per H2D batch, pushes of 31 + 44), time-major
    eng.begin(mu, case["dt"])
    eng.set_tuning(batch_snapshots=7)
    M = pinned_empty((9 * nF, n_snap + 5))
    M[:] = -1.0
    eng.set_wss_matrix(M, first_column=2)
    assert eng.push(case["u"][:31], flags=PUSH_GLOBAL_FIRST) is M
    eng.push(case["u"][31:])
    out = eng.finalize()
    assert np.array_equal(M[:, 2:2 + n_snap], steps.reshape(n_snap, -1).T)       # bitwise the same tau
    assert (M[:, :2] == -1.0).all() and (M[:, 2 + n_snap:] == -1.0).all()        # nothing else is touched
    for name in H.FIELDS:
        assert np.array_equal(out[name], ref_out[name]) or H.rel_l2(out[name], ref_out[name]) < 1e-13
    want = np.asarray(res["wss"]).reshape(n_snap, -1).T
    assert H.rel_l2(M[:, 2:2 + n_snap], want) < 1e-10
    with pytest.raises(RuntimeError, match="WSS matrix has"):
        eng.push(case["u"][:4])                                                  # only 3 columns left
    with pytest.raises(ValueError):
        eng.push(case["u"][:2], keep_wss=True)
    # a time shard: halo snapshot first, columns start at 0 of the shard's own matrix
    eng.set_wss_matrix(None)
    eng.begin(mu, case["dt"])
    Ms = pinned_empty((9 * nF, 20))
    eng.set_wss_matrix(Ms)
    eng.push(case["u"][39:60], flags=PUSH_HALO_FIRST)
    assert np.array_equal(Ms, steps[40:60].reshape(20, -1).T)
    eng.set_wss_matrix(None)
    back = eng.push(case["u"][:3], flags=PUSH_GLOBAL_FIRST, keep_wss=True)       # the default layout is back
    assert np.array_equal(back, steps[:3])
    eng.close()


def test_device_resident_matrix(engine_lib):
    src = H.load_fluid("cylinder")
    n_snap, mu = 70, 1.0
    case = H.make_case(src["xyz"], src["tets"], 2, n_snap=n_snap)
    eng = H.engine_for(case, mu)
    nF = eng.nF
    steps = eng.push(case["u"], flags=PUSH_GLOBAL_FIRST, keep_wss=True)
    eng.begin(mu, case["dt"])
    d_u = eng.device_alloc(case["u"].nbytes)
    eng.h2d(d_u, case["u"])
    d_m = eng.device_alloc(9 * nF * n_snap * 8)
    eng.set_wss_layout(n_snap)
    stride = case["u"].shape[1] * 8
    eng.push_device(d_u, 33, stride, PUSH_GLOBAL_FIRST, d_m)
    eng.push_device(d_u + 33 * stride, n_snap - 33, stride, 0, d_m)
    M = np.empty((9 * nF, n_snap))
    eng.sync()
    eng.d2h(M, d_m)
    assert np.array_equal(M, steps.reshape(n_snap, -1).T)
    eng.device_free(d_u)
    eng.device_free(d_m)
    eng.close()


def test_cli_direct_matrix_equals_the_file_route(tmp_path):
    cache = {}

    def u_syn(p, t):
        if "b" not in cache:
            cache["b"] = synth.velocity_basis(p, seed=4)
        coef = np.array([[1 + 0.5 * np.sin(2 * np.pi * t), 0.2 * np.sin(4 * np.pi * t + 1), 0.1 * np.cos(6 * np.pi * t),
                          0.3 * np.sin(2 * np.pi * t + 2)]])
        return synth.velocity_series(cache["b"], coef)[0]

    _make_folder(tmp_path, u_syn, 11, 0.05, 3.5e-3)
    code = ("from pathlib import Path; from vasp_b200.compute_hemodynamics import compute_hemodyanamics as f; "
            f"p = Path({str(tmp_path)!r}); "
            "f(p / 'Visualization_separate_domain', p / 'Mesh' / 'mesh.h5', 3.5e-3, 1, block_snapshots=4, "
            "wss_matrix_folder=p / 'direct')")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True)
    assert "wss_mag.npz is saved" in out
    direct = np.load(tmp_path / "direct" / "wss_mag.npz")["component"]
    dt_files, dof_info, _ = wss_matrix.create_transformed_matrix_wss(tmp_path / "Hemodynamic_indices",
                                                                     tmp_path / "from_files", 0.0, 1e9, 1)
    via_files = np.load(tmp_path / "from_files" / "wss_mag.npz")["component"]
    assert direct.shape == via_files.shape == (dof_info["cell_dofs"].size, 10)
    assert np.array_equal(direct, via_files)
    assert abs(dt_files - 0.05) < 1e-12
