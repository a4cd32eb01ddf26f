"""The host-side restatements against outputs of THE REFERENCE'S OWN CODE.

``tests/golden/reference_goldens.npz`` was produced by ``tests/golden/make_reference_goldens.py``, which imports the
unmodified reference modules from ``/root/reference/src`` (with shims for the absent ``h5py`` / ``matplotlib`` /
``dolfin`` modules) and runs ``output_file_lists``, ``get_domain_ids``, ``read_parameters_from_file``,
``parse_arguments`` and ``create_transformed_matrix(quantity="wss")`` on the inputs stored in the same file.  Here the
same inputs go through this repository's versions; results must be identical (bitwise for arrays)."""
import json
from pathlib import Path

import numpy as np
import pytest

from vasp_b200 import compute_hemodynamics as ch
from vasp_b200 import io_dolfin, io_turtle, wss_matrix
from vasp_b200.h5lite import H5Writer

G = np.load(Path(__file__).resolve().parent / "golden" / "reference_goldens.npz")


def _j(key):
    return json.loads(str(G[key]))


def test_output_file_lists_on_a_restarted_turtlefsi_series(tmp_path):
    (tmp_path / "velocity.xdmf").write_text(str(G["ofl_turtle_xdmf"]))
    h5s, ts, idx = io_turtle.output_file_lists(tmp_path / "velocity.xdmf")
    want = _j("ofl_turtle")
    assert [list(h5s), list(ts), list(idx)] == want
    assert want[0].count("velocity_run_1.h5") == 3 and want[2][6:] == [0, 1, 2]      # the restart really is in there


def _write_wss_case(folder):
    vals, times = G["ctm_vals"], G["ctm_times"].tolist()
    for name, vector in (("WSS", True), ("MaxPrincipalStrain", False)):
        w = io_dolfin.CheckpointWriter(folder, name, G["ctm_btopo"], G["ctm_bgeom"], vector)
        for v, t in zip(vals, times):
            w.write(v if vector else np.linalg.norm(v, axis=2), t)
        w.close()


def test_output_file_lists_on_write_checkpoint_files(tmp_path):
    _write_wss_case(tmp_path)
    h5s, ts, idx = io_turtle.output_file_lists(tmp_path / "WSS.xdmf")
    assert [list(h5s), list(ts), list(idx)] == _j("ofl_checkpoint")


@pytest.mark.parametrize("k", range(5))
def test_wss_matrix_equals_the_reference_create_transformed_matrix(tmp_path, k):
    """The reference function itself read WSS.h5 / WSS.xdmf written by this repository's writer (through the h5py
    shim) and produced these matrices; the restatement and the column selection must give the same file."""
    st, et, stride = G["ctm_cases"][k]
    case = tmp_path / "case" / "Hemodynamic_indices"
    case.mkdir(parents=True)
    _write_wss_case(case)
    dt_files, dof_info, dof_amp = wss_matrix.create_transformed_matrix_wss(case, tmp_path / "npz", float(st), float(et),
                                                                          int(stride))
    got = np.load(tmp_path / "npz" / "wss_mag.npz")["component"]
    want = G[f"ctm{k}_matrix"]
    assert got.shape == want.shape == (54, 7) and np.array_equal(got, want)
    assert dt_files == float(G[f"ctm{k}_dt"])
    for name in wss_matrix.DOF_INFO_NAMES:
        key = name.replace("/", "__")
        assert np.array_equal(dof_info[name], G["ctm_dofinfo_" + key]), name
        assert np.array_equal(dof_amp[name], G["ctm_dofamp_" + key]), name
    # the direct route's column picker agrees with what the reference kept
    cols = wss_matrix.select_columns(G["ctm_times"].tolist(), float(st), float(et), int(stride))
    flat = G["ctm_vals"].reshape(len(G["ctm_times"]), -1)
    ref_cols = np.zeros_like(want)
    if cols:
        ref_cols[:, :len(cols)] = flat[cols].T
    assert np.array_equal(ref_cols, want)


@pytest.mark.parametrize("name", ["cylinder", "stenosis", "aneurysm"])
def test_get_domain_ids_on_the_reference_test_meshes(tmp_path, name):
    fid, sid = _j(f"ids_{name}_query")
    p = tmp_path / "mesh.h5"
    with H5Writer(p) as w:
        w.create_dataset("/domains/values", G[f"ids_{name}_domains"].astype("<u8"))
        w.create_dataset("/domains/topology", G[f"ids_{name}_topology"].astype("<i8"))
    f, s, a = io_turtle.get_domain_ids(p, fid, sid)
    for got, key in ((f, "fluid"), (s, "solid"), (a, "all")):
        want = G[f"ids_{name}_{key}"]
        assert np.array_equal(np.asarray(got), want) and len(want) > 0, key


def test_read_parameters_from_file_cases(tmp_path):
    (tmp_path / "p1" / "Checkpoint").mkdir(parents=True)
    (tmp_path / "p1" / "Checkpoint" / "default_variables.json").write_text(str(G["par_input"]))
    (tmp_path / "p2" / "Checkpoint").mkdir(parents=True)
    (tmp_path / "p2" / "Checkpoint" / "default_variables.json").write_text("{ not json")
    got = [ch.read_parameters_from_file(tmp_path / d) for d in ("p1", "p2", "p3")]
    assert got == _j("par") and got[1] is None and got[2] is None


def test_parse_arguments_matches_the_reference_parser():
    a = _j("args")
    for argv, want in zip(a["argv"], a["namespaces"]):
        ns = vars(ch.parse_arguments(argv))
        got = {k: (str(v) if isinstance(v, Path) else v) for k, v in ns.items() if k in want}
        assert got == want, argv
        assert set(ns) - set(want) == {"velocity_degree", "device"}      # the two documented extensions
        assert ns["velocity_degree"] == 2 and ns["device"] is None       # ... default to the reference behaviour
