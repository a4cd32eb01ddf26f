"""End to end through the drop-in entry point on the GPU, modelled on the reference's own
tests/test_compute_hemodynamics.py: build the folder layout, run ``vasp-compute-hemo --folder``, read TAWSS back."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import hemo_oracle as ho
from tests import helpers as H
from vasp_b200 import io_dolfin, synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _make_folder(tmp: Path, u_of_points, n_snap: int, dt: float, mu, shuffle_seed=21):
    src = H.load_pipe()
    xyz, tets = src["xyz"], src["tets"]
    (tmp / "Mesh").mkdir()
    (tmp / "Checkpoint").mkdir()
    (tmp / "Visualization_separate_domain").mkdir()
    io_dolfin.write_mesh(tmp / "Mesh" / "mesh.h5", xyz, tets)
    io_dolfin.write_mesh(tmp / "Mesh" / "mesh_fluid.h5", xyz, tets)
    rx, rt = synth.refine_uniform(xyz, tets, seed=shuffle_seed)
    io_dolfin.write_mesh(tmp / "Mesh" / "mesh_refined_fluid.h5", rx, rt)
    params = dict(src["params"], mu_f=mu, dt=dt)
    (tmp / "Checkpoint" / "default_variables.json").write_text(json.dumps(params))
    times = [dt * (k + 1) for k in range(n_snap)]
    vecs = [u_of_points(rx, t) for t in times]
    io_dolfin.write_velocity_series(tmp / "Visualization_separate_domain" / "u.h5", rt, len(rx), vecs, times)
    return xyz, tets, rx, np.array(vecs), times


def test_poiseuille_through_the_cli(tmp_path):
    def u_pois(p, t):
        return np.concatenate([1.0 - p[:, 1] ** 2 - p[:, 2] ** 2, np.zeros(2 * len(p))])

    xyz, tets, rx, vecs, times = _make_folder(tmp_path, u_pois, 9, 0.1, 1)
    out = subprocess.check_output([sys.executable, "-m", "vasp_b200.compute_hemodynamics", "--folder", str(tmp_path)],
                                  cwd=ROOT, text=True)
    assert "Start post processing" in out and "Calculating WSS at Timestep: 0.1" in out
    assert "--- TAWSS is saved in" in out and "Running in serial mode" in out
    hemo = tmp_path / "Hemodynamic_indices"
    for name in ("RRT", "OSI", "ECAP", "WSS", "TAWSS", "TWSSG"):
        assert (hemo / f"{name}.xdmf").exists() and (hemo / f"{name}.h5").exists()
    r = io_dolfin.read_checkpoint(hemo, "TAWSS", 0)
    tawss, topo, geom = r["values"], r["topology"], r["geometry"]
    p = geom[topo]                                                    # (nF,3,3)
    area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
    wall = (p[:, :, 0] > 0.1).all(axis=1) & (p[:, :, 0] < 4.9).all(axis=1)
    surface_average = float((tawss.mean(axis=1) * area)[wall].sum() / area[wall].sum())
    assert 1.95 < surface_average < 2.05                              # reference test :73
    osi = io_dolfin.read_checkpoint(hemo, "OSI", 0)["values"]
    tol = 1e-12
    assert -tol <= osi.min() < 0.5 and -tol < osi.max() <= 0.5 + tol   # reference test :84-88
    # BoundaryMesh(mesh, "exterior") has order=True by default: every boundary cell lists its vertices ascending (the
    # right-oriented cells of order=False would all have outward normals; the ordered ones have both signs)
    assert np.all(np.diff(topo, axis=1) > 0)
    nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    side = (np.abs(p[:, :, 0].mean(axis=1) - 2.5) < 2.4)
    radial = p.mean(axis=1) * np.array([0, 1, 1])
    signs = np.sign(np.einsum("ij,ij->i", nrm, radial)[side])
    assert (signs > 0).any() and (signs < 0).any()


def test_pulsatile_series_matches_oracle_through_the_cli(tmp_path):
    basis_cache = {}

    def u_syn(p, t):
        if "b" not in basis_cache:
            basis_cache["b"] = synth.velocity_basis(p, seed=9)
        coef = np.array([[1 + 0.6 * np.sin(2 * np.pi * t), 0.2 * np.sin(4 * np.pi * t + 1), 0.1 * np.cos(6 * np.pi * t),
                          0.3 * np.sin(2 * np.pi * t + 2)]])
        return synth.velocity_series(basis_cache["b"], coef)[0]

    mu = [3.5e-3, 1.0]   # list => first entry, with a notice (:439-442)
    xyz, tets, rx, vecs, times = _make_folder(tmp_path, u_syn, 17, 0.05, mu)
    out = subprocess.check_output([sys.executable, "-m", "vasp_b200.compute_hemodynamics", "--folder", str(tmp_path),
                                   "--stride", "2"], cwd=ROOT, text=True)
    assert "two fluid regions are detected" in out
    sel = list(range(0, 17, 2))
    cn, edges = ho.p2_cell_nodes(tets)
    node_of_p2 = ho.match_points(ho.p2_node_coordinates(xyz, edges), rx, 1e-9)
    S = ho.SurfaceStress(xyz, tets, 3.5e-3, 2, node_of_p2)
    n = len(rx)
    res = ho.run_time_loop(S, vecs[sel], times[2] - times[0], (0, n, 2 * n), keep_wss=True)
    fin = ho.finalize(res["wss_sum"], res["tawss_sum"], res["twssg_sum"], res["count"])
    hemo = tmp_path / "Hemodynamic_indices"
    for name in H.FIELDS:
        got = io_dolfin.read_checkpoint(hemo, name, 0)["values"]
        assert H.rel_l2(got, fin[name]) < 1e-10, name
    for k in range(len(sel)):
        got = io_dolfin.read_checkpoint(hemo, "WSS", k)["values"].reshape(-1, 3, 3).transpose(0, 2, 1)
        assert H.rel_l2(got, res["wss"][k]) < 1e-10
    topo = io_dolfin.read_checkpoint(hemo, "WSS", 0)["topology"]
    assert np.array_equal(topo, S.maps.btopology)


def test_raw_turtlefsi_output_gives_the_same_fields_as_u_h5(tmp_path):
    """No ``Visualization_separate_domain``: the reference would run create_hdf5() first (:389-431); here the raw
    ``VisualisationVector`` arrays (whole domain, interleaved, restarted run in two files) are sliced on the GPU.
    The result must be bit-identical to the run over the u.h5 that create_hdf5 would have written."""
    basis_cache = {}

    def u_syn(p, t):
        if "b" not in basis_cache:
            basis_cache["b"] = synth.velocity_basis(p, seed=4)
        coef = np.array([[1 + 0.5 * np.sin(2 * np.pi * t), 0.3 * np.sin(4 * np.pi * t + 1), 0.2 * np.cos(6 * np.pi * t),
                          0.1 * np.sin(2 * np.pi * t + 2)]])
        return synth.velocity_series(basis_cache["b"], coef)[0]

    raw = tmp_path / "raw"
    raw.mkdir()
    info = H.write_turtle_folder(raw, u_syn, n_snap=11, dt=0.01, mu=3.5e-3, save_step=5, split_at=6)
    out = subprocess.check_output([sys.executable, "-m", "vasp_b200.compute_hemodynamics", "--folder", str(raw)],
                                  cwd=ROOT, text=True)
    assert "Visualization_separate_domain folder not found" in out and "save_time_step: 0.05" in out
    assert not (raw / "Visualization_separate_domain").exists()
    # the converted layout of the same series
    conv = tmp_path / "conv"
    conv.mkdir()
    for sub in ("Mesh", "Checkpoint"):
        (conv / sub).mkdir()
        for f in (raw / sub).iterdir():
            (conv / sub / f.name).write_bytes(f.read_bytes())
    (conv / "Visualization_separate_domain").mkdir()
    io_dolfin.write_velocity_series(conv / "Visualization_separate_domain" / "u.h5", info["rt"], len(info["rx"]),
                                    info["vecs"], info["times"])
    subprocess.check_output([sys.executable, "-m", "vasp_b200.compute_hemodynamics", "--folder", str(conv)],
                            cwd=ROOT, text=True)
    for name in H.FIELDS:
        a = io_dolfin.read_checkpoint(raw / "Hemodynamic_indices", name, 0)["values"]
        b = io_dolfin.read_checkpoint(conv / "Hemodynamic_indices", name, 0)["values"]
        assert np.array_equal(a, b), name
    for k in (0, 10):
        a = io_dolfin.read_checkpoint(raw / "Hemodynamic_indices", "WSS", k)["values"]
        b = io_dolfin.read_checkpoint(conv / "Hemodynamic_indices", "WSS", k)["values"]
        assert np.array_equal(a, b)
    # and against the oracle
    cn, edges = ho.p2_cell_nodes(info["tets"])
    node_of_p2 = ho.match_points(ho.p2_node_coordinates(info["xyz"], edges), info["rx"], 1e-9)
    S = ho.SurfaceStress(info["xyz"], info["tets"], 3.5e-3, 2, node_of_p2)
    n = len(info["rx"])
    res = ho.run_time_loop(S, info["vecs"], 0.05, (0, n, 2 * n))
    fin = ho.finalize(res["wss_sum"], res["tawss_sum"], res["twssg_sum"], res["count"])
    for name in H.FIELDS:
        got = io_dolfin.read_checkpoint(raw / "Hemodynamic_indices", name, 0)["values"]
        assert H.rel_l2(got, fin[name]) < 1e-10, name


def test_derived_refined_numbering_on_the_gpu(tmp_path):
    """SURVEY.md §8f-4 on the real engine: ``--derive-refined-mesh`` matches the P2 nodes of the wall cells against the
    geometry stored in the raw velocity file (all nodes of the whole domain, solid ones included) instead of
    ``mesh_refined_fluid.h5``; ``mesh_refined.h5`` / ``mesh_refined_fluid.h5`` are deleted.  Fields and WSS series must be
    bit-identical to the route that uses them, and within 1e-10 of the oracle."""
    basis_cache = {}

    def u_syn(p, t):
        if "b" not in basis_cache:
            basis_cache["b"] = synth.velocity_basis(p, seed=4)
        coef = np.array([[1 + 0.5 * np.sin(2 * np.pi * t), 0.3 * np.sin(4 * np.pi * t + 1), 0.2 * np.cos(6 * np.pi * t),
                          0.1 * np.sin(2 * np.pi * t + 2)]])
        return synth.velocity_series(basis_cache["b"], coef)[0]

    a, b = tmp_path / "with", tmp_path / "without"
    for d in (a, b):
        d.mkdir()
        info = H.write_turtle_folder(d, u_syn, n_snap=9, dt=0.01, mu=3.5e-3, save_step=5, split_at=6)
    (b / "Mesh" / "mesh_refined.h5").unlink()
    (b / "Mesh" / "mesh_refined_fluid.h5").unlink()
    run = [sys.executable, "-m", "vasp_b200.compute_hemodynamics", "--folder"]
    subprocess.check_output(run + [str(a)], cwd=ROOT, text=True)
    assert subprocess.run(run + [str(b)], cwd=ROOT, capture_output=True, text=True).returncode != 0   # reference behaviour
    subprocess.check_output(run + [str(b), "--derive-refined-mesh"], cwd=ROOT, text=True)
    for name in H.FIELDS:
        x = io_dolfin.read_checkpoint(a / "Hemodynamic_indices", name, 0)["values"]
        y = io_dolfin.read_checkpoint(b / "Hemodynamic_indices", name, 0)["values"]
        assert np.array_equal(x, y), name
    for k in range(9):
        x = io_dolfin.read_checkpoint(a / "Hemodynamic_indices", "WSS", k)["values"]
        y = io_dolfin.read_checkpoint(b / "Hemodynamic_indices", "WSS", k)["values"]
        assert np.array_equal(x, y), k
    case = {"xyz": info["xyz"], "tets": info["tets"], "order": 2, "points": info["rx"], "u": info["vecs"],
            "dt": 0.05, "n_nodes": len(info["rx"])}
    _, res, fin = H.oracle_run(case, 3.5e-3)
    for name in H.FIELDS:
        y = io_dolfin.read_checkpoint(b / "Hemodynamic_indices", name, 0)["values"]
        assert H.rel_l2(y, fin[name]) < 1e-10, name
