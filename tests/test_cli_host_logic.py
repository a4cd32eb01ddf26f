"""The entry point's host logic in the CPU suite: ``main()`` / ``compute_hemodyanamics()`` run in-process with the
oracle-backed stand-in of ``tests/fake_engine.py`` in place of the CUDA engine (the real engine runs the same
scenarios in ``tests/test_gpu_cli.py``).  What is exercised here is everything *around* the kernels: flags, stride,
``dt``, block reader, order and content of the ``WSS`` series, the index files, raw turtleFSI input, the WSS matrix."""
import numpy as np
import pytest

from oracle import hemo_oracle as ho
from tests import helpers as H
from tests.fake_engine import OracleHemoEngine
from tests.test_gpu_cli import _make_folder
from vasp_b200 import compute_hemodynamics as ch
from vasp_b200 import engine as engine_mod
from vasp_b200 import io_dolfin, synth, wss_matrix


@pytest.fixture()
def cpu_engine(monkeypatch):
    monkeypatch.setattr(ch, "HemoEngine", OracleHemoEngine)
    monkeypatch.setattr(ch, "pinned_empty", lambda shape: np.zeros(shape))
    monkeypatch.setattr(ch, "device_count", lambda: 1)
    monkeypatch.setattr(engine_mod, "pinned_empty", lambda shape: np.zeros(shape))
    for n in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "PMI_RANK", "PMI_SIZE",
              "SLURM_PROCID", "SLURM_NTASKS"):
        monkeypatch.delenv(n, raising=False)


def _u_syn(seed):
    cache = {}

    def f(p, t):
        if "b" not in cache:
            cache["b"] = synth.velocity_basis(p, seed=seed)
        coef = np.array([[1 + 0.6 * np.sin(2 * np.pi * t), 0.2 * np.sin(4 * np.pi * t + 1), 0.1 * np.cos(6 * np.pi * t),
                          0.3 * np.sin(2 * np.pi * t + 2)]])
        return synth.velocity_series(cache["b"], coef)[0]
    return f


def test_stride_dt_series_order_and_index_files(cpu_engine, tmp_path, capsys):
    xyz, tets, rx, vecs, times = _make_folder(tmp_path, _u_syn(9), 9, 0.05, [3.5e-3, 1.0])
    ch.main(["--folder", str(tmp_path), "--stride", "2"])
    out = capsys.readouterr().out
    assert "two fluid regions are detected" in out and "Running in serial mode" in out
    sel = list(range(0, 9, 2))
    for k in sel:
        assert f"Calculating WSS at Timestep: {times[k]} =" in out
    assert f"Calculating WSS at Timestep: {times[1]} =" not in out
    node_of_p2 = ho.match_points(ho.p2_node_coordinates(xyz, ho.p2_cell_nodes(tets)[1]), rx, 1e-9)
    S = ho.SurfaceStress(xyz, tets, 3.5e-3, 2, node_of_p2)
    n = len(rx)
    res = ho.run_time_loop(S, vecs[sel], times[2] - times[0], (0, n, 2 * n), keep_wss=True)   # dt = gap of the selection
    fin = ho.finalize(res["wss_sum"], res["tawss_sum"], res["twssg_sum"], res["count"])
    hemo = tmp_path / "Hemodynamic_indices"
    assert sorted(p.name for p in hemo.iterdir()) == sorted(f"{n}.{e}" for n in ch.INDEX_NAMES for e in ("h5", "xdmf"))
    for name in H.FIELDS:
        assert H.rel_l2(io_dolfin.read_checkpoint(hemo, name, 0)["values"], fin[name]) < 1e-12, name
    for k in range(len(sel)):
        got = io_dolfin.read_checkpoint(hemo, "WSS", k)["values"].reshape(-1, 3, 3).transpose(0, 2, 1)
        assert np.array_equal(got, res["wss"][k])
    _, ts, idx = wss_matrix.output_file_lists(hemo / "WSS.xdmf")
    assert ts == [times[k] for k in sel] and idx == list(range(len(sel)))
    assert np.array_equal(io_dolfin.read_checkpoint(hemo, "WSS", 0)["topology"], S.maps.btopology)


def test_raw_turtlefsi_route_equals_the_u_h5_route(cpu_engine, tmp_path, capsys):
    raw = tmp_path / "raw"
    raw.mkdir()
    info = H.write_turtle_folder(raw, _u_syn(4), n_snap=7, dt=0.01, mu=3.5e-3, save_step=5, split_at=4)
    ch.main(["--folder", str(raw)])
    out = capsys.readouterr().out
    assert "Visualization_separate_domain folder not found" in out and "save_time_step: 0.05" in out
    assert not (raw / "Visualization_separate_domain").exists()
    conv = tmp_path / "conv"
    for sub in ("Mesh", "Checkpoint", "Visualization_separate_domain"):
        (conv / sub).mkdir(parents=True)
    for sub in ("Mesh", "Checkpoint"):
        for f in (raw / sub).iterdir():
            (conv / sub / f.name).write_bytes(f.read_bytes())
    io_dolfin.write_velocity_series(conv / "Visualization_separate_domain" / "u.h5", info["rt"], len(info["rx"]),
                                    info["vecs"], info["times"])
    ch.main(["--folder", str(conv)])
    for name in H.FIELDS:
        a = io_dolfin.read_checkpoint(raw / "Hemodynamic_indices", name, 0)["values"]
        b = io_dolfin.read_checkpoint(conv / "Hemodynamic_indices", name, 0)["values"]
        assert np.array_equal(a, b), name
    for k in range(7):
        a = io_dolfin.read_checkpoint(raw / "Hemodynamic_indices", "WSS", k)["values"]
        b = io_dolfin.read_checkpoint(conv / "Hemodynamic_indices", "WSS", k)["values"]
        assert np.array_equal(a, b), k
    assert (raw / "Hemodynamic_indices" / "WSS.xdmf").read_text() == (conv / "Hemodynamic_indices" / "WSS.xdmf").read_text()


def test_direct_wss_matrix_equals_the_file_route(cpu_engine, tmp_path, capsys):
    _make_folder(tmp_path, _u_syn(2), 7, 0.05, 3.5e-3)
    ch.compute_hemodyanamics(tmp_path / "Visualization_separate_domain", tmp_path / "Mesh" / "mesh.h5", 3.5e-3, 1,
                             block_snapshots=3, wss_matrix_folder=tmp_path / "direct")
    assert "wss_mag.npz is saved" in capsys.readouterr().out
    direct = np.load(tmp_path / "direct" / "wss_mag.npz")["component"]
    dt_files, dof_info, _ = wss_matrix.create_transformed_matrix_wss(tmp_path / "Hemodynamic_indices",
                                                                     tmp_path / "files", 0.0, 1e9, 1)
    via_files = np.load(tmp_path / "files" / "wss_mag.npz")["component"]
    assert direct.shape == via_files.shape == (dof_info["cell_dofs"].size, 6) and np.array_equal(direct, via_files)
    assert abs(dt_files - 0.05) < 1e-12


def test_p1_extension_reads_vertex_data(cpu_engine, tmp_path):
    """--velocity-degree 1: velocity on the vertices of the un-refined mesh (the BASELINE P1 configs; the reference
    refuses save_deg != 2 at compute_hemodynamics.py:435-436)."""
    src = H.load_fluid("cylinder")
    xyz, tets = src["xyz"], src["tets"]
    for sub in ("Mesh", "Checkpoint", "Visualization_separate_domain"):
        (tmp_path / sub).mkdir()
    io_dolfin.write_mesh(tmp_path / "Mesh" / "mesh.h5", xyz, tets)
    io_dolfin.write_mesh(tmp_path / "Mesh" / "mesh_fluid.h5", xyz, tets)
    (tmp_path / "Checkpoint" / "default_variables.json").write_text('{"mu_f": 0.0035, "dt": 0.1, "save_step": 1, "save_deg": 1}')
    case = H.make_case(xyz, tets, 1, n_snap=6)
    times = [0.1 * (k + 1) for k in range(6)]
    io_dolfin.write_velocity_series(tmp_path / "Visualization_separate_domain" / "u.h5", tets, len(xyz), case["u"], times)
    with pytest.raises(AssertionError, match="save_deg = 2"):
        ch.main(["--folder", str(tmp_path)])
    ch.main(["--folder", str(tmp_path), "--velocity-degree", "1"])
    S, res, fin = H.oracle_run(dict(case, dt=0.1), 0.0035)
    for name in H.FIELDS:
        got = io_dolfin.read_checkpoint(tmp_path / "Hemodynamic_indices", name, 0)["values"]
        assert H.rel_l2(got, fin[name]) < 1e-12, name


@pytest.mark.parametrize("route", ["u_h5", "raw"])
def test_compact_route_reads_only_the_wall_layer_out_of_the_file_mapping(cpu_engine, tmp_path, monkeypatch, route):
    """With wall-layer compaction active the block reader never copies a snapshot whole: the engine gathers the
    wall-layer dofs straight out of the read-only mapping of ``u.h5`` (or of the raw turtleFSI arrays) and the entry
    point pushes compact rows.  Output files must equal those of the whole-vector route byte for byte; the stand-in
    poisons everything outside the wall layer, so a dof that should not matter cannot leak in."""
    folders = {}
    for mode in (False, True):
        monkeypatch.setattr(OracleHemoEngine, "compact_route", mode)
        d = tmp_path / f"run{int(mode)}"
        d.mkdir()
        if route == "raw":
            H.write_turtle_folder(d, _u_syn(6), n_snap=7, dt=0.01, mu=3.5e-3, save_step=2, split_at=3)
        else:
            _make_folder(d, _u_syn(6), 7, 0.05, 3.5e-3)
        ch.main(["--folder", str(d)])
        folders[mode] = d / "Hemodynamic_indices"
    for name in ch.INDEX_NAMES:
        assert (folders[True] / f"{name}.h5").read_bytes() == (folders[False] / f"{name}.h5").read_bytes(), name


def test_block_reader_in_compact_mode_handles_ragged_blocks(cpu_engine, tmp_path, monkeypatch):
    monkeypatch.setattr(OracleHemoEngine, "compact_route", True)
    xyz, tets, rx, vecs, times = _make_folder(tmp_path, _u_syn(8), 8, 0.05, 3.5e-3)
    series = io_dolfin.VelocitySeries(tmp_path / "Visualization_separate_domain" / "u.h5", "velocity", 1)
    eng = OracleHemoEngine()
    eng.set_mesh(xyz, tets)
    _, rt = io_dolfin.read_mesh(tmp_path / "Mesh" / "mesh_refined_fluid.h5")
    comp_offset, node_stride, perm = series.layout(rt, len(rx))
    eng.set_velocity_layout(2, refined_xyz=rx, node_perm=perm, comp_offset=comp_offset, node_stride=node_stride)
    reader = ch._BlockReader(series, 1, 8, 3, eng)          # snapshots 1..7 in blocks of 3, 4 (no single trailing one)
    assert reader.compact and reader.ranges == [(1, 4), (4, 8)] and reader.max_rows == 4
    sl, nwp, n = eng._slots().astype(np.int64), eng.compact_len // 3, len(rx)
    for a, b, c in reader:
        for k in range(3):
            assert np.array_equal(c[:, k * nwp:k * nwp + len(sl)], vecs[a:b, k * n + sl])
    series.close()


def test_derived_refined_numbering_needs_no_refined_mesh_files(cpu_engine, tmp_path, capsys):
    """with --derive-refined-mesh the raw route matches the P2 nodes of mesh_fluid.h5 against the
    geometry stored in Visualization/velocity.h5; mesh_refined.h5 and mesh_refined_fluid.h5 (the products of
    create_refined_mesh.py:50-151 and separate_mesh.py:56-107) can be absent.  Same output files as the route that
    uses them."""
    a, b = tmp_path / "with", tmp_path / "without"
    for d in (a, b):
        d.mkdir()
        H.write_turtle_folder(d, _u_syn(4), n_snap=6, dt=0.01, mu=3.5e-3, save_step=5, split_at=4)
    (b / "Mesh" / "mesh_refined.h5").unlink()
    (b / "Mesh" / "mesh_refined_fluid.h5").unlink()
    ch.main(["--folder", str(a)])
    with pytest.raises(AssertionError, match="mesh_refined.h5 not found"):
        ch.main(["--folder", str(b)])                        # the reference's behaviour stays the default
    ch.main(["--folder", str(b), "--derive-refined-mesh"])
    capsys.readouterr()
    for name in ch.INDEX_NAMES:
        assert (a / "Hemodynamic_indices" / f"{name}.h5").read_bytes() == \
            (b / "Hemodynamic_indices" / f"{name}.h5").read_bytes(), name
