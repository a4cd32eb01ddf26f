"""Host-side logic that needs no GPU: dataset selection, dof-layout detection, outputs, CLI contract, shard plans,
and that the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest

from tests import helpers as H
from vasp_b200 import _lib, io_dolfin, synth, timeshard
from vasp_b200.h5lite import H5File, H5Writer

ROOT = Path(__file__).resolve().parents[1]


def test_library_loads_and_exports_every_declared_symbol():
    header = (ROOT / "include" / "vasp_hemo.h").read_text()
    declared = set(re.findall(r"\b(vh_[a-z0-9_]+)\s*\(", header)) - {"vh_handle", "vh_status"}
    assert len(declared) >= 35
    assert _lib.LIB_PATH.exists(), "libvasp_hemo.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))   # loading needs no GPU; compute calls would fail loudly
    missing = sorted(s for s in declared if not hasattr(lib, s))
    assert not missing, missing
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)


def test_no_gpu_means_loud_failure_not_a_fallback():
    lib = _lib.load()
    n = ctypes.c_int(0)
    rc = lib.vh_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible here")
    h = ctypes.c_void_p()
    assert lib.vh_create(0, ctypes.byref(h)) != 0
    from vasp_b200.engine import HemoEngine
    with pytest.raises(_lib.VaspHemoError):
        HemoEngine(0)


def test_host_gather_of_the_wall_layer_is_a_pure_copy():
    """csrc/compact.cu: the host-side compaction in front of the bus (handle-free entry point, no GPU needed) picks
    vec[comp_offset[c] + slot[i]] bit for bit -- blocked and interleaved layouts, rows by stride and by address."""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    n_nodes, n_w, n = 70001, 20011, 7
    slots = np.sort(rng.choice(n_nodes, n_w, replace=False)).astype(np.int32)
    for off, node_stride in (((0, n_nodes, 2 * n_nodes), 1), ((0, 1, 2), 3)):
        u = rng.normal(size=(n, 3 * n_nodes + 5))
        sl = (slots * node_stride).astype(np.int32)
        want = np.stack([np.concatenate([u[r, o + sl.astype(np.int64)] for o in off]) for r in range(n)])
        offs = (ctypes.c_int64 * 3)(*off)
        for threads in (1, 3):
            got = np.full((n, 3 * n_w + 2), -7.0)
            rc = lib.vh_host_gather(None, u.ctypes.data, u.strides[0], n, sl.ctypes.data, n_w, offs, got.ctypes.data,
                                    got.strides[0], threads)
            assert rc == 0 and np.array_equal(got[:, :3 * n_w], want) and np.all(got[:, 3 * n_w:] == -7.0)
        rows = np.array([u[r].ctypes.data for r in (4, 0, 6)], dtype=np.uint64)
        got = np.empty((3, 3 * n_w))
        assert lib.vh_host_gather(rows.ctypes.data, None, 0, 3, sl.ctypes.data, n_w, offs, got.ctypes.data,
                                  got.strides[0], 2) == 0
        assert np.array_equal(got, want[[4, 0, 6]])
    assert lib.vh_host_gather(None, None, 0, 1, slots.ctypes.data, n_w, offs, got.ctypes.data, got.strides[0], 1) != 0


def test_product_never_imports_the_oracle():
    for p in (ROOT / "vasp_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), p
        assert "torch" not in re.findall(r"^\s*(?:from|import)\s+(\w+)", src, re.M) or p.name == "timeshard.py", p


def test_get_dataset_names_selection_rule(tmp_path):
    p = tmp_path / "u.h5"
    with H5Writer(p) as w:
        for k in (2, 3, 4, 5, 6, 8, 9):
            w.create_dataset(f"/velocity/vector_{k}", np.zeros(3), attrs={"timestamp": float(k)})
        w.create_dataset("/velocity/cells", np.zeros(1))
    with H5File(p) as f:
        g = f["velocity"]
        assert io_dolfin.get_dataset_names(g, step=1) == [f"vector_{k}" for k in (2, 3, 4, 5, 6, 8, 9)]
        assert io_dolfin.get_dataset_names(g, step=2) == [f"vector_{k}" for k in (2, 4, 6, 8)]
        assert io_dolfin.get_dataset_names(g, step=3) == [f"vector_{k}" for k in (3, 6, 9)]


@pytest.mark.parametrize("kind", ["blocked", "interleaved", "blocked_perm", "no_tables"])
def test_velocity_layout_detection(tmp_path, kind):
    src = H.load_fluid("cylinder")
    rx, rt = synth.refine_uniform(src["xyz"], src["tets"], seed=3)
    n, nc = len(rx), len(rt)
    rng = np.random.default_rng(0)
    q = rng.permutation(n)
    vals = rng.random((3, n))                       # value of (comp, refined vertex)
    if kind in ("blocked", "no_tables"):
        idx = np.arange(3)[:, None] * n + np.arange(n)[None, :]
    elif kind == "interleaved":
        idx = 3 * q[None, :] + np.arange(3)[:, None]
    else:
        idx = np.arange(3)[:, None] * n + q[None, :]
    vec = np.empty(3 * n)
    vec[idx] = vals
    p = tmp_path / "u.h5"
    with H5Writer(p) as w:
        for k in range(2):
            w.create_dataset(f"/velocity/vector_{k}", vec, attrs={"timestamp": 0.1 * k})
        if kind != "no_tables":
            cell_dofs = idx[:, rt].transpose(1, 0, 2).reshape(-1)   # per cell: comp-major, 4 vertices
            w.create_dataset("/velocity/cell_dofs", cell_dofs.astype("<i8"))
            w.create_dataset("/velocity/x_cell_dofs", (12 * np.arange(nc + 1)).astype("<i8"))
            w.create_dataset("/velocity/cells", np.arange(nc, dtype="<i8"))
    s = io_dolfin.VelocitySeries(p)
    off, stride, perm = s.layout(rt, n)
    pv = np.arange(n) if perm is None else perm
    for c in range(3):
        assert np.array_equal(vec[off[c] + stride * pv], vals[c])
    assert s.timestamps.tolist() == [0.0, 0.1] and s.vec_len == 3 * n
    buf = np.zeros((2, 3 * n + 5))
    s.read_into(buf, 0, 2)
    assert np.array_equal(buf[1, :3 * n], vec)
    s.close()


def test_series_reads_in_pieces_over_a_thread_pool(tmp_path, monkeypatch):
    """The CLI's block reader cuts long vectors and groups short ones into ~4 MiB jobs for several pread threads
    (vasp_b200/io_dolfin.py read_chunks); whatever the cut, the rows must come back byte for byte."""
    from concurrent.futures import ThreadPoolExecutor
    from vasp_b200 import compute_hemodynamics as ch
    from vasp_b200.compute_hemodynamics import _BlockReader, default_block_snapshots
    monkeypatch.setattr(ch, "pinned_empty", lambda shape: np.zeros(shape))   # no CUDA driver in the CPU suite
    rng = np.random.default_rng(8)
    n, n_snap = 211, 23
    vecs = rng.normal(size=(n_snap, 3 * n))
    p = tmp_path / "u.h5"
    with H5Writer(p) as w:
        for k in range(n_snap):
            w.create_dataset(f"/velocity/vector_{k}", vecs[k], attrs={"timestamp": 0.1 * k})
    s = io_dolfin.VelocitySeries(p)
    for chunk in (64, 1000, 3 * n * 8, 10 ** 6):
        jobs = s.read_chunks(np.zeros((5, 3 * n)), 3, 8, chunk_bytes=chunk)
        assert sum(len(mv) for job in jobs for mv, _ in job) == 5 * 3 * n * 8
        assert all(sum(len(mv) for mv, _ in job) >= min(chunk, 8) for job in jobs[:-1])
    buf = np.zeros((n_snap, 3 * n + 3))
    with ThreadPoolExecutor(3) as pool:
        s.read_into(buf, 0, n_snap, pool)
    assert np.array_equal(buf[:, :3 * n], vecs) and (buf[:, 3 * n:] == 0).all()
    got = []
    for a, b, u in _BlockReader(s, 2, n_snap, 4):   # two buffers, one block ahead; no single trailing snapshot
        assert b - a >= 2
        got.append(np.array(u[:, :3 * n]))
    assert np.array_equal(np.concatenate(got), vecs[2:])
    assert default_block_snapshots(3 * 2997) == 116 and default_block_snapshots(3 * 13_400_000) == 2
    # ~8 MiB of the larger of what is read (compact rows when the wall layer is gathered on the way) and of the WSS block
    assert default_block_snapshots(3 * 2997, 0, 2560) == 45
    assert default_block_snapshots(3 * 13_400_000, 3 * 998_688, 224_768) == 2
    assert default_block_snapshots(3 * 351_329, 3 * 70_688) == 4 and default_block_snapshots(3 * 351_329, 3 * 70_688, 72_960) == 2
    s.close()


def test_checkpoint_writer_layout(tmp_path):
    """Member names and shapes the reference's own consumer dereferences
    (postprocessing_h5py_common.py:234-242,271,337-343) and regexes it parses XDMF with
    (postprocessing_common.py:92-94)."""
    nF, nBV = 5, 7
    rng = np.random.default_rng(1)
    topo, geom = rng.integers(0, nBV, (nF, 3)), rng.random((nBV, 3))
    w = io_dolfin.CheckpointWriter(tmp_path, "WSS", topo, geom, True)
    vals = [rng.random((nF, 3, 3)) for _ in range(3)]
    for k, v in enumerate(vals):
        w.write(v, 0.5 + 0.1 * k)
    w.close()
    with H5File(tmp_path / "WSS.h5") as f:
        assert list(f.keys()) == ["WSS"]
        assert sorted(f["WSS"].keys()) == ["WSS_0", "WSS_1", "WSS_2"]
        for member in ("cell_dofs", "cells", "mesh/geometry", "mesh/topology", "x_cell_dofs"):
            assert member in f["WSS/WSS_0"]
        assert f["WSS/WSS_2/vector"].shape == (9 * nF, 1)
        assert np.array_equal(f["WSS/WSS_1/vector"].read().reshape(nF, 3, 3), vals[1])   # node-interleaved xyz
        assert f["WSS/WSS_0/mesh/topology"].shape == (nF, 3) and f["WSS/WSS_0/mesh/geometry"].shape == (nBV, 3)
    for k in range(3):
        r = io_dolfin.read_checkpoint(tmp_path, "WSS", k)["values"]            # per cell: comp-major (UFC)
        assert np.array_equal(r.reshape(nF, 3, 3).transpose(0, 2, 1), vals[k])
    x = (tmp_path / "WSS.xdmf").read_text()
    assert re.findall('<Time Value="(.+?)"', x) == ["0.5", "0.6", "0.7"]
    assert re.findall(r'"HDF">(.*?):', x)[0] == "WSS.h5"
    assert [int(i) for i in re.findall(r'_([0-9]+)\/vector', x)] == [0, 1, 2]
    assert 'ElementFamily="DG" ElementDegree="1" ElementCell="triangle"' in x and 'AttributeType="Vector"' in x
    s = io_dolfin.CheckpointWriter(tmp_path, "TAWSS", topo, geom, False)
    s.write(rng.random((nF, 3)), 0)
    s.close()
    assert 'AttributeType="Scalar"' in (tmp_path / "TAWSS.xdmf").read_text()
    with H5File(tmp_path / "TAWSS.h5") as f:
        assert f["TAWSS/TAWSS_0/vector"].shape == (3 * nF, 1)


def test_shard_plan_covers_the_series_once():
    for n, world in ((10, 1), (10, 3), (7, 8), (2000, 8), (5, 5)):
        shards = [timeshard.plan_shard(n, r, world) for r in range(world)]
        assert shards[0].start == 0 and shards[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
        assert max(s.count for s in shards) - min(s.count for s in shards) <= 1
        assert not shards[0].has_halo and shards[0].first_push_flags() == 1
        for s in shards[1:]:
            assert s.has_halo == (s.count > 0)
            if s.has_halo:
                assert s.read_start == s.start - 1 and s.first_push_flags() == 2


def test_cli_flag_set_and_error_conventions(tmp_path, monkeypatch):
    from vasp_b200 import compute_hemodynamics as ch
    a = ch.parse_arguments(["--folder", "x", "--mesh-path", "m.h5", "--stride", "3", "-st", "0.1", "-et", "0.9",
                            "--extract-entire-domain", "--log-level", "10"])
    assert (a.folder, a.mesh_path, a.stride, a.start_time, a.end_time, a.extract_entire_domain, a.log_level) == \
        (Path("x"), Path("m.h5"), 3, 0.1, 0.9, True, 10)
    d = ch.parse_arguments(["--folder", "x"])
    assert (d.mesh_path, d.stride, d.start_time, d.end_time, d.extract_entire_domain, d.log_level) == \
        (None, 1, None, None, False, 20)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    with pytest.raises(AssertionError, match="not found"):
        ch.main(["--folder", str(tmp_path / "missing")])
    with pytest.raises(RuntimeError, match="Error reading parameters from file."):
        ch.main(["--folder", str(tmp_path)])                       # no Checkpoint/default_variables.json
    (tmp_path / "Checkpoint").mkdir()
    (tmp_path / "Checkpoint" / "default_variables.json").write_text("{not json")
    with pytest.raises(RuntimeError):
        ch.main(["--folder", str(tmp_path)])
    (tmp_path / "Checkpoint" / "default_variables.json").write_text(json.dumps({"save_deg": 1, "mu_f": [1.0, 2.0]}))
    (tmp_path / "Visualization_separate_domain").mkdir()
    with pytest.raises(AssertionError, match="save_deg = 2"):
        ch.main(["--folder", str(tmp_path)])
    (tmp_path / "Checkpoint" / "default_variables.json").write_text(json.dumps({"save_deg": 2, "mu_f": 1.0}))
    with pytest.raises(AssertionError, match="Mesh file"):
        ch.main(["--folder", str(tmp_path)])
    assert ch.compute_hemodynamics is ch.compute_hemodyanamics
    assert ch.read_parameters_from_file(tmp_path) == {"save_deg": 2, "mu_f": 1.0}


def test_rank_and_world_from_mpi_and_slurm_launchers(monkeypatch):
    """The reference is started with `mpirun -np N vasp-compute-hemo` (docs/postprocess.md:165); the same command line
    must find its rank without mpi4py: Open MPI, MPICH/hydra and srun variables, torchrun's taking precedence."""
    from vasp_b200 import timeshard
    names = [n for trio in timeshard._LAUNCHERS for n in trio] + ["PMIX_NAMESPACE", "SLURM_JOB_ID", "SLURM_STEP_ID",
                                                                 "OMPI_MCA_ess_base_jobid", "PMI_JOBID",
                                                                 "TORCHELASTIC_RUN_ID", "MASTER_PORT", "SLURM_NTASKS"]
    for n in names:
        monkeypatch.delenv(n, raising=False)
    assert timeshard.env_rank_world() == (0, 0, 1)
    plain = timeshard._rendezvous_path(2, "/tmp")
    monkeypatch.setenv("OMPI_COMM_WORLD_RANK", "3")
    monkeypatch.setenv("OMPI_COMM_WORLD_SIZE", "8")
    monkeypatch.setenv("OMPI_COMM_WORLD_LOCAL_RANK", "1")
    assert timeshard.env_rank_world() == (3, 1, 8)
    monkeypatch.setenv("PMIX_NAMESPACE", "prterun-node-123@1")
    a = timeshard._rendezvous_path(8, "/tmp")
    assert "prterun-node-123@1" in a.name and a != plain
    monkeypatch.setenv("SLURM_PROCID", "5")
    monkeypatch.setenv("SLURM_NTASKS", "6")
    monkeypatch.setenv("SLURM_STEP_NUM_TASKS", "6")
    assert timeshard.env_rank_world() == (3, 1, 8)            # first launcher in the table wins
    for n in ("OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_RANK", "PMIX_NAMESPACE"):
        monkeypatch.delenv(n)
    assert timeshard.env_rank_world() == (5, 5, 6)            # no local id given: the rank itself
    monkeypatch.delenv("SLURM_STEP_NUM_TASKS")
    assert timeshard.env_rank_world() == (0, 0, 1)            # batch-script environment (no srun step): a plain run
    monkeypatch.setenv("SLURM_STEP_NUM_TASKS", "6")
    monkeypatch.setenv("SLURM_JOB_ID", "77")
    monkeypatch.setenv("SLURM_STEP_ID", "0")
    assert timeshard._rendezvous_path(6, "/tmp").name == "vasp_b200_nccl_0_77_0_6.id"
    monkeypatch.setenv("PMI_RANK", "1")
    monkeypatch.setenv("PMI_SIZE", "2")
    assert timeshard.env_rank_world() == (1, 1, 2)
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("LOCAL_RANK", "0")
    assert timeshard.env_rank_world() == (0, 0, 4)            # torchrun's contract first


def test_bench_series_is_one_series_whatever_the_sharding():
    """bench.py: every rank synthesises its own snapshots of ONE long series (weak: rank r owns [r n, (r + 1) n) plus a
    halo; strong: plan_shard ranges), and rank 0 re-synthesises the whole of it for the parity check -- the pieces must
    be the series, bit for bit."""
    import bench
    whole, dt = bench.series_coefficients(37, 0, 20)
    for first, n in ((0, 9), (8, 11), (19, 18)):
        part, dt_p = bench.series_coefficients(n, first, 20)
        assert dt_p == dt and np.array_equal(part, whole[first:first + n])
    sh = [timeshard.plan_shard(37, r, 4) for r in range(4)]
    assert [s.start for s in sh] == [0, 10, 19, 28] and sh[-1].stop == 37
    assert bench.algorithmic_bytes_per_unit(2) == 240 and bench.algorithmic_bytes_per_unit(1, True) == 168
    assert set(bench.OTHER_WORKLOADS) <= set(bench.WORKLOADS) and all(r >= 256 for r, _ in bench.OTHER_WORKLOADS.values())


def test_reference_arm_of_the_bench_prints_the_contract_line():
    """``bench.py --impl reference`` (the CPU arm the driver runs beside ours: the C twin of the oracle on the host cores,
    a bounded sample of the same workload) prints ONE JSON line with the keys of the bench contract; under a launcher
    only rank 0 prints."""
    import subprocess
    import sys
    env = {k: v for k, v in __import__("os").environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--snapshots", "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "wall_facet_snapshots_per_s" and d["higher_is_better"] is True
    assert d["unit"] == "facet*snapshots/s" and d["value"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("stenosis_p1") and d["config"]["facets"] == 2560
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "snapshots" in d["cpu_baseline"]["sample"]
    rank1 = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300,
                           env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert rank1.returncode == 0 and not [ln for ln in rank1.stdout.splitlines() if ln.startswith("{")]
