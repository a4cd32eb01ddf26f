"""N > 1 host path on CPU: two processes over gloo run the product's shard plan / push loop / reduction contract
with the C oracle standing in for the GPU engine, and must reproduce the sequential result."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch.distributed as dist
from tests import helpers as H
from oracle import c_oracle, hemo_oracle as ho
from vasp_b200 import timeshard

class OracleEngine:
    """CPU stand-in with HemoEngine's push / sums / set_sums contract (tests only)."""
    def __init__(self, stress, dt, n_nodes):
        self.co, self.dt, self.n = c_oracle.COracle(stress), dt, n_nodes
        self.nF = stress.nF
        self._s = np.zeros((15, self.nF)); self._count = 0; self._prev = None
    def push(self, u, flags=0, keep_wss=False, wss_out=None):
        u = np.ascontiguousarray(u)
        off = (0, self.n, 2 * self.n)
        if flags & timeshard.PUSH_HALO_FIRST:
            self._prev = self.co.run(u[:1], self.dt, off)["tau_last"]
            u = u[1:]
        elif flags & timeshard.PUSH_GLOBAL_FIRST:
            self._prev = None
        r = self.co.run(u, self.dt, off, tau_prev=self._prev, keep_wss=keep_wss)
        self._prev = r["tau_last"]
        self._s[:9] += r["wss_sum"].reshape(self.nF, 9).T
        self._s[9:12] += r["tawss_sum"].T
        self._s[12:] += r["twssg_sum"].T
        self._count += len(u)
        return r.get("wss")
    def sums(self): return self._s.copy(), self._count
    def set_sums(self, s, c): self._s, self._count = np.array(s), int(c)

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
src = H.load_fluid("stenosis")
order = int(os.environ["ORDER"])
case = H.make_case(src["xyz"], src["tets"], order, n_snap=13)
S, res, fin = H.oracle_run(case, 3.5e-3)
eng = OracleEngine(S, case["dt"], case["n_nodes"])
shard = timeshard.plan_shard(13, rank, world)
seen = []
done = timeshard.run_shard(eng, shard, lambda a, b: case["u"][a:b], block=3,
                           on_wss=lambda k, w: seen.append((k, len(w))))
assert done == shard.count
assert sum(n for _, n in seen) == shard.count and seen[0][0] == shard.start
comm = timeshard.TorchDistComm(eng)
comm.allreduce_sums()
comm.barrier()
s, c = eng.sums()
assert c == 13
out = ho.finalize(s[:9].T.reshape(-1, 3, 3), s[9:12].T, s[12:].T, c)
for name in H.FIELDS:
    assert H.rel_l2(out[name], fin[name]) < 1e-12, name
assert abs(comm.max(float(rank)) - (world - 1)) == 0
dist.destroy_process_group()
print("ok", rank)
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(order: int):
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), REPO_ROOT=str(ROOT), ORDER=str(order), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert all("ok" in o for o in outs)


def test_two_rank_time_shards_reproduce_sequential_p2():
    _run(2)


def test_two_rank_time_shards_reproduce_sequential_p1():
    _run(1)


def test_unique_id_exchange_through_a_file(tmp_path, monkeypatch):
    from vasp_b200 import timeshard
    monkeypatch.setenv("MASTER_PORT", "12345")
    uid = bytes(range(128))
    got0 = timeshard.exchange_unique_id(0, 2, lambda: uid, directory=str(tmp_path))
    got1 = timeshard.exchange_unique_id(1, 2, lambda: b"", directory=str(tmp_path), timeout=5)
    assert got0 == got1 == uid
    timeshard.cleanup_unique_id(2, directory=str(tmp_path))
    assert not list(tmp_path.iterdir())


CLI_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch.distributed as dist
from tests.fake_engine import OracleHemoEngine
from vasp_b200 import compute_hemodynamics as ch, engine as engine_mod, timeshard

dist.init_process_group("gloo")
ch.HemoEngine = OracleHemoEngine
ch.pinned_empty = engine_mod.pinned_empty = lambda shape: np.zeros(shape)
ch.device_count = lambda: 1
ch.NcclComm = lambda eng, rank, world, rendezvous_dir=None: timeshard.TorchDistComm(eng)
ch.main(["--folder", os.environ["FOLDER"]])
dist.destroy_process_group()
'''


def test_entry_point_under_two_ranks_merges_the_wss_series(tmp_path):
    """The multi-rank plumbing of the entry point on CPU (stand-in engine, gloo communicator): time shards with a halo
    snapshot, per-rank WSS blocks merged by rank 0 after a barrier, one reduction -- same six files as one rank."""
    import shutil
    from tests.test_gpu_cli import _make_folder
    from tests.test_cli_host_logic import _u_syn
    from vasp_b200 import io_dolfin
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir()
    _make_folder(one, _u_syn(9), 11, 0.04, 3.5e-3)
    shutil.copytree(one, two)
    port = _free_port()

    def launch(folder, world):
        procs = []
        for rank in range(world):
            env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                       MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port + world), REPO_ROOT=str(ROOT), FOLDER=str(folder),
                       OMP_NUM_THREADS="1")
            procs.append(subprocess.Popen([sys.executable, "-c", CLI_WORKER], env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o
        return outs

    launch(one, 1)
    outs = launch(two, 2)
    assert "Calculating WSS at Timestep" in outs[0] and "Calculating WSS at Timestep" not in outs[1]
    h1, h2 = one / "Hemodynamic_indices", two / "Hemodynamic_indices"
    assert sorted(f.name for f in h2.iterdir()) == sorted(f.name for f in h1.iterdir())   # no shard files left behind
    for k in range(11):
        a, b = io_dolfin.read_checkpoint(h1, "WSS", k), io_dolfin.read_checkpoint(h2, "WSS", k)
        assert np.array_equal(a["values"], b["values"]), k
    assert (h1 / "WSS.xdmf").read_text() == (h2 / "WSS.xdmf").read_text()
    for name in ("TAWSS", "OSI", "RRT", "ECAP", "TWSSG"):
        a, b = io_dolfin.read_checkpoint(h1, name, 0)["values"], io_dolfin.read_checkpoint(h2, name, 0)["values"]
        assert np.linalg.norm(a - b) <= 1e-12 * np.linalg.norm(a), name
