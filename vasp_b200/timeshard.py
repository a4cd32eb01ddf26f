"""Multi-GPU time sharding: contiguous snapshot ranges per rank, one halo snapshot, one all-reduce.

The reference's snapshot loop is strictly sequential (``compute_hemodynamics.py:272-318``); the only coupling
between snapshots is additive (``TAWSS``, ``WSS_mean``, ``TWSSG`` sums, ``:303-312``) plus TWSSG's dependence on
the previous step's tau (``:309,315-316``).  So rank ``g`` takes snapshots ``[k_g, k_{g+1})``, additionally reads
snapshot ``k_g - 1`` to seed ``tau_prev`` (rank 0 starts from zero as the reference does, ``:244``), and the
``15 * nF`` partial sums are added across ranks once before the final formulas (``:326-346``).  SURVEY.md §8e.

One process per GPU.  The launcher contract is torchrun's environment (``RANK``, ``LOCAL_RANK``, ``WORLD_SIZE``,
``MASTER_PORT``) but torch itself is not needed: the NCCL unique id travels through a file on the node.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Optional, Protocol, Tuple

import numpy as np

from ._lib import PUSH_GLOBAL_FIRST, PUSH_HALO_FIRST


# (rank, world size, local rank) variables of the launchers a VaSP user may start the tool with: torchrun's contract
# first, then Open MPI (the reference is run as `mpirun -np N vasp-compute-hemo`, docs/postprocess.md:165), MPICH /
# Intel MPI (hydra), Slurm's srun.  mpi4py is not needed: the ranks only have to know who they are.
_LAUNCHERS = (
    ("RANK", "WORLD_SIZE", "LOCAL_RANK"),
    ("OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_RANK"),
    ("PMI_RANK", "PMI_SIZE", "MPI_LOCALRANKID"),
    # srun only: SLURM_STEP_NUM_TASKS exists inside a job step, not in the batch script's own environment (where
    # SLURM_NTASKS = n would make a single plain process wait for n - 1 ranks that never start)
    ("SLURM_PROCID", "SLURM_STEP_NUM_TASKS", "SLURM_LOCALID"),
)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the launcher's environment; (0, 0, 1) when run plainly."""
    for r, w, l in _LAUNCHERS:
        if r in os.environ and w in os.environ:
            rank, world = int(os.environ[r]), int(os.environ[w])
            return rank, int(os.environ.get(l, str(rank))), world
    return 0, 0, 1


def _job_tag() -> str:
    """Something all ranks of one job share and no other job has (names the rendezvous file of the NCCL id)."""
    for names in (("PMIX_NAMESPACE",), ("OMPI_MCA_ess_base_jobid",), ("PMI_JOBID",), ("SLURM_JOB_ID", "SLURM_STEP_ID")):
        if all(n in os.environ for n in names):
            return "_".join(os.environ[n] for n in names).replace("/", "-")
    # torchrun and plain subprocess launchers: all workers share their parent process
    return f"{os.getppid()}_{os.environ.get('TORCHELASTIC_RUN_ID', 'none')}"


@dataclass(frozen=True)
class Shard:
    """Snapshots ``[start, stop)`` of the selected series belong to this rank; ``read_start`` is where reading
    begins (one earlier than ``start`` when a halo snapshot is needed)."""
    rank: int
    world: int
    start: int
    stop: int

    @property
    def has_halo(self) -> bool:
        return self.start > 0 and self.stop > self.start

    @property
    def read_start(self) -> int:
        return self.start - 1 if self.has_halo else self.start

    @property
    def count(self) -> int:
        return self.stop - self.start

    def first_push_flags(self) -> int:
        return PUSH_HALO_FIRST if self.has_halo else PUSH_GLOBAL_FIRST


def plan_shard(n_snap: int, rank: int, world: int) -> Shard:
    """Contiguous, balanced ranges: the first ``n_snap % world`` ranks get one extra snapshot."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_snap, world)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return Shard(rank, world, start, stop)


class Engine(Protocol):
    """What :func:`run_shard` needs from a compute engine (``HemoEngine`` satisfies it)."""

    def push(self, u: np.ndarray, flags: int = 0, keep_wss: bool = False, wss_out=None): ...


def run_shard(engine: Engine, shard: Shard, read_block: Callable[[int, int], np.ndarray], block: int,
              on_wss: Optional[Callable[[int, np.ndarray], None]] = None) -> int:
    """Stream this rank's snapshots through ``engine`` in blocks of ``block`` snapshots.

    ``read_block(a, b)`` returns snapshots ``[a, b)`` as a ``(b - a, vec_len)`` float64 array.  ``on_wss(k, tau)``
    receives the per-step WSS of global snapshots ``k, k+1, ...`` when given.  Returns the snapshots processed.
    """
    pos = shard.read_start
    first = True
    done = 0
    while pos < shard.stop:
        end = min(pos + block, shard.stop)
        if first and shard.has_halo and end - pos < 2:
            end = min(pos + 2, shard.stop)  # the halo must travel with at least one real snapshot
        u = read_block(pos, end)
        flags = shard.first_push_flags() if first else 0
        wss = engine.push(u, flags=flags, keep_wss=on_wss is not None)
        n_real = (end - pos) - (1 if (first and shard.has_halo) else 0)
        if on_wss is not None and n_real:
            on_wss(shard.start + done, wss)
        done += n_real
        pos = end
        first = False
    return done


# ------------------------------------------------------------------------------------------------------------------
# communicators
# ------------------------------------------------------------------------------------------------------------------
def exchange_unique_id(rank: int, world: int, make_id: Callable[[], bytes], timeout: float = 300.0,
                       directory: Optional[str] = None) -> bytes:
    """Rank 0 creates the NCCL unique id and publishes it through a file; the other ranks of the node poll for it.

    The file name carries the launcher's pid (all workers of one torchrun share their parent) and MASTER_PORT, so
    concurrent or consecutive jobs cannot pick up each other's id; rank 0 removes any stale file first and every
    rank checks a nonce made of the launcher pid and start time."""
    path = _rendezvous_path(world, directory)
    if rank == 0:
        uid = make_id()
        tmp = path.with_suffix(f".tmp{os.getpid()}")
        tmp.write_bytes(uid)
        os.replace(tmp, path)
        return uid
    t0 = time.time()
    my_start = _process_start_time()
    while time.time() - t0 < timeout:
        try:
            st = path.stat()
            # a file older than this job's processes is a leftover of a crashed run with a recycled pid
            if st.st_size == 128 and st.st_mtime >= my_start - 120.0:
                return path.read_bytes()
        except FileNotFoundError:
            pass
        time.sleep(0.02)
    raise TimeoutError(f"rank {rank}: NCCL unique id file {path} did not appear within {timeout}s")


def _rendezvous_path(world: int, directory: Optional[str] = None) -> Path:
    tag = f"{os.environ.get('MASTER_PORT', '0')}_{_job_tag()}"
    d = Path(directory or os.environ.get("VASP_B200_RDZV_DIR", "/tmp"))
    return d / f"vasp_b200_nccl_{tag}_{world}.id"


def _process_start_time() -> float:
    try:
        return os.stat(f"/proc/{os.getpid()}").st_ctime
    except OSError:
        return time.time()


def cleanup_unique_id(world: int, directory: Optional[str] = None) -> None:
    try:
        _rendezvous_path(world, directory).unlink()
    except FileNotFoundError:
        pass


class NcclComm:
    """Device-side reduction (the product path): NCCL for the plumbing, and -- when the ranks share an NVSwitch node
    and can map each other's memory -- the fused peer-memory reduction + final formulas."""

    def __init__(self, engine, rank: int, world: int, peer: bool = True, rendezvous_dir: Optional[str] = None):
        """``rendezvous_dir``: where rank 0 leaves the NCCL unique id for the others.  The default (``/tmp`` or
        ``VASP_B200_RDZV_DIR``) only reaches the ranks of one node; launchers that span nodes (``mpirun``, ``srun``:
        the reference is run as ``mpirun -np N vasp-compute-hemo``, docs/postprocess.md:165) need a directory every
        rank sees -- the entry point passes its output folder, which has to be shared anyway."""
        import socket
        import zlib
        from .engine import HemoEngine, VaspHemoError
        self.engine, self.rank, self.world = engine, rank, world
        try:
            uid = exchange_unique_id(rank, world, HemoEngine.nccl_unique_id, directory=rendezvous_dir)
        except TimeoutError as e:
            raise TimeoutError(f"{e}.  If the ranks run on several hosts, the rendezvous directory "
                               f"({rendezvous_dir or os.environ.get('VASP_B200_RDZV_DIR', '/tmp')}) must be on a file "
                               "system all of them share (set VASP_B200_RDZV_DIR).") from e
        engine.nccl_init(uid, rank, world)
        engine.barrier()
        if rank == 0:
            cleanup_unique_id(world, rendezvous_dir)
        # do all ranks sit on one host?  (CUDA IPC peer mappings -- the fused reduction -- only exist inside a node)
        tag = float(zlib.crc32(socket.gethostname().encode()) % (1 << 24))
        self.one_host = engine.allreduce_max(tag) == -engine.allreduce_max(-tag)
        self.fused = False
        if not self.one_host and rank == 0:
            print("--- ranks span several hosts: reducing with ncclAllReduce (the peer-memory reduction is node-local)")
        if peer and self.one_host and world <= 8 and os.environ.get("VASP_B200_PEER_REDUCE", "1") != "0":
            try:
                engine.peer_init()   # collective: fails on every rank or on none
                self.fused = True
            except VaspHemoError as e:
                if rank == 0:
                    print(f"--- peer-memory reduction unavailable ({e}); using ncclAllReduce")

    def allreduce_sums(self) -> None:
        self.engine.allreduce_sums()

    def reduce_finalize(self, n_total: int, host: bool = True):
        """Global TAWSS/OSI/RRT/ECAP/TWSSG on every rank: fused peer reduction if mapped, else all-reduce + K4."""
        if self.fused:
            if host:
                # ranks may arrive minutes apart (file I/O); wait on the host so that the kernel's bounded wait
                # (VH_PEER_WAIT_NS) only ever covers stream skew
                self.engine.barrier()
            return self.engine.peer_reduce_finalize(n_total, host=host)
        self.engine.allreduce_sums()
        if host:
            return self.engine.finalize(n_total)
        self.engine.finalize_async(n_total)
        return None

    def max(self, value: float) -> float:
        return self.engine.allreduce_max(value)

    def barrier(self) -> None:
        self.engine.barrier()

    def close(self) -> None:
        """Collective teardown while every rank is still alive.  A rank that tears its communicator down after a peer
        process has already exited waits for that peer until NCCL's own timeout (measured: a two-rank run of the
        entry point took 110 s, almost all of it in rank 0's exit), so callers whose ranks finish at different times
        (rank 0 writes the output files) close right after the last collective."""
        self.engine.barrier()
        self.engine.nccl_destroy()
        self.fused = False


class TorchDistComm:
    """Same contract over an initialised ``torch.distributed`` process group (any backend).

    The partial sums make one host round trip; used when the caller already owns a process group (and by the
    world_size-2 ``gloo`` tests, where the engine is a CPU stand-in)."""

    def __init__(self, engine):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed process group is not initialised")
        self.engine, self._dist = engine, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allreduce_sums(self) -> None:
        import torch
        sums, count = self.engine.sums()
        t = torch.from_numpy(np.concatenate([sums.ravel(), [float(count)]]))
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        out = t.numpy()
        self.engine.set_sums(out[:-1].reshape(sums.shape), int(round(out[-1])))

    def reduce_finalize(self, n_total: int, host: bool = True):
        """Same contract as :meth:`NcclComm.reduce_finalize`: global indices on every rank."""
        self.allreduce_sums()
        return self.engine.finalize(n_total)

    def max(self, value: float) -> float:
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t[0])

    def barrier(self) -> None:
        self._dist.barrier()

    def close(self) -> None:
        self._dist.barrier()
