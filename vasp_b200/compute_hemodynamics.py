"""Drop-in for VaSP's ``vasp-compute-hemo`` / ``compute_hemodyanamics()`` with the numerics on a B200.

Mirrors ``src/vasp/postprocessing/postprocessing_fenics/compute_hemodynamics.py`` of the reference: same function
name (misspelling included) and signature (``:160-161``), same command-line flags
(``postprocessing_fenics_common.py:17-28``), same input files (``<stem>_fluid.h5``, ``<stem>_refined_fluid.h5``,
``Visualization_separate_domain/u.h5``, ``Checkpoint/default_variables.json``), same outputs
(``Hemodynamic_indices/{RRT,OSI,ECAP,WSS,TAWSS,TWSSG}.xdmf|.h5``), same stdout progress lines and the same error
conventions (``AssertionError`` for missing inputs / ``save_deg``, ``RuntimeError`` for unreadable parameters, the OSI
range assertion after the files are written, ``:366-372``).

One extension of the input side: when ``Visualization_separate_domain/`` is missing the reference
first converts the raw turtleFSI output with ``create_hdf5()`` (``:389-431``); here the raw
``Visualization/velocity*.h5`` arrays are read in place (:mod:`vasp_b200.io_turtle`) and sliced on the GPU.

What changed underneath: dolfin's per-snapshot assemble + LU + Python dof matching is replaced by the CUDA engine
(:mod:`vasp_b200.engine`); snapshots are ``pread`` from ``u.h5`` into pinned buffers by a reader thread while the
GPU works on the previous block; under a launcher (one process per GPU, torchrun-style ``RANK``/``WORLD_SIZE``)
every rank takes a contiguous time range and the partial sums meet in one NCCL all-reduce.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, Optional, Union

import numpy as np

from . import io_dolfin, io_turtle
from .engine import HemoEngine, device_count, pinned_empty
from .timeshard import NcclComm, env_rank_world, plan_shard

INDEX_NAMES = ["RRT", "OSI", "ECAP", "WSS", "TAWSS", "TWSSG"]  # compute_hemodynamics.py:253


def parse_arguments(argv=None) -> argparse.Namespace:
    """The reference's flag set (``postprocessing_fenics_common.py:10-28``) plus two extensions that default to the
    reference behaviour."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--folder", type=Path, help="Path to simulation results")
    parser.add_argument('--mesh-path', type=Path, default=None,
                        help="Path to the mesh file (default: <folder_path>/Mesh/mesh.h5)")
    parser.add_argument("--stride", type=int, default=1, help="Save frequency of simulation")
    parser.add_argument("-st", "--start-time", type=float, default=None, help="Desired start time for postprocessing")
    parser.add_argument("-et", "--end-time", type=float, default=None, help="Desired end time for postprocessing")
    parser.add_argument("--extract-entire-domain", action="store_true", help="Extract displacement from entire domain")
    parser.add_argument("--log-level", type=int, default=20,
                        help="Specify the log level (default is 20, which is INFO)")
    # extensions
    parser.add_argument("--velocity-degree", type=int, choices=[1, 2], default=2,
                        help="2 (reference behaviour: P1 data on the refined mesh = P2 on the mesh) or 1 (P1 data on "
                             "the un-refined mesh; the reference refuses save_deg != 2)")
    parser.add_argument("--device", type=int, default=None, help="CUDA device (default: LOCAL_RANK)")
    parser.add_argument("--derive-refined-mesh", action="store_true",
                        help="raw turtleFSI input only: take the node coordinates from Visualization/velocity.h5 "
                             "instead of Mesh/mesh_refined.h5 and Mesh/mesh_refined_fluid.h5 (vasp-refine-mesh and "
                             "vasp-separate-mesh need not have been run on the refined mesh)")
    return parser.parse_args(argv)


def read_parameters_from_file(folder: Union[str, Path]) -> Optional[Dict]:
    """``postprocessing_common.py:124-145``."""
    file_path = Path(folder) / "Checkpoint" / "default_variables.json"
    try:
        with open(file_path, 'r') as json_file:
            return json.load(json_file)
    except FileNotFoundError:
        logging.error(f"File not found: {file_path}")
        return None
    except json.JSONDecodeError as e:
        logging.error(f"Error parsing JSON file: {e}")
        return None


class _Drain:
    """One worker thread that stores finished WSS blocks in submission order while the caller computes the next
    one; ``wait_free(slot)`` blocks until the block last submitted from ``slot`` has been stored."""

    def __init__(self, store):
        self._store, self.seconds = store, 0.0
        self._q: "queue.Queue" = queue.Queue()
        self._free = [threading.Event(), threading.Event()]
        for e in self._free:
            e.set()
        self._error: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            slot, buf, first_step, n_steps = item
            try:
                if self._error is None:
                    t0 = time.perf_counter()
                    self._store(buf, first_step, n_steps)
                    self.seconds += time.perf_counter() - t0
            except BaseException as e:  # surfaced in the caller
                self._error = e
            finally:
                self._free[slot].set()

    def _check(self) -> None:
        if self._error is not None:
            raise self._error

    def wait_free(self, slot: int) -> None:
        self._free[slot].wait()
        self._check()

    def submit(self, slot: int, buf, first_step: int, n_steps: int) -> None:
        self._check()
        self._free[slot].clear()
        self._q.put((slot, buf, first_step, n_steps))

    def close(self) -> None:
        self._q.put(None)
        self._thread.join()
        self._check()


def default_block_snapshots(vec_len: int, compact_len: int = 0, n_facets: int = 0) -> int:
    """Snapshots per pinned buffer: ~8 MiB of whichever is larger per snapshot -- what lands in the read buffer
    (``compact_len`` doubles when the wall layer is gathered on the way, else the whole vector) or the WSS block that
    comes back (72 bytes per facet) -- at least 2 and at most 512.  Small blocks matter twice: nothing overlaps the
    first block's read, and pinning memory costs about as much per byte as reading it (measured: 2 x 260 MB buffers
    cost more than reading the 1 GB file from the page cache, profiles/r1io_*; 64-snapshot blocks of compact rows made
    the 2 M-tet entry point 0.17 s slower than 9-snapshot ones, profiles/r2h_*).  The kernels do not care: a push of a
    few columns costs what one 64-column pass costs, far less than reading those snapshots."""
    row = max((compact_len or vec_len) * 8, 72 * n_facets)
    return int(min(512, max(2, (8 << 20) // row)))


class _BlockReader:
    """Reads snapshot blocks of ``u.h5`` into two pinned buffers, one block ahead of the consumer.

    With an ``engine`` whose wall-layer compaction is active the blocks hold COMPACT rows: the engine's thread pool
    gathers the wall-layer dofs straight out of the read-only mapping of the file (the page cache), so a snapshot is
    never copied whole and only ``24 * n_wall_nodes`` bytes per snapshot cross the bus (``compact`` is then True and
    the consumer pushes with ``push_compact``)."""

    def __init__(self, series, first: int, last: int, block: int, engine=None):
        self.series, self.block = series, max(2, block)
        self.compact = bool(engine is not None and engine.compaction_active and hasattr(series, "row_addresses"))
        self._engine = engine
        self.ranges = []
        pos = first
        while pos < last:
            end = min(pos + self.block, last)
            if last - end == 1:  # never leave a single trailing snapshot (halo blocks need two)
                end = last
            self.ranges.append((pos, end))
            pos = end
        self.max_rows = rows = max((b - a for a, b in self.ranges), default=1)
        if self.compact:
            self._addr = series.row_addresses()
            self.bufs = [pinned_empty((rows, engine.compact_len)) for _ in range(2)]
        else:
            self.bufs = [pinned_empty((rows, series.vec_len)) for _ in range(2)]
        self.io_seconds = 0.0
        # page cache -> pinned memory is a memcpy per pread: one thread moves ~7 GB/s, the PCIe link takes 52
        n_threads = int(os.environ.get("VASP_B200_READ_THREADS", min(4, os.cpu_count() or 1)))
        self._pool = ThreadPoolExecutor(n_threads) if n_threads > 1 and not self.compact else None
        self._ready = [threading.Event(), threading.Event()]
        self._free = [threading.Event(), threading.Event()]
        for e in self._free:
            e.set()
        self._error: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _read(self, buf, a: int, b: int, nxt) -> None:
        if self.compact:
            if nxt is not None:
                self.series.advise(*nxt)  # the kernel reads ahead while this block is gathered
            self._engine.compact_rows(self._addr[a:b], buf)
        else:
            self.series.read_into(buf, a, b, self._pool)

    def _run(self) -> None:
        try:
            for i, (a, b) in enumerate(self.ranges):
                slot = i & 1
                self._free[slot].wait()
                self._free[slot].clear()
                t0 = time.perf_counter()
                self._read(self.bufs[slot], a, b, self.ranges[i + 1] if i + 1 < len(self.ranges) else None)
                self.io_seconds += time.perf_counter() - t0
                self._ready[slot].set()
        except BaseException as e:  # surfaced in the consumer
            self._error = e
            for e2 in self._ready:
                e2.set()
        finally:
            if self._pool is not None:
                self._pool.shutdown(wait=False)

    def __iter__(self):
        for i, (a, b) in enumerate(self.ranges):
            slot = i & 1
            self._ready[slot].wait()
            self._ready[slot].clear()
            if self._error is not None:
                raise self._error
            yield a, b, self.bufs[slot][:b - a]
            self._free[slot].set()


def compute_hemodyanamics(visualization_separate_domain_folder: Path, mesh_path: Path,
                          mu_f: float, stride: int = 1, velocity_degree: int = 2,
                          device: Optional[int] = None, block_snapshots: Optional[int] = None,
                          series=None, wss_matrix_folder: Optional[Path] = None) -> None:
    """
    Compute hemodynamic indices from velocity field (reference ``compute_hemodynamics.py:160-372``).

    Args:
        visualization_separate_domain_folder (Path): Path to the folder containing u.h5
        mesh_path (Path): Path to the mesh folder
        mu_f (float): Dynamic viscosity
        stride (int): Save frequency of output data
        series: (extension) an already opened velocity series, e.g. :class:`io_turtle.TurtleVelocitySeries` over the
            raw turtleFSI output; ``u.h5`` is not needed then and ``stride`` has been applied by whoever opened it
        wss_matrix_folder: (extension) also write ``wss_mag.npz`` there: the (dof x time) matrix the
            spectral tools otherwise rebuild from ``WSS.h5`` (``create_transformed_matrix``, quantity ``"wss"``, whole
            series, stride 1).  K2 then writes tau time-major and the per-step vectors of ``WSS.h5`` are its columns
    """
    rank, local_rank, world = env_rank_world()
    visualization_separate_domain_folder = Path(visualization_separate_domain_folder)
    mesh_path = Path(mesh_path)
    t_begin = time.perf_counter()
    if series is None:
        file_path_u = visualization_separate_domain_folder / "u.h5"
        assert file_path_u.exists(), f"Velocity file {file_path_u} not found.  Make sure to run create_hdf5.py first."
        series = io_dolfin.VelocitySeries(file_path_u, "velocity", stride)

    if rank == 0:
        print("--- Read the original mesh and also the refined mesh \n")
    mesh_name = mesh_path.stem
    fluid_mesh_path = mesh_path.parent / f"{mesh_name}_fluid.h5"
    assert fluid_mesh_path.exists(), f"Mesh file {fluid_mesh_path} not found."
    xyz, tets = io_dolfin.read_mesh(fluid_mesh_path, "mesh")

    if device is None:
        # one process per GPU: the launcher's local rank picks the device.  Under `srun --gpus-per-task=1` every task
        # sees a single device 0 whatever its local id is, hence the modulo.
        device = local_rank % max(device_count(), 1)
    eng = HemoEngine(device)
    eng.set_mesh(xyz, tets)  # BoundaryMesh(mesh, "exterior") and all index maps, on the device

    if rank == 0:
        print("--- Define function spaces \n")
    if velocity_degree == 2 and getattr(series, "geometry", None) is not None:
        # (extension) the raw turtleFSI arrays carry the geometry they are indexed by: the P2 nodes of
        # the wall cells are matched against it directly and <mesh>_refined_fluid.h5 is not needed
        rxyz = series.geometry
        comp_offset, node_stride, perm = series.layout(None, len(rxyz))
        eng.set_velocity_layout(2, refined_xyz=rxyz, node_perm=perm, comp_offset=comp_offset, node_stride=node_stride)
    elif velocity_degree == 2:
        refined_mesh_path = mesh_path.parent / f"{mesh_name}_refined_fluid.h5"
        assert refined_mesh_path.exists(), f"Mesh file {refined_mesh_path} not found."
        rxyz, rtets = io_dolfin.read_mesh(refined_mesh_path, "mesh")
        comp_offset, node_stride, perm = series.layout(rtets, len(rxyz))
        eng.set_velocity_layout(2, refined_xyz=rxyz, node_perm=perm, comp_offset=comp_offset, node_stride=node_stride)
    else:
        comp_offset, node_stride, perm = series.layout(tets, len(xyz))
        eng.set_velocity_layout(1, n_nodes=len(xyz), node_perm=perm, comp_offset=comp_offset, node_stride=node_stride)
    if rank == 0:
        print("--- Define functions")

    maps = eng.maps()
    hemodynamic_indices_path = visualization_separate_domain_folder.parent / "Hemodynamic_indices"
    hemodynamic_indices_path.mkdir(parents=True, exist_ok=True)
    bgeom = xyz[maps["bvert_parent"]]

    if rank == 0:
        print("=" * 10, "Start post processing", "=" * 10)

    n_snap = len(series)
    assert n_snap >= 2, "at least two snapshots are needed (dt is the gap between the first two)"
    # Get time difference between two consecutive time steps
    dt = float(series.timestamps[1] - series.timestamps[0])
    eng.begin(mu_f, dt)
    # the NCCL id travels through the output folder: every rank must see it anyway (WSS shards are merged from there)
    comm = NcclComm(eng, rank, world, rendezvous_dir=str(hemodynamic_indices_path)) if world > 1 else None

    shard = plan_shard(n_snap, rank, world)
    nF = eng.nF
    compacting = eng.compaction_active and hasattr(series, "row_addresses")
    if block_snapshots is None:
        block_snapshots = default_block_snapshots(series.vec_len, eng.compact_len if compacting else 0, nF)
        if wss_matrix_folder is not None:
            # the matrix comes back as a pitched copy of 9 nF rows of (block x 8) bytes: keep the rows >= 256 bytes
            # unless that would pin more than 1 GiB per read buffer
            block_snapshots = max(block_snapshots, min(32, max(2, (1 << 30) // (series.vec_len * 8))))
    eng.set_tuning(batch_snapshots=block_snapshots, chunk_snapshots=0)

    wss_writer = None
    shard_file = None
    if rank == 0:
        wss_writer = io_dolfin.CheckpointWriter(hemodynamic_indices_path, "WSS", maps["btopology"], bgeom, True)
    elif shard.count:
        shard_file = np.lib.format.open_memmap(hemodynamic_indices_path / f".WSS_shard{rank}.npy", mode="w+",
                                               dtype=np.float64, shape=(shard.count, nF, 3, 3))
    reader = _BlockReader(series, shard.read_start, shard.stop, block_snapshots, eng)
    direct = None
    if wss_matrix_folder is not None:
        assert world == 1, "wss_matrix_folder: the direct WSS matrix is assembled by a single rank"
        from .wss_matrix import DirectWssMatrix, write_npz
        ts = [float(t) for t in series.timestamps]
        direct = DirectWssMatrix(eng, ts, ts[0], ts[-1], 1)
        direct.attach()
        wss_buf = None
    else:
        wss_buf = pinned_empty((reader.max_rows, nF, 3, 3))  # the reader may merge a trailing snapshot into a block

    first, done = True, 0
    t_setup = time.perf_counter() - t_begin
    t_push = 0.0

    # The WSS steps of block i are written (WSS.h5 on rank 0, the shard file elsewhere) by a thread while block i + 1
    # is on the GPU: two result buffers, one being filled by the push, one being drained.
    def store(buf, first_step, n_steps):
        k0 = shard.start + first_step
        ts = [float(t) for t in series.timestamps[k0:k0 + n_steps]]
        if rank == 0:
            print("".join(f"========== Calculating WSS at Timestep: {t} ==========\n" for t in ts), end="")
            # Write temporal WSS: the whole block of steps in one write
            wss_writer.write_block(buf[:n_steps], ts)
        else:
            shard_file[first_step:first_step + n_steps] = buf[:n_steps]

    drain = _Drain(store)
    wss_bufs = [wss_buf, pinned_empty(wss_buf.shape)] if wss_buf is not None else [None, None]
    for i, (a, b, u) in enumerate(reader):
        flags = shard.first_push_flags() if first else 0
        n_real = (b - a) - (1 if (first and shard.has_halo) else 0)
        slot = i & 1
        drain.wait_free(slot)  # the block written two pushes ago has left this buffer
        t0 = time.perf_counter()
        push = eng.push_compact if reader.compact else eng.push
        if direct is not None:
            m = push(u, flags=flags)  # columns [done, done + n_real) of the time-major matrix
            out_block = np.ascontiguousarray(m[:, done:done + n_real].T).reshape(n_real, nF, 3, 3)
        else:
            push(u, flags=flags, wss_out=wss_bufs[slot])
            out_block = wss_bufs[slot]
        t_push += time.perf_counter() - t0
        drain.submit(slot, out_block, done, n_real)
        done += n_real
        first = False
    drain.close()
    t_write = drain.seconds
    series_io = reader.io_seconds
    if direct is not None:
        direct.detach()
        write_npz(wss_matrix_folder, direct.matrix())
        if rank == 0:
            print(f"--- wss_mag.npz is saved in {wss_matrix_folder}")
    if shard_file is not None:
        shard_file.flush()
        del shard_file
    if comm is not None:
        comm.barrier()  # every rank's WSS block is on disk before rank 0 starts merging them
    # the single collective (15*nF partial sums + the snapshot count) happens below, fused with the final formulas

    if rank == 0:
        # WSS of the other ranks' time ranges, in order
        for r in range(1, world):
            sh = plan_shard(n_snap, r, world)
            part = hemodynamic_indices_path / f".WSS_shard{r}.npy"
            if sh.count:
                arr = np.load(part, mmap_mode="r")
                for i in range(0, sh.count, 256):
                    j = min(i + 256, sh.count)
                    wss_writer.write_block(arr[i:j], [float(t) for t in series.timestamps[sh.start + i:sh.start + j]])
                del arr
                part.unlink()
        wss_writer.close()
        print("=" * 10, "Saving hemodynamic indices", "=" * 10)

    out = comm.reduce_finalize(n_snap) if comm is not None else eng.finalize(n_snap)
    if comm is not None:
        comm.close()  # ranks > 0 are done here; rank 0 goes on to write the index files
    timers = eng.timers()
    series.close()
    if rank == 0:
        # Write indices to file
        for name in INDEX_NAMES:
            if name == "WSS":
                continue
            w = io_dolfin.CheckpointWriter(hemodynamic_indices_path, name, maps["btopology"], bgeom, False)
            w.write(out[name], 0)
            w.close()
            print(f"--- {name} is saved in {hemodynamic_indices_path}")
        total = time.perf_counter() - t_begin
        print(f"--- timing: total {total:.3f} s | mesh + maps {t_setup:.3f} s | u.h5 -> pinned host {series_io:.3f} s "
              f"(reader thread) | push calls {t_push:.3f} s (host -> device {timers['h2d_ms'] * 1e-3:.3f} s, kernels "
              f"{timers['kernel_ms'] * 1e-3:.3f} s) | WSS.h5 steps {t_write:.3f} s | "
              f"{nF} wall facets x {n_snap} snapshots on {world} GPU(s)")
    eng.close()

    if rank == 0:
        # assert that OSI is within 0 to 0.5
        min_, max_ = np.min(out["OSI"]), np.max(out["OSI"])
        tol = 1e-12
        assert -tol <= min_ < 0.5, "OSI min is not within 0 to 0.5"
        assert -tol < max_ <= 0.5 + tol, "OSI max is not within 0 to 0.5"


# correctly spelt alias
compute_hemodynamics = compute_hemodyanamics


def main(argv=None) -> None:
    rank, _, world = env_rank_world()
    if world == 1:
        print("--- Running in serial mode, you can use MPI to speed up the postprocessing. \n")

    args = parse_arguments(argv)
    folder_path = Path(args.folder)

    assert folder_path.exists(), f"Folder {folder_path} not found."

    visualization_separate_domain_folder = folder_path / "Visualization_separate_domain"
    parameters = read_parameters_from_file(args.folder)
    if parameters is None:
        raise RuntimeError("Error reading parameters from file.")

    series = None
    if visualization_separate_domain_folder.exists():
        if rank == 0:
            print("--- Visualization_separate_domain folder found \n")
    else:
        if rank == 0:
            print("--- Visualization_separate_domain folder not found \n")
        # The reference converts the raw turtleFSI output to u.h5 / d_solid.h5 here (create_hdf5(),
        # compute_hemodynamics.py:389-431) and then reads u.h5 back.  Same parameters, same mesh checks, same step
        # selection -- but the fluid-node slice of every raw array is taken on the GPU and nothing is re-written
        # (SURVEY.md §8f-1).  The displacement file of the solid post-processing is not this path's business:
        # vasp-create-hdf5 still makes it.
        visualization_path = folder_path / "Visualization"
        save_deg = parameters["save_deg"]
        dt = parameters["dt"]
        save_step = parameters["save_step"]
        save_time_step = dt * save_step
        logging.info(f"save_time_step: {save_time_step} \n")
        fluid_domain_id = parameters["dx_f_id"]
        solid_domain_id = parameters["dx_s_id"]
        logging.info(f"--- Fluid domain ID: {fluid_domain_id} and Solid domain ID: {solid_domain_id} \n")
        if args.derive_refined_mesh:
            domain_mesh_path = None
            logging.info("--- Taking the node coordinates from the velocity file \n")
        elif args.mesh_path:
            domain_mesh_path = Path(args.mesh_path)
            logging.info("--- Using user-defined mesh \n")
            assert domain_mesh_path.exists(), f"Mesh file {domain_mesh_path} not found."
        elif save_deg == 2:
            domain_mesh_path = folder_path / "Mesh" / "mesh_refined.h5"
            logging.info("--- Using refined mesh \n")
            assert domain_mesh_path.exists(), f"Mesh file {domain_mesh_path} not found."
        else:
            domain_mesh_path = folder_path / "Mesh" / "mesh.h5"
            logging.info("--- Using non-refined mesh \n")
            assert domain_mesh_path.exists(), f"Mesh file {domain_mesh_path} not found."
        if rank == 0:
            print(f"save_time_step: {save_time_step} \n")
            print("--- Reading the fluid velocity straight from Visualization/velocity.h5 (no u.h5 is written) \n")
        series = io_turtle.TurtleVelocitySeries(visualization_path, domain_mesh_path, save_time_step, args.stride,
                                                args.start_time, args.end_time, fluid_domain_id, solid_domain_id,
                                                **({"derive_refined_mesh": True} if args.derive_refined_mesh else {}))

    save_deg = parameters["save_deg"]
    if args.velocity_degree == 2:
        assert save_deg == 2, "This script only works for save_deg = 2"
    mu_f = parameters["mu_f"]

    if isinstance(mu_f, list):
        if rank == 0:
            print("--- two fluid regions are detected. Using the first fluid region for viscosity \n")
        mu_f = mu_f[0]

    if args.mesh_path:
        mesh_path = Path(args.mesh_path)
        if rank == 0:
            print("--- Using user-defined mesh \n")
        assert mesh_path.exists(), f"Mesh file {mesh_path} not found."
    else:
        mesh_path = folder_path / "Mesh" / "mesh.h5"
        if rank == 0:
            print("--- Using mesh from default turrtleFSI Mesh folder \n")
        assert mesh_path.exists(), f"Mesh file {mesh_path} not found."

    compute_hemodyanamics(visualization_separate_domain_folder, mesh_path, mu_f, args.stride,
                          velocity_degree=args.velocity_degree, device=args.device, series=series)


if __name__ == "__main__":
    main()
