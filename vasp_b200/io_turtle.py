"""turtleFSI's raw ``Visualization/velocity.xdmf`` + ``velocity*.h5`` as the velocity series of the hot path.

The reference converts these files to ``Visualization_separate_domain/u.h5`` first (``create_hdf5()``,
``postprocessing_fenics/create_hdf5.py:26-189``; ``main()`` of ``compute_hemodynamics.py:389-431`` calls it when the
folder is missing): every saved step's ``VisualisationVector/<i>`` array (all nodes of the refined whole-domain mesh,
``(N_all, 3)``) is sliced to the fluid nodes, flattened component-blocked and written again.  Here the slice happens
on the GPU instead: the raw arrays are copied to the device as they are and K1 gathers ``3 * fluid_ids[v] + c`` --
no second copy of the time series on disk.

What is reproduced from the reference, line by line:

* ``output_file_lists`` (``postprocessing_common.py:65-121``): time values, h5 file and dataset index of every step,
  from the XDMF text (restarted simulations spread over several h5 files);
* ``get_domain_ids`` (``postprocessing_common.py:16-62``): ``fluid_ids = unique(topology[domains == dx_f_id])``;
* the step selection of ``create_hdf5`` (``create_hdf5.py:118-131``): ``int(start / save_time_step) - 1`` up to
  ``int(end / save_time_step)`` in steps of ``stride`` -- and then, because ``main()`` hands the same ``--stride`` to
  ``compute_hemodyanamics`` (``:431,455``), every ``stride``-th of *those* (``get_dataset_names(step=stride)``,
  ``:179``).  The double application is the reference's behaviour and is kept.
"""
from __future__ import annotations

import logging
import os
import re
from pathlib import Path
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from .h5lite import H5File


def output_file_lists(xdmf_file: Union[str, Path]) -> Tuple[List[str], List[float], List[int]]:
    """``(h5 file of every step, time value of every step, dataset index of every step)``."""
    lines = Path(xdmf_file).read_text().splitlines()
    checkpoint = any("FiniteElementFunction" in ln for ln in lines)
    names: List[str] = []
    times: List[float] = []
    index: List[int] = []
    for ln in lines:
        if "<Time Value" in ln:
            times.append(float(re.findall('<Time Value="(.+?)"', ln)[0]))
        if checkpoint and "vector" in ln:
            names.append(re.findall(r'"HDF">(.*?):', ln)[0])
            index.append(int(re.findall(r"_([0-9]+)\/vector", ln)[0]))
        elif not checkpoint and "VisualisationVector" in ln:
            names.append(re.findall('"HDF">(.+?):/', ln)[0])
            index.append(int(re.findall("VisualisationVector/(.+?)</DataItem", ln)[0]))
    return names, times, index


def _ids(domains: np.ndarray, topology: np.ndarray, domain_id) -> np.ndarray:
    if isinstance(domain_id, (list, tuple)):
        sel = (domains == domain_id[0]) | (domains == domain_id[1])
    else:
        sel = domains == domain_id
    return np.unique(topology[sel])


def get_domain_ids(mesh_path: Union[str, Path], fluid_domain_id, solid_domain_id):
    """Node ids (rows of the whole-domain mesh's coordinates) of the fluid, of the solid and of everything."""
    mesh_path = Path(mesh_path)
    assert mesh_path.exists() and mesh_path.is_file(), f"Mesh file {mesh_path} does not exist"
    with H5File(mesh_path) as f:
        domains = f["domains/values"].read().ravel()
        topology = f["domains/topology"].read().astype(np.int64)
    return (_ids(domains, topology, fluid_domain_id), _ids(domains, topology, solid_domain_id), np.unique(topology))


def select_steps(timevalues: Sequence[float], save_time_step: float, stride: int = 1,
                 start_time: Optional[float] = None, end_time: Optional[float] = None) -> List[int]:
    """Positions in the XDMF step list that ``create_hdf5`` converts (``create_hdf5.py:118-133``)."""
    start_time = start_time if start_time is not None else timevalues[0]
    if end_time is not None:
        assert end_time > start_time, "end_time must be greater than start_time"
        assert end_time <= timevalues[-1], "end_time must be less than the last time step"
    end_time = end_time if end_time is not None else timevalues[-1]
    first = int(start_time / save_time_step) - 1
    last = int(end_time / save_time_step)
    n = len(timevalues)
    if first < -n or last > n:
        # the reference would raise IndexError on timevalue_list[file_counter]
        raise ValueError(f"start/end time select steps [{first}, {last}) outside the {n} saved steps "
                         f"(save_time_step = {save_time_step})")
    if first < 0:
        # int(start / save_time_step) - 1 < 0 happens when the first saved time is an accumulated float just below
        # save_time_step.  The reference then indexes its lists with a negative file_counter, i.e. from the END
        # (create_hdf5.py:128-131): same steps here, with a warning, so that both tools accept the same runs.
        logging.warning(f"WARNING : start index {first} is negative; like the reference, counting from the last step")
    return [i if i >= 0 else n + i for i in range(first, last, stride)]


class TurtleVelocitySeries:
    """Same face as :class:`vasp_b200.io_dolfin.VelocitySeries` (``names``, ``timestamps``, ``vec_len``,
    ``layout``, ``read_into``, ``close``), backed by the raw turtleFSI output."""

    def __init__(self, visualization_path: Union[str, Path], mesh_path: Union[str, Path], save_time_step: float,
                 stride: int = 1, start_time: Optional[float] = None, end_time: Optional[float] = None,
                 fluid_domain_id=1, solid_domain_id=2, compute_stride: Optional[int] = None,
                 derive_refined_mesh: bool = False):
        self.path = Path(visualization_path)
        xdmf = self.path / "velocity.xdmf"
        assert xdmf.exists(), f"Velocity file {xdmf} not found."
        files, times, index = output_file_lists(xdmf)
        if not (len(files) == len(times) == len(index)) or not files:
            raise ValueError(f"{xdmf}: {len(times)} time values, {len(files)} data items")
        steps = select_steps(times, save_time_step, stride, start_time, end_time)
        # ... written as vector_0, vector_1, ... and read back by get_dataset_names(step=stride)
        steps = steps[::(stride if compute_stride is None else compute_stride)]
        self._files = {}
        self.geometry: Optional[np.ndarray] = None
        if derive_refined_mesh:
            # no mesh_refined.h5 / mesh_refined_fluid.h5 needed.  turtleFSI stores the geometry its
            # arrays are indexed by next to them (/Mesh/0/mesh/geometry of the first file -- the authority
            # create_refined_mesh.py:64-83 itself renumbers the refined mesh against); the P2 nodes of the fluid mesh
            # are matched against those coordinates directly (K0), so neither the refinement (create_refined_mesh.py:
            # 50-151) nor the separation (separate_mesh.py:56-107) has to be run first.
            f0 = self._files[files[steps[0]]] = H5File(self.path / files[steps[0]])
            self.geometry = f0["Mesh/0/mesh/geometry"].read().astype(np.float64)
            if self.geometry.ndim != 2 or self.geometry.shape[1] != 3:
                raise ValueError(f"{self.path / files[steps[0]]}: /Mesh/0/mesh/geometry must be (N, 3)")
            self.fluid_ids = None
            self.n_all = len(self.geometry)
        else:
            self.fluid_ids, _, all_ids = get_domain_ids(mesh_path, fluid_domain_id, solid_domain_id)
            self.n_all = int(all_ids.max()) + 1 if len(all_ids) else 0
        self.names: List[str] = []
        self._fd: List[int] = []
        offsets = []
        for s in steps:
            fn = files[s]
            if fn not in self._files:
                self._files[fn] = H5File(self.path / fn)
            ds = self._files[fn][f"VisualisationVector/{index[s]}"]
            if ds.dtype != np.dtype("<f8") or len(ds.shape) != 2 or ds.shape[1] != 3:
                raise ValueError(f"{self.path / fn}: VisualisationVector/{index[s]} must be (N, 3) little-endian "
                                 f"float64, got {ds.shape} {ds.dtype}")
            if ds.shape[0] < self.n_all:
                raise ValueError(f"{self.path / fn}: VisualisationVector/{index[s]} has {ds.shape[0]} nodes, the "
                                 f"mesh numbers {self.n_all}")
            if ds.offset is None:
                raise ValueError(f"{self.path / fn}: VisualisationVector/{index[s]} is not stored contiguously")
            self.names.append(f"{fn}:/VisualisationVector/{index[s]}")
            self._fd.append(self._files[fn]._fh.fileno())
            offsets.append(int(ds.offset))
            self._rows = int(ds.shape[0])
        self.offsets = np.array(offsets, dtype=np.int64)
        self.timestamps = np.array([times[s] for s in steps], dtype=np.float64)
        self.vec_len = 3 * self._rows if steps else 0

    def __len__(self) -> int:
        return len(self.names)

    def layout(self, refined_tets: np.ndarray, n_nodes: int):
        """Component c of fluid-mesh vertex v sits at ``c + 3 * fluid_ids[v]`` of a raw array: the vertices of the
        separated fluid mesh are the whole-domain nodes ``unique(fluid topology)`` in ascending order
        (``separate_mesh.py:79-92``)."""
        if self.fluid_ids is None:  # derived route: the velocity nodes ARE the rows of the raw arrays
            if n_nodes != self.n_all:
                raise ValueError(f"{self.path}: the stored geometry has {self.n_all} nodes, asked for {n_nodes}")
            return (0, 1, 2), 3, None
        if len(self.fluid_ids) != n_nodes:
            raise ValueError(f"{self.path}: the domain table has {len(self.fluid_ids)} fluid nodes, the fluid mesh "
                             f"{n_nodes} vertices")
        return (0, 1, 2), 3, self.fluid_ids.astype(np.int64)

    def read_into(self, out: np.ndarray, first: int, last: int, pool=None) -> np.ndarray:
        """Same contract as :meth:`io_dolfin.VelocitySeries.read_into`: pieces of ~4 MiB, spread over ``pool``."""
        nbytes, chunk = self.vec_len * 8, 4 << 20
        jobs, cur, cur_bytes = [], [], 0
        for r, k in enumerate(range(first, last)):
            mv = memoryview(out[r, :self.vec_len]).cast("B")
            for lo in range(0, nbytes, chunk):
                hi = min(lo + chunk, nbytes)
                cur.append((mv[lo:hi], self._fd[k], int(self.offsets[k]) + lo))
                cur_bytes += hi - lo
                if cur_bytes >= chunk:
                    jobs.append(cur)
                    cur, cur_bytes = [], 0
        if cur:
            jobs.append(cur)

        def run(job) -> None:
            for mv, fd, off in job:
                got, n_all = 0, len(mv)
                while got < n_all:
                    n = os.preadv(fd, [mv[got:]], off + got)
                    if n <= 0:
                        raise IOError(f"{self.path}: short read at offset {off + got}")
                    got += n

        if pool is None or len(jobs) < 2:
            for j in jobs:
                run(j)
        else:
            list(pool.map(run, jobs))
        return out

    def row_addresses(self) -> np.ndarray:
        """Host address of every selected raw array inside the read-only ``mmap`` of its file (see
        :meth:`io_dolfin.VelocitySeries.row_addresses`)."""
        if not getattr(self, "_views", None):
            self._views = {fd: np.frombuffer(f._buf, dtype=np.uint8) for f in self._files.values()
                           for fd in [f._fh.fileno()]}
        base = np.array([self._views[fd].ctypes.data for fd in self._fd], dtype=np.uint64)
        return base + self.offsets.astype(np.uint64)

    def advise(self, first: int, last: int) -> None:
        pass  # several files: left to the kernel's read-ahead

    def close(self) -> None:
        self._views = None
        for f in self._files.values():
            f.close()
        self._files = {}
        self._fd = []
