"""The dolfin-flavoured files around the hot path: mesh HDF5, ``u.h5`` time series, ``write_checkpoint`` outputs.

Formats are documented in SURVEY.md §5.9 from the reference's own writers and readers:

* mesh files        -- ``HDF5File.write(mesh, "/mesh")`` (``preprocessing/preprocessing_common.py:243-247``), read at
  ``compute_hemodynamics.py:187-197``;
* ``u.h5``          -- ``HDF5File.write(u, "/velocity", time)`` per step (``create_hdf5.py:158-174``), read through
  ``get_dataset_names`` / ``file_u.read`` / ``attributes(...)["timestamp"]`` at ``compute_hemodynamics.py:176-179,
  269,274,277``;
* ``<Name>.xdmf/.h5`` -- ``XDMFFile.write_checkpoint`` (``compute_hemodynamics.py:286,361``); the member names are the
  ones the reference's own consumer dereferences (``postprocessing_h5py_common.py:234-242,337-343``) and the XDMF
  text follows its template (``:639-662``).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import List, Sequence, Tuple, Union

import numpy as np

from .h5lite import H5File, H5FormatError, H5Writer, dataset_table


# ---------------------------------------------------------------------------------------------------------------
# meshes
# ---------------------------------------------------------------------------------------------------------------
def read_mesh(path: Union[str, Path], group: str = "mesh") -> Tuple[np.ndarray, np.ndarray]:
    """``(coordinates (nv,3) f8, topology (nc,4) i8)``; file row order is the cell numbering."""
    with H5File(path) as f:
        xyz = f[f"{group}/coordinates"].read().astype(np.float64, copy=False)   # read() already is a private copy
        tets = f[f"{group}/topology"].read().astype(np.int64, copy=False)
    if xyz.ndim != 2 or xyz.shape[1] != 3 or tets.ndim != 2 or tets.shape[1] != 4:
        raise ValueError(f"{path}: expected a tetrahedral mesh in 3-D")
    return xyz, tets


def write_mesh(path: Union[str, Path], xyz: np.ndarray, tets: np.ndarray, group: str = "mesh") -> None:
    """Serial dolfin layout: coordinates, topology (+celltype, partition), cell_indices."""
    with H5Writer(path) as w:
        w.create_dataset(f"/{group}/coordinates", np.asarray(xyz, dtype="<f8"))
        w.create_dataset(f"/{group}/topology", np.asarray(tets, dtype="<i8"),
                         attrs={"celltype": "tetrahedron", "partition": np.array([0], dtype=np.uint64)})
        w.create_dataset(f"/{group}/cell_indices", np.arange(len(tets), dtype="<i8"))


# ---------------------------------------------------------------------------------------------------------------
# u.h5
# ---------------------------------------------------------------------------------------------------------------
def get_dataset_names(group, step: int = 1, start: int = 0, num_files: int = 100000,
                      vector_filename: str = "vector_%d") -> List[str]:
    """VaMPy's ``get_dataset_names`` (called at ``compute_hemodynamics.py:179``; source not in the reference repo):
    advance ``start`` by ``step`` to the first existing dataset, then keep every ``start + i*step``,
    ``i < num_files``, that exists."""
    prefix, suffix = vector_filename.split("%d")
    present = set()
    for k in group.keys():
        if k.startswith(prefix) and k.endswith(suffix):
            core = k[len(prefix):len(k) - len(suffix)] if suffix else k[len(prefix):]
            if core.isdigit():
                present.add(int(core))
    if not present:
        raise ValueError("no velocity vectors in file")
    top = max(present)
    while start not in present:
        start += step
        if start > top:
            raise ValueError("no velocity vector reachable with this stride")
    last = min(top, start + (num_files - 1) * step)
    return [vector_filename % i for i in range(start, last + 1, step) if i in present]


class VelocitySeries:
    """``u.h5``: dataset table (name, timestamp, file offset) + the dof layout of the vectors."""

    def __init__(self, path: Union[str, Path], group: str = "velocity", stride: int = 1):
        self.path = Path(path)
        self._f = H5File(self.path)
        g = self._f[group]
        self.group = group
        self.names = get_dataset_names(g, step=stride)
        try:
            # one numpy pass over the object headers (they differ in data address and timestamp only)
            self.offsets, self.timestamps, self._ds = dataset_table(g, self.names, "timestamp")
        except H5FormatError as e:
            raise ValueError(f"{path}: velocity vectors differ: {e}") from e
        self.vec_len = int(np.prod(self._ds[0].shape))
        if self._ds[0].dtype != np.dtype("<f8"):
            raise ValueError(f"{path}: vectors must be little-endian float64, got {self._ds[0].dtype}")
        self._g = g

    def __len__(self) -> int:
        return len(self.names)

    def layout(self, refined_tets: np.ndarray, n_nodes: int):
        """``(comp_offset, node_stride, node_perm|None)`` such that component c of refined-mesh vertex v sits at
        ``comp_offset[c] + node_stride * node_perm[v]``.

        Derived from ``cell_dofs`` / ``x_cell_dofs`` / ``cells`` when the file has them (dolfin always writes
        them); otherwise the blocked layout of ``create_hdf5.py:158-163`` is assumed."""
        if self.vec_len != 3 * n_nodes:
            raise ValueError(f"{self.path}: vectors have {self.vec_len} entries, mesh needs 3*{n_nodes}")
        g = self._g
        if not all(k in g.keys() for k in ("cell_dofs", "x_cell_dofs", "cells")):
            return (0, n_nodes, 2 * n_nodes), 1, None
        cell_dofs = g["cell_dofs"].read().astype(np.int64).ravel()
        x = g["x_cell_dofs"].read().astype(np.int64).ravel()
        cells = g["cells"].read().astype(np.int64).ravel()
        tets = np.sort(np.asarray(refined_tets, dtype=np.int64), axis=1)
        if len(cells) != len(tets) or len(x) != len(tets) + 1 or np.any(np.diff(x) != 12):
            raise ValueError(f"{self.path}: cell_dofs tables do not describe a P1 vector field on the refined mesh")
        dofs = cell_dofs.reshape(-1, 3, 4)  # UFC: component-major, 4 vertices each
        src = np.full((3, n_nodes), -1, dtype=np.int64)
        verts = tets[cells]
        for c in range(3):
            src[c, verts] = dofs[:, c, :]
        if np.any(src < 0):
            raise ValueError(f"{self.path}: some refined-mesh vertices have no dof")
        ident = np.arange(n_nodes)
        if all(np.array_equal(src[c], c * n_nodes + ident) for c in range(3)):
            return (0, n_nodes, 2 * n_nodes), 1, None
        q = src[0] // 3
        if all(np.array_equal(src[c], 3 * q + c) for c in range(3)):
            return (0, 1, 2), 3, (None if np.array_equal(q, ident) else q)
        q = src[0]
        if all(np.array_equal(src[c], c * n_nodes + q) for c in range(3)):
            return (0, n_nodes, 2 * n_nodes), 1, q
        raise ValueError(f"{self.path}: unsupported dof numbering in cell_dofs")

    def read_chunks(self, out: np.ndarray, first: int, last: int, chunk_bytes: int = 4 << 20):
        """Jobs of about ``chunk_bytes`` -- each a list of ``(memoryview, file offset)`` pieces -- that together fill
        rows ``0 .. last-first`` of ``out`` with snapshots ``[first, last)``; independent of each other, so several
        threads can ``pread`` them (long vectors are cut, short ones grouped)."""
        nbytes = self.vec_len * 8
        jobs, cur, cur_bytes = [], [], 0
        for r, k in enumerate(range(first, last)):
            mv = memoryview(out[r, :self.vec_len]).cast("B")
            for lo in range(0, nbytes, chunk_bytes):
                hi = min(lo + chunk_bytes, nbytes)
                cur.append((mv[lo:hi], int(self.offsets[k]) + lo))
                cur_bytes += hi - lo
                if cur_bytes >= chunk_bytes:
                    jobs.append(cur)
                    cur, cur_bytes = [], 0
        if cur:
            jobs.append(cur)
        return jobs

    def pread_chunk(self, job) -> None:
        fd = self._f._fh.fileno()
        for mv, off in job:
            got, n_all = 0, len(mv)
            while got < n_all:
                n = os.preadv(fd, [mv[got:]], off + got)  # releases the GIL
                if n <= 0:
                    raise IOError(f"{self.path}: short read at offset {off + got}")
                got += n

    def read_into(self, out: np.ndarray, first: int, last: int, pool=None) -> np.ndarray:
        """Raw ``pread`` of snapshots ``[first, last)`` into the rows of ``out`` (no HDF5 library in the loop);
        ``pool`` (a ``ThreadPoolExecutor``) spreads the pieces over its threads."""
        jobs = self.read_chunks(out, first, last)
        if pool is None or len(jobs) < 2:
            for j in jobs:
                self.pread_chunk(j)
        else:
            list(pool.map(self.pread_chunk, jobs))
        return out

    def row_addresses(self) -> np.ndarray:
        """Host address of every selected vector inside the read-only ``mmap`` of the file (``uint64``): the
        wall-layer gather of the engine (``HemoEngine.compact_rows``) reads the page cache in place."""
        if getattr(self, "_base", None) is None:
            self._view = np.frombuffer(self._f._buf, dtype=np.uint8)
            self._base = int(self._view.ctypes.data)
        return (np.uint64(self._base) + self.offsets.astype(np.uint64)).astype(np.uint64)

    def advise(self, first: int, last: int) -> None:
        """Tell the kernel that snapshots ``[first, last)`` are about to be read through the mapping."""
        if last <= first:
            return
        try:
            import mmap as _mmap
            page = _mmap.PAGESIZE
            lo = int(self.offsets[first]) // page * page
            hi = int(self.offsets[last - 1]) + self.vec_len * 8
            self._f._buf.madvise(_mmap.MADV_WILLNEED, lo, hi - lo)
        except (AttributeError, OSError, ValueError):
            pass

    def close(self) -> None:
        self._ds = []
        self._g = None
        self._view = None
        self._base = None
        self._f.close()


def write_velocity_series(path: Union[str, Path], refined_tets: np.ndarray, n_nodes: int,
                          vectors: Sequence[np.ndarray], times: Sequence[float], group: str = "velocity") -> None:
    """Fixture writer in the layout ``create_hdf5.py:158-174`` produces (blocked vectors + dolfin's dof tables)."""
    tets = np.sort(np.asarray(refined_tets, dtype=np.int64), axis=1)
    nc = len(tets)
    cell_dofs = (np.arange(3)[None, :, None] * n_nodes + tets[:, None, :]).reshape(-1)
    with H5Writer(path) as w:
        for k, (v, t) in enumerate(zip(vectors, times)):
            v = np.asarray(v, dtype="<f8").ravel()
            if v.size != 3 * n_nodes:
                raise ValueError("vector length must be 3 * n_nodes")
            w.create_dataset(f"/{group}/vector_{k}", v,
                             attrs={"timestamp": float(t), "partition": np.array([0], dtype=np.uint64)})
        w.create_dataset(f"/{group}/cell_dofs", cell_dofs.astype("<i8"))
        w.create_dataset(f"/{group}/x_cell_dofs", (12 * np.arange(nc + 1)).astype("<i8"))
        w.create_dataset(f"/{group}/cells", np.arange(nc, dtype="<i8"))
        w.create_group(f"/{group}", attrs={"count": np.uint64(len(times))})


# ---------------------------------------------------------------------------------------------------------------
# write_checkpoint outputs
# ---------------------------------------------------------------------------------------------------------------
class CheckpointWriter:
    """``XDMFFile.write_checkpoint(f, name, t, HDF5, append)`` for a DG1 function on the boundary triangle mesh.

    Global dof numbering: scalar ``3*i + j``; vector ``3*(3*i + j) + c`` (facet i, boundary dof j, component c), so
    the ``vector`` dataset is node-interleaved like dolfin's (the reference reshapes it as ``(ndofs, 3)`` at
    ``compute_hemodynamics.py:289-296``); ``cell_dofs`` lists a cell's dofs component-major as UFC does.
    """

    def __init__(self, folder: Union[str, Path], name: str, btopology: np.ndarray, bgeometry: np.ndarray,
                 vector_valued: bool):
        self.folder, self.name = Path(folder), name
        self.nF, self.nBV = int(btopology.shape[0]), int(bgeometry.shape[0])
        self.vector_valued = vector_valued
        self.ncomp = 3 if vector_valued else 1
        self._btopo = np.asarray(btopology, dtype="<i8")
        self._bgeom = np.asarray(bgeometry, dtype="<f8")
        i = np.arange(self.nF)[:, None, None]
        c = np.arange(self.ncomp)[None, :, None]
        j = np.arange(3)[None, None, :]
        self._cell_dofs = (self.ncomp * (3 * i + j) + c).reshape(-1).astype("<i8")
        self._x = (3 * self.ncomp * np.arange(self.nF + 1)).astype("<i8")
        self._cells = np.arange(self.nF, dtype="<i8")
        for ext in (".h5", ".xdmf"):  # a stale series from a previous run would otherwise be appended to
            p = self.folder / f"{name}{ext}"
            if p.exists():
                p.unlink()
        self._w = H5Writer(self.folder / f"{name}.h5")
        self._times: List[float] = []

    def write(self, values: np.ndarray, time: float) -> None:
        vec = np.ascontiguousarray(values, dtype="<f8").reshape(1, -1)
        self.write_block(vec, [time])

    def write_block(self, values: np.ndarray, times: Sequence[float]) -> None:
        """``len(times)`` consecutive steps at once: ``values`` is ``(n, dofs...)``, e.g. the block of per-step WSS
        vectors a push returned -- its bytes go to the file in one write, each step's ``vector`` dataset points into
        them."""
        n = len(times)
        if n == 0:
            return
        vals = np.asarray(values)
        if vals.dtype != np.dtype("<f8") or not vals.flags.c_contiguous:
            vals = np.ascontiguousarray(vals, dtype="<f8")
        vals = vals.reshape(vals.shape[0], -1)[:n]
        ndofs = 3 * self.ncomp * self.nF
        if vals.shape != (n, ndofs):
            raise ValueError(f"{self.name}: expected {n} x {ndofs} dofs, got {vals.shape}")
        k0 = len(self._times)
        first = f"/{self.name}/{self.name}_0"
        if k0 == 0:
            self._w.create_dataset(f"{first}/mesh/topology", self._btopo, attrs={"celltype": "triangle"})
            self._w.create_dataset(f"{first}/mesh/geometry", self._bgeom)
            self._w.create_dataset(f"{first}/cell_dofs", self._cell_dofs)
            self._w.create_dataset(f"{first}/x_cell_dofs", self._x)
            self._w.create_dataset(f"{first}/cells", self._cells)
        for k in range(max(k0, 1), k0 + n):
            # dolfin re-writes the (identical) mesh and dof tables every step; here every later step links to the
            # objects of step 0 (HDF5 hard links: same paths, same content, written once)
            base = f"/{self.name}/{self.name}_{k}"
            for member in ("mesh", "cell_dofs", "x_cell_dofs", "cells"):
                self._w.link(f"{base}/{member}", f"{first}/{member}")
        self._w.create_dataset_block([f"/{self.name}/{self.name}_{k}/vector" for k in range(k0, k0 + n)], vals,
                                     shape=(ndofs, 1))
        self._times.extend(float(t) for t in times)

    def close(self) -> None:
        self._w.close()
        (self.folder / f"{self.name}.xdmf").write_text(self._xdmf())

    def _xdmf(self) -> str:
        n, nF, nBV = self.name, self.nF, self.nBV
        ndof = 3 * self.ncomp * nF
        att = "Vector" if self.vector_valued else "Scalar"
        out = ['<?xml version="1.0"?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>',
               '<Xdmf Version="3.0" xmlns:xi="http://www.w3.org/2001/XInclude">', '  <Domain>',
               f'    <Grid Name="{n}" GridType="Collection" CollectionType="Temporal">']
        for k, t in enumerate(self._times):
            p = f"{n}.h5:/{n}/{n}_{k}"
            out += [
                f'      <Grid Name="{n}_{k}" GridType="Uniform">',
                f'        <Topology NumberOfElements="{nF}" TopologyType="Triangle" NodesPerElement="3">',
                f'          <DataItem Dimensions="{nF} 3" NumberType="UInt" Format="HDF">{p}/mesh/topology</DataItem>',
                '        </Topology>',
                '        <Geometry GeometryType="XYZ">',
                f'          <DataItem Dimensions="{nBV} 3" Format="HDF">{p}/mesh/geometry</DataItem>',
                '        </Geometry>',
                f'        <Time Value="{t!r}" />',
                f'        <Attribute ItemType="FiniteElementFunction" ElementFamily="DG" ElementDegree="1" '
                f'ElementCell="triangle" Name="{n}" Center="Other" AttributeType="{att}">',
                f'          <DataItem Dimensions="{ndof} 1" NumberType="UInt" Format="HDF">{p}/cell_dofs</DataItem>',
                f'          <DataItem Dimensions="{ndof} 1" NumberType="Float" Format="HDF">{p}/vector</DataItem>',
                f'          <DataItem Dimensions="{nF + 1} 1" NumberType="UInt" Format="HDF">{p}/x_cell_dofs</DataItem>',
                f'          <DataItem Dimensions="{nF} 1" NumberType="UInt" Format="HDF">{p}/cells</DataItem>',
                '        </Attribute>',
                '      </Grid>',
            ]
        out += ['    </Grid>', '  </Domain>', '</Xdmf>', '']
        return "\n".join(out)


def read_checkpoint(folder: Union[str, Path], name: str, step: int = 0) -> dict:
    """Read one step back honouring ``cell_dofs`` (what dolfin's ``read_checkpoint`` does): values per
    (cell, local dof)."""
    with H5File(Path(folder) / f"{name}.h5") as f:
        base = f"{name}/{name}_{step}"
        vec = f[f"{base}/vector"].read().ravel()
        cell_dofs = f[f"{base}/cell_dofs"].read().ravel()
        x = f[f"{base}/x_cell_dofs"].read().ravel()
        topo = f[f"{base}/mesh/topology"].read()
        geom = f[f"{base}/mesh/geometry"].read()
    per_cell = int(x[1] - x[0])
    vals = vec[cell_dofs].reshape(-1, per_cell)
    return {"values": vals, "topology": topo, "geometry": geom}
