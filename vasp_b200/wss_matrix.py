"""The (dof x time) WSS matrix of VaSP's spectral post-processing.

The reference's spectral tools (``vasp-create-spectrograms-chromagrams --quantity wss``, ``spectrograms.py:309-311``)
do not consume ``WSS.h5`` step by step: ``create_transformed_matrix`` (``postprocessing_h5py_common.py:154-407``)
first re-reads every ``WSS/WSS_<i>/vector`` dataset and transposes the series into one ``(n_dofs, n_steps - 1)``
array, stored as ``wss_mag.npz`` under the key ``component`` (``:226-246,337-343,384-399``).  Row ``r`` is position
``r`` of the checkpoint vector (``9 f + 3 j + c``: facet, boundary dof, component -- the "mag" in the file name is
a misnomer, the whole vector is kept, ``:343``); column ``k`` is the ``k``-th *selected* step; unselected trailing
columns stay zero (``:263,337,369``).

Two producers of that file:

* :func:`create_transformed_matrix_wss` -- the reference's route, from an existing ``WSS.xdmf`` / ``WSS.h5`` pair
  (host only: it is pure file I/O).
* :func:`select_columns` + :meth:`HemoEngine.set_wss_matrix` -- the direct route: K2 writes tau time-major while it
  computes it (its lanes run along time, so a row segment is one 256-byte line), and the matrix never makes the
  round trip through ``WSS.h5``.  ``compute_hemodyanamics(..., wss_matrix_folder=...)`` uses it.

Both give bit-identical files (``tests/test_wss_matrix.py``, ``tests/test_gpu_wss_matrix.py``).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .h5lite import H5File
from .io_turtle import output_file_lists

DOF_INFO_NAMES = ("cell_dofs", "cells", "mesh/geometry", "mesh/topology", "x_cell_dofs")


def select_columns(times: Sequence[float], start_t: float, end_t: float, stride: int = 1) -> List[int]:
    """Steps of the series that become columns, in order (``postprocessing_h5py_common.py:292-294,311,337``):
    ``i in range(0, num_ts - 1)`` -- the last step is never used -- with ``start_t <= t_i <= end_t`` and
    ``i % stride == 0``."""
    n = len(times)
    return [i for i in range(0, n - 1) if start_t <= times[i] <= end_t and i % stride == 0]


def time_between_files(times: Sequence[float]) -> float:
    """``time_ts[2] - time_ts[1]`` (``:220``): the reference needs at least three steps."""
    if len(times) < 3:
        raise IndexError("list index out of range")  # what the reference raises
    return times[2] - times[1]


def write_npz(output_folder: Union[str, Path], matrix: np.ndarray, quantity: str = "wss") -> Path:
    """``<quantity>_mag.npz`` with the key ``component``; an existing file is replaced (``:384-399``)."""
    output_folder = Path(output_folder)
    output_folder.mkdir(parents=True, exist_ok=True)
    path = output_folder / f"{quantity}_mag.npz"
    if path.exists():
        path.unlink()
    np.savez_compressed(path, component=matrix)
    return path


def create_transformed_matrix_wss(input_path: Union[str, Path], output_folder: Union[str, Path], start_t: float,
                                  end_t: float, stride: int = 1
                                  ) -> Tuple[float, Dict[str, np.ndarray], Optional[Dict[str, np.ndarray]]]:
    """``create_transformed_matrix(..., quantity="wss", ...)`` from the checkpoint files in ``input_path``.

    Returns ``(time_between_files, dof_info_dict, dof_info_dict_amplitude)`` like the reference; the third item
    comes from ``MaxPrincipalStrain.xdmf`` (output of the solid post-processing, ``:257-266``) and is ``None`` when
    that file is absent (the reference raises there; nothing on this path needs it)."""
    input_path = Path(input_path)
    names, times, index = output_file_lists(input_path / "WSS.xdmf")
    dt_files = time_between_files(times)
    cols = select_columns(times, start_t, end_t, stride)
    current, f = None, None
    try:
        f = H5File(input_path / names[0])
        current = names[0]
        top = f.keys()[0]  # name_of_quantity_in_h5 (:232)
        first = f"{top}/{top}_0"
        dof_info = {n: np.array(f[f"{first}/{n}"].read()) for n in DOF_INFO_NAMES}
        n_rows = f[f"WSS/WSS_{index[0]}/vector"].read().shape[0]
        matrix = np.zeros((n_rows, len(times) - 1))
        for k, i in enumerate(cols):
            if names[i] != current:
                f.close()
                f = H5File(input_path / names[i])
                current = names[i]
            matrix[:, k] = f[f"WSS/WSS_{index[i]}/vector"].read()[:, 0]
    finally:
        if f is not None:
            f.close()
    amplitude = None
    mps = input_path / "MaxPrincipalStrain.xdmf"
    if mps.exists():
        a_names, _, _ = output_file_lists(mps)
        with H5File(input_path / a_names[0]) as g:
            top = "MaxPrincipalStrain"
            amplitude = {n: np.array(g[f"{top}/{top}_0/{n}"].read()) for n in DOF_INFO_NAMES}
    write_npz(output_folder, matrix)
    return dt_files, dof_info, amplitude


class DirectWssMatrix:
    """Host side of the direct route: owns the pinned ``(9 nF, n_steps - 1)`` matrix and tells the engine which
    pushed snapshots are columns.

    The engine writes tau of *every* pushed snapshot into consecutive columns (the time averages need them all);
    the reference keeps a subset (``select_columns``) packed to the left and never the last step, so
    :meth:`matrix` picks those columns out of the all-steps buffer at the end."""

    def __init__(self, engine, times: Sequence[float], start_t: float, end_t: float, stride: int = 1):
        from .engine import pinned_empty
        self.engine, self.times = engine, list(times)
        self.cols = select_columns(self.times, start_t, end_t, stride)
        self.dt_files = time_between_files(self.times)
        n = len(self.times)
        n_rows = 9 * engine.nF
        # the engine may be asked to push all n steps (the hot path needs them all for the time averages), so the
        # buffer it writes into has n columns either way; the reference's matrix is its first n - 1 columns
        self._all = pinned_empty((n_rows, n))
        self._all[:] = 0.0

    def attach(self, first_column: int = 0) -> None:
        self.engine.set_wss_matrix(self._all, first_column)

    def detach(self) -> None:
        self.engine.set_wss_matrix(None)

    def matrix(self) -> np.ndarray:
        n = len(self.times)
        out = np.zeros((self._all.shape[0], n - 1))
        if self.cols:
            out[:, :len(self.cols)] = self._all[:, self.cols]
        return out
