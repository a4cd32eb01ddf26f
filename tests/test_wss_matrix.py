"""CPU tests of the (dof x time) WSS matrix: the column selection is checked against a literal
restatement of the reference loop (postprocessing_h5py_common.py:284-371, quantity "wss"), the file route against
checkpoint files written by this repository's own writer."""
import numpy as np
import pytest

from vasp_b200 import io_dolfin, wss_matrix
from vasp_b200.io_turtle import output_file_lists


def reference_loop(vectors, time_ts, start_t, end_t, stride):
    """postprocessing_h5py_common.py:263-371 for quantity "wss", with the h5 reads replaced by ``vectors[i]``."""
    num_ts = int(len(time_ts))
    num_rows = vectors[0].shape[0]
    num_cols = num_ts - 1                                   # :263
    quantity_magnitude = np.zeros((num_rows, num_cols))     # :271
    idx_zeroed = 0
    start, stop = 0, num_ts - 1                             # :291-294
    for i in range(start, stop):                            # :298
        time_file = time_ts[i]
        if start_t <= time_file <= end_t and i % stride == 0:   # :311
            vector_array_full = vectors[i].reshape(-1, 1)
            quantity_magnitude[:, idx_zeroed] = vector_array_full[np.arange(num_rows), 0]   # :337-343
            idx_zeroed += 1
    return quantity_magnitude


@pytest.mark.parametrize("stride,window", [(1, (0.0, 10.0)), (2, (0.0, 10.0)), (3, (0.25, 0.75)), (1, (0.4, 0.4)),
                                           (1, (5.0, 6.0))])
def test_column_selection_matches_the_reference_loop(stride, window):
    rng = np.random.default_rng(3)
    times = [0.1 * (k + 1) for k in range(11)]
    vectors = [rng.normal(size=45) for _ in times]
    want = reference_loop(vectors, times, window[0], window[1], stride)
    cols = wss_matrix.select_columns(times, window[0], window[1], stride)
    got = np.zeros_like(want)
    if cols:
        got[:, :len(cols)] = np.stack([vectors[i] for i in cols], axis=1)
    assert np.array_equal(got, want)
    assert all(c < len(times) - 1 for c in cols)            # the last step never becomes a column


def _write_series(folder, n_steps, seed=0):
    rng = np.random.default_rng(seed)
    nF = 7
    btopo = rng.integers(0, 9, size=(nF, 3))
    bgeom = rng.normal(size=(9, 3))
    w = io_dolfin.CheckpointWriter(folder, "WSS", btopo, bgeom, True)
    times, vals = [], []
    for k in range(n_steps):
        v = rng.normal(size=(nF, 3, 3))
        t = 0.05 * (k + 1)
        w.write(v, t)
        times.append(t)
        vals.append(v.reshape(-1))
    w.close()
    return times, vals, btopo, bgeom


def test_file_route_reproduces_the_reference_semantics(tmp_path):
    times, vals, btopo, bgeom = _write_series(tmp_path, 9)
    names, ts, idx = output_file_lists(tmp_path / "WSS.xdmf")      # the reference's own XDMF scan finds every step
    assert names == ["WSS.h5"] * 9 and idx == list(range(9)) and np.allclose(ts, times)
    out = tmp_path / "npz"
    dt_files, dof_info, amp = wss_matrix.create_transformed_matrix_wss(tmp_path, out, 0.0, 1.0, 2)
    assert dt_files == ts[2] - ts[1] and amp is None
    assert set(dof_info) == {"cell_dofs", "cells", "mesh/geometry", "mesh/topology", "x_cell_dofs"}
    assert np.array_equal(dof_info["mesh/topology"], btopo) and np.array_equal(dof_info["mesh/geometry"], bgeom)
    m = np.load(out / "wss_mag.npz")["component"]
    assert m.shape == (63, 8)
    assert np.array_equal(m, reference_loop(vals, times, 0.0, 1.0, 2))
    # an existing file is replaced, not appended to
    wss_matrix.create_transformed_matrix_wss(tmp_path, out, 0.0, 1.0, 1)
    assert np.array_equal(np.load(out / "wss_mag.npz")["component"], reference_loop(vals, times, 0.0, 1.0, 1))


def test_fewer_than_three_steps_fail_like_the_reference(tmp_path):
    _write_series(tmp_path, 2)
    with pytest.raises(IndexError):                                 # time_ts[2] at :220
        wss_matrix.create_transformed_matrix_wss(tmp_path, tmp_path / "npz", 0.0, 1.0, 1)
