"""Host logic of the raw turtleFSI reader: XDMF step table, domain ids, the step selection of
``create_hdf5`` pinned on the reference's own test cases, and the in-place layout handed to the engine."""
import numpy as np
import pytest

from tests import helpers as H
from vasp_b200 import io_turtle


def _u(p, t):
    return np.concatenate([(1 + t) * p[:, 0], t * p[:, 1] - 2.0, np.sin(p[:, 2]) + t])


def test_step_selection_matches_the_reference_tests():
    """tests/test_create_hdf5_and_separate_viz.py: turtleFSI -dt 0.001 -T 0.002 saves three steps; the plain run
    converts vector_0..2 (:41-44), --stride 2 keeps the first and the last (:143-146), --start-time 0.001
    --end-time 0.002 keeps the first two (:199-202)."""
    times = [0.001, 0.002, 0.003]
    assert io_turtle.select_steps(times, 0.001) == [0, 1, 2]
    assert io_turtle.select_steps(times, 0.001, stride=2) == [0, 2]
    assert io_turtle.select_steps(times, 0.001, start_time=0.001, end_time=0.002) == [0, 1]
    with pytest.raises(AssertionError):          # create_hdf5.py:122-123
        io_turtle.select_steps(times, 0.001, start_time=0.002, end_time=0.001)
    with pytest.raises(AssertionError):
        io_turtle.select_steps(times, 0.001, end_time=0.004)
    # first saved time just below save_time_step: int(t0 / dt) - 1 = -1, and the reference's range(-1, last) indexes
    # its lists from the END for that first step (create_hdf5.py:128-131); same selection here
    assert io_turtle.select_steps([0.0, 0.001], 0.001) == [1, 0]
    assert io_turtle.select_steps([0.00099999, 0.002, 0.003], 0.001) == [2, 0, 1, 2]
    with pytest.raises(ValueError):              # IndexError in the reference
        io_turtle.select_steps([0.001, 0.002], 0.001, end_time=0.002, start_time=-0.005)


def test_raw_series_is_the_slice_create_hdf5_would_write(tmp_path):
    info = H.write_turtle_folder(tmp_path, _u, n_snap=7, dt=0.05, mu=1.0, split_at=4)
    files, times, index = io_turtle.output_file_lists(tmp_path / "Visualization" / "velocity.xdmf")
    assert files == [it[0] for it in info["items"]] and index == [it[1] for it in info["items"]]
    assert np.allclose(times, info["times"], rtol=0, atol=0)
    fluid, solid, allids = io_turtle.get_domain_ids(tmp_path / "Mesh" / "mesh_refined.h5", 1, 2)
    assert np.array_equal(fluid, info["fluid_ids"])
    assert len(np.intersect1d(fluid, solid)) > 0 and len(allids) <= info["n_all"]
    s = io_turtle.TurtleVelocitySeries(tmp_path / "Visualization", tmp_path / "Mesh" / "mesh_refined.h5", 0.05,
                                       fluid_domain_id=1, solid_domain_id=2)
    assert len(s) == 7 and np.allclose(s.timestamps, info["times"])
    comp_offset, node_stride, perm = s.layout(info["rt"], len(info["rx"]))
    assert comp_offset == (0, 1, 2) and node_stride == 3 and np.array_equal(perm, info["fluid_ids"])
    buf = np.empty((7, s.vec_len))
    s.read_into(buf, 0, 7)
    n = len(info["rx"])
    for k in range(7):
        got = np.concatenate([buf[k, c + 3 * perm] for c in range(3)])   # what K1 gathers
        assert np.array_equal(got, info["vecs"][k])                       # = u.h5 vector of create_hdf5.py:158-163
        assert got.shape == (3 * n,)
    s.close()
    # main() hands --stride to create_hdf5 AND to compute_hemodyanamics (:431,455): applied twice
    s = io_turtle.TurtleVelocitySeries(tmp_path / "Visualization", tmp_path / "Mesh" / "mesh_refined.h5", 0.05,
                                       stride=2)
    assert np.allclose(s.timestamps, np.array(info["times"])[[0, 4]])
    s.close()
    with pytest.raises(ValueError):
        s2 = io_turtle.TurtleVelocitySeries(tmp_path / "Visualization", tmp_path / "Mesh" / "mesh_refined.h5", 0.05)
        s2.layout(info["rt"], len(info["rx"]) + 1)
