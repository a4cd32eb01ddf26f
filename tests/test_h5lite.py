"""The self-written HDF5 subset: reader against dolfin-written files, writer by round trip."""
from pathlib import Path

import numpy as np
import pytest

from vasp_b200.h5lite import H5File, H5FormatError, H5Writer

REF = Path("/root/reference/tests/test_data")
DOLFIN_FILES = {
    "hemodynamics_data/Mesh/mesh_fluid.h5": (2442, 11940),
    "cylinder/cylinder.h5": (352, 1647),
    "offset_stenosis/offset_stenosis.h5": (1287, 6590),
    "aneurysm/small_aneurysm.h5": (1224, 6073),
}


@pytest.mark.parametrize("rel", sorted(DOLFIN_FILES))
def test_reads_dolfin_written_meshes(rel):
    """Known shapes from the reference's own tests.  The files live in the read-only reference
    checkout, which exists in the build container only."""
    path = REF / rel
    if not path.exists():
        pytest.skip("reference checkout not present on this machine")
    nv, nc = DOLFIN_FILES[rel]
    with H5File(path) as f:
        assert sorted(f.keys()) == ["boundaries", "domains", "mesh"]
        xyz, topo = f["mesh/coordinates"], f["mesh/topology"]
        assert xyz.shape == (nv, 3) and xyz.dtype == np.dtype("<f8")
        assert topo.shape == (nc, 4) and topo.dtype == np.dtype("<i8")
        assert topo.attrs["celltype"].tobytes().rstrip(b"\0") == b"tetrahedron"
        t = topo.read()
        assert t.min() == 0 and t.max() == nv - 1
        assert f["domains/values"].shape == (nc,)
        # raw byte range really is the data (what the snapshot streamer relies on)
        raw = np.fromfile(path, dtype="<f8", count=3 * nv, offset=xyz.offset).reshape(nv, 3)
        assert np.array_equal(raw, xyz.read())


def test_golden_arrays_came_through_the_reader():
    g = np.load(Path(__file__).parent / "golden" / "pipe_mesh.npz")
    assert g["xyz"].shape == (2442, 3) and g["tets"].shape == (11940, 4)
    assert np.allclose(g["xyz"].min(axis=0), [0, -1, -1]) and np.allclose(g["xyz"].max(axis=0), [5, 1, 1])


def test_writer_round_trip_with_deep_btree(tmp_path):
    rng = np.random.default_rng(0)
    p = tmp_path / "t.h5"
    vals = {}
    with H5Writer(p) as w:
        w.create_dataset("/mesh/coordinates", rng.random((10, 3)))
        w.create_dataset("/mesh/topology", rng.integers(0, 10, (7, 4)),
                         attrs={"celltype": "tetrahedron", "partition": np.array([0], dtype=np.uint64)})
        for k in range(3000):  # > 8*32 entries: forces a three-level group B-tree
            vals[k] = rng.random(5)
            w.create_dataset(f"/velocity/vector_{k}", vals[k], attrs={"timestamp": 0.001 * k})
        w.create_group("/velocity", attrs={"count": np.uint64(3000)})
        w.create_dataset("/alias", None, alias_of="/mesh/coordinates")
        w.create_dataset("/empty", np.zeros((0, 3)))
        w.create_dataset("/ints32", np.arange(5, dtype=np.int32))
    with H5File(p) as f:
        assert f.leaf_k == 4 and f.internal_k == 16
        assert sorted(f.keys()) == ["alias", "empty", "ints32", "mesh", "velocity"]
        g = f["velocity"]
        assert len(g.keys()) == 3000 and int(g.attrs["count"]) == 3000
        for k in (0, 1, 999, 2999):
            d = g[f"vector_{k}"]
            assert np.array_equal(d.read(), vals[k]) and float(d.attrs["timestamp"]) == 0.001 * k
        assert f["mesh/topology"].attrs["celltype"].tobytes() == b"tetrahedron"
        assert np.array_equal(f["alias"].read(), f["mesh/coordinates"].read())
        assert f["alias"].offset == f["mesh/coordinates"].offset
        assert f["empty"].shape == (0, 3) and f["ints32"].dtype == np.dtype("<i4")
        assert "nope" not in f and "mesh/topology" in f
        with pytest.raises(KeyError):
            f["mesh/nope"]


def test_rejects_non_hdf5(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5" * 100)
    with pytest.raises(H5FormatError):
        H5File(p)


def test_dataset_table_fast_path_and_fallback(tmp_path):
    """``h5lite.dataset_table`` lifts offset and timestamp out of like object headers in one numpy pass; a header that
    differs anywhere else (here: an extra attribute, a continuation-free but longer header) takes the slow parse, and
    a vector of another length is refused."""
    from vasp_b200 import io_dolfin
    from vasp_b200.h5lite import H5File, H5Writer, dataset_table
    rng = np.random.default_rng(1)
    data = rng.normal(size=(9, 12))
    with H5Writer(tmp_path / "u.h5") as w:
        for k in range(9):
            attrs = {"timestamp": 0.25 * k, "partition": np.array([0], dtype=np.uint64)}
            if k == 4:
                attrs["extra"] = 7.0
            w.create_dataset(f"/velocity/vector_{k}", data[k], attrs=attrs)
    with H5File(tmp_path / "u.h5") as f:
        g = f["velocity"]
        names = [f"vector_{k}" for k in range(9)]
        off, ts, _ = dataset_table(g, names, "timestamp")
        assert np.array_equal(ts, 0.25 * np.arange(9))
        for k in range(9):
            assert off[k] == g[names[k]].offset
            assert np.array_equal(np.frombuffer(f._buf, "<f8", 12, int(off[k])), data[k])
    s = io_dolfin.VelocitySeries(tmp_path / "u.h5", "velocity", 2)
    assert list(s.timestamps) == [0.0, 0.5, 1.0, 1.5, 2.0] and s.vec_len == 12
    s.close()
    with H5Writer(tmp_path / "bad.h5") as w:
        w.create_dataset("/velocity/vector_0", data[0], attrs={"timestamp": 0.0})
        w.create_dataset("/velocity/vector_1", data[1, :7], attrs={"timestamp": 1.0})
    with pytest.raises(ValueError):
        io_dolfin.VelocitySeries(tmp_path / "bad.h5", "velocity", 1)
