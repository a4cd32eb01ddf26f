"""The HDF5 files this repository WRITES, held structurally to what legacy dolfin (HDF5 1.12, earliest format) writes.

No libhdf5 exists in the image, so the writer (``vasp_b200.h5lite.H5Writer``) cannot be checked by opening its files with
the real library.  Instead ``tests/h5struct.py`` -- a structural walker that shares no code with ``h5lite``'s reader --
fingerprints superblock, object headers, header messages (type, order, version, flags, encodings), group B-trees,
symbol-table nodes and local heaps, and this test requires the fingerprints of our files to equal those of the
dolfin-written files of the reference's own test data (``tests/golden/h5_structure.json``, made by
``tests/golden/make_h5_structure_golden.py``).  The XDMF text is held the same way to what the reference's own writer
of ``write_checkpoint`` series emits (``postprocessing_h5py_common.py:594-682``), whose member names are also the ones
its reader dereferences (``:234-242``).
"""
import json
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np
import pytest

from tests import h5struct as hs
from tests import helpers as H
from vasp_b200 import io_dolfin

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "h5_structure.json").read_text())


def _sig(obj):
    return json.loads(json.dumps(hs.object_signature(obj)))        # same key types as the stored golden


def _dataset_patterns():
    """Every (datatype, header) pattern a dolfin-written dataset shows in the reference's files, rank left out."""
    pats = []
    for f in GOLD["files"].values():
        for o in f["objects"].values():
            if o["kind"] == "dataset":
                pats.append(_strip_rank(o))
    return pats


def _strip_rank(o):
    o = json.loads(json.dumps(o))
    o["messages"]["1"].pop("rank", None)
    o.pop("attributes", None)
    return o


@pytest.mark.parametrize("name", ["pipe", "cylinder", "stenosis"])
def test_mesh_files_have_the_structure_dolfin_writes(tmp_path, name):
    src = H.load_pipe() if name == "pipe" else H.load_fluid(name)
    io_dolfin.write_mesh(tmp_path / "m.h5", src["xyz"], src["tets"])
    ours = hs.fingerprint(tmp_path / "m.h5")
    for ref_name, ref in GOLD["files"].items():
        assert ours["superblock"] == ref["superblock"], ref_name          # version 0, K values, sizes, flags, addresses
    serial = GOLD["files"]["cylinder/cylinder.h5"]["objects"]             # written by one dolfin process, like ours
    for path in ("/mesh", "/mesh/coordinates", "/mesh/topology", "/mesh/cell_indices"):
        want = json.loads(json.dumps(serial[path]))
        got = _sig(ours["objects"][path])
        if path == "/mesh/topology":                                     # same attributes, same encodings
            assert set(got["attributes"]) == set(want["attributes"]) == {"celltype", "partition"}
            want["attributes"]["celltype"]["datatype"]["size"] = got["attributes"]["celltype"]["datatype"]["size"]
        assert got == want, path
    root = _sig(ours["objects"]["/"])
    want = json.loads(json.dumps(serial["/"]))
    want["btree"]["snod_max_entries"] = root["btree"]["snod_max_entries"] = None   # the reference file has 3 members
    assert root == want


def test_checkpoint_files_have_dolfin_structure_and_the_reference_xdmf_shape(tmp_path):
    src = H.load_fluid("cylinder")
    from oracle import hemo_oracle as ho
    S = ho.SurfaceStress(src["xyz"], src["tets"], 1.0, 1)
    m = S.maps
    bgeom = src["xyz"][m.bvert_parent]
    n_steps = 300                                                        # > 2K * 2K' members: a two-level group B-tree
    w = io_dolfin.CheckpointWriter(tmp_path, "WSS", m.btopology, bgeom, True)
    rng = np.random.default_rng(0)
    for k in range(n_steps):
        w.write(rng.normal(size=(S.nF, 3, 3)), 0.5 + 0.25 * k)
    w.close()
    ws = io_dolfin.CheckpointWriter(tmp_path, "TAWSS", m.btopology, bgeom, False)
    ws.write(rng.normal(size=(S.nF, 3)), 0)
    ws.close()
    pats = _dataset_patterns()
    group_sig = json.loads(json.dumps(GOLD["files"]["cylinder/cylinder.h5"]["objects"]["/mesh"]))
    for name, steps in (("WSS", n_steps), ("TAWSS", 1)):
        fp = hs.fingerprint(tmp_path / f"{name}.h5")
        assert fp["superblock"] == GOLD["files"]["cylinder/cylinder.h5"]["superblock"]
        objs = fp["objects"]
        # the members the reference's reader dereferences (postprocessing_h5py_common.py:234-242,337-343)
        for k in (0, steps - 1):
            for member in ("vector", "cell_dofs", "x_cell_dofs", "cells", "mesh/topology", "mesh/geometry"):
                assert f"/{name}/{name}_{k}/{member}" in objs
        for path, o in objs.items():
            sig = _sig(o)
            if sig["kind"] == "dataset":
                assert _strip_rank(sig) in pats, path                    # a header dolfin writes for this datatype
            else:
                for key in ("kind", "header_version", "message_order", "messages", "heap"):
                    assert sig[key] == group_sig[key], (path, key)
                bt = sig["btree"]
                assert bt["snod_version"] == [1] and bt["snod_within_2k"] and bt["names_sorted"], path
        assert objs[f"/{name}"]["btree"]["depth"] == (2 if steps > 256 else 1)
        topo_attrs = _sig(objs[f"/{name}/{name}_0/mesh/topology"])["attributes"]
        assert list(topo_attrs) == ["celltype"] and topo_attrs["celltype"]["datatype"]["class"] == 3

    # ---- XDMF: element tree of the reference's own writer of checkpoint series, attribute names and fixed values --------
    def shape(elem):
        fixed = {k: v for k, v in elem.attrib.items()
                 if k in ("GridType", "CollectionType", "GeometryType", "ItemType", "ElementFamily", "ElementDegree",
                          "Center", "AttributeType", "NumberType", "Format", "NodesPerElement", "Version")}
        return (elem.tag, tuple(sorted(elem.attrib)), tuple(sorted(fixed.items())), tuple(shape(c) for c in elem))

    for name, att in (("WSS", "Vector"), ("TAWSS", "Scalar")):
        ours = ET.fromstring((tmp_path / f"{name}.xdmf").read_text().split("<!DOCTYPE")[0] +
                             (tmp_path / f"{name}.xdmf").read_text().split("[]>", 1)[-1])
        ref = ET.fromstring(GOLD[f"checkpoint_xdmf_{att}"])
        g_ours, g_ref = ours.find("Domain/Grid"), ref.find("Domain/Grid")
        assert (g_ours.tag, sorted(g_ours.attrib)) == (g_ref.tag, sorted(g_ref.attrib))
        step_ours, step_ref = g_ours.findall("Grid")[0], g_ref.findall("Grid")[0]
        so, sr = shape(step_ours), shape(step_ref)
        # the reference template is for tetrahedra (NodesPerElement 4); ours is the boundary triangle mesh
        assert json.dumps(so).replace('"3"', '"N"') == json.dumps(sr).replace('"4"', '"N"')
        for item_o, item_r in zip(step_ours.iter("DataItem"), step_ref.iter("DataItem")):
            # same members of the same group, in the same order
            assert item_o.text.split(":")[1].lstrip("/").split("/", 2)[2] == item_r.text.split(":")[1].split("/", 2)[2]
        assert step_ours.find("Attribute").attrib["ElementCell"] == "triangle"
