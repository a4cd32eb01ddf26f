"""Regenerate the committed fixtures from the reference's own test data (run in the build container only;
``/root/reference`` does not exist on the GPU box, so nothing in the test-suite reads it at run time).

    python tests/golden/make_golden.py

Writes, next to this script:

* ``pipe_mesh.npz``      -- ``tests/test_data/hemodynamics_data/Mesh/mesh_fluid.h5`` (the mesh of the reference's
  only hot-path test, ``tests/test_compute_hemodynamics.py``), coordinates float64, cells int32, plus the
  fixture's run parameters from ``Checkpoint/default_variables.json``.
* ``dolfin_facets.npz``  -- for the three serially written FSI meshes (cylinder, offset_stenosis, small_aneurysm):
  cells, domain ids and the exterior facets *in dolfin's own facet order*, taken from the ``/boundaries`` mesh
  function that dolfin wrote (facet index order); golden vector for R1.
* ``fluid_meshes.npz``   -- fluid sub-meshes (domain id 1) of those three meshes in ``separate_mesh.py:56-107``'s
  numbering (fluid nodes ascending parent id, cells in parent order), for parity cases beyond the pipe.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from vasp_b200.h5lite import H5File  # noqa: E402

REF = Path("/root/reference/tests/test_data")


def main() -> None:
    f = H5File(REF / "hemodynamics_data/Mesh/mesh_fluid.h5")
    params = json.loads((REF / "hemodynamics_data/Checkpoint/default_variables.json").read_text())
    keep = {k: params[k] for k in ("dt", "T", "save_step", "save_deg", "mu_f", "dx_f_id", "dx_s_id")}
    np.savez_compressed(HERE / "pipe_mesh.npz", xyz=f["mesh/coordinates"].read(),
                        tets=f["mesh/topology"].read().astype(np.int32), params=json.dumps(keep))
    facets, fluid = {}, {}
    for name, rel in (("cylinder", "cylinder/cylinder.h5"), ("stenosis", "offset_stenosis/offset_stenosis.h5"),
                      ("aneurysm", "aneurysm/small_aneurysm.h5")):
        f = H5File(REF / rel)
        tets = f["mesh/topology"].read()
        xyz = f["mesh/coordinates"].read()
        dom = f["domains/values"].read()
        btopo = f["boundaries/topology"].read()  # every facet, dolfin facet index order (serial write)
        # exterior facets = those whose sorted triple occurs in exactly one cell
        keepf = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
        faces = np.sort(np.sort(tets, axis=1)[:, keepf].reshape(-1, 3), axis=1)
        nv = len(xyz)
        key = (faces[:, 0] * nv + faces[:, 1]) * nv + faces[:, 2]
        uk, cnt = np.unique(key, return_counts=True)
        ext_keys = set(uk[cnt == 1].tolist())
        bs = np.sort(btopo, axis=1)
        bkey = (bs[:, 0] * nv + bs[:, 1]) * nv + bs[:, 2]
        is_ext = np.array([k in ext_keys for k in bkey.tolist()])
        facets[f"{name}_tets"] = tets.astype(np.int32)
        facets[f"{name}_exterior_in_dolfin_order"] = btopo[is_ext].astype(np.int32)
        # fluid sub-mesh, numbering of separate_mesh.py:65-107
        ft = tets[dom == 1]
        ids = np.unique(ft)
        remap = np.full(nv, -1, dtype=np.int64)
        remap[ids] = np.arange(len(ids))
        fluid[f"{name}_xyz"] = xyz[ids]
        fluid[f"{name}_tets"] = remap[ft].astype(np.int32)
    np.savez_compressed(HERE / "dolfin_facets.npz", **facets)
    np.savez_compressed(HERE / "fluid_meshes.npz", **fluid)
    for p in sorted(HERE.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
