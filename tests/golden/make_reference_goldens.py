"""Golden vectors produced by RUNNING THE REFERENCE'S OWN PYTHON (build container only; nothing in the test-suite
reads ``/root/reference`` at run time).

    python tests/golden/make_reference_goldens.py        ->  tests/golden/reference_goldens.npz

The numerics of the hot path live in dolfin and cannot run here, but several functions around it are plain Python
and can: they are imported from ``/root/reference/src`` *unmodified*, with three absent third-party modules replaced
by shims in ``sys.modules`` -- ``h5py`` (a thin adapter over this repository's ``h5lite`` reader, only ``File``,
``[...]``, ``keys``, ``close``), ``matplotlib`` and ``dolfin`` (empty placeholders: none of their names is *called* by
the functions used here) -- and namespace stubs for the ``vasp`` packages themselves (their ``__init__`` files ask
``importlib.metadata`` for an installed distribution and import every sibling tool).  What is executed and recorded:

====  =====================================================================  ==========================================
key   reference function (file:line)                                          restated in this repository as
====  =====================================================================  ==========================================
ofl   ``output_file_lists`` (postprocessing_common.py:63-121)                 ``io_turtle.output_file_lists``
ids   ``get_domain_ids`` (postprocessing_common.py:16-60)                      ``io_turtle.get_domain_ids``
args  ``parse_arguments`` (postprocessing_fenics_common.py:10-28)              ``compute_hemodynamics.parse_arguments``
par   ``read_parameters_from_file`` (postprocessing_common.py:124-145)         ``compute_hemodynamics.read_parameters_...``
ctm   ``create_transformed_matrix(quantity="wss")``                            ``wss_matrix.create_transformed_matrix_wss``
      (postprocessing_h5py_common.py:154-407)
====  =====================================================================  ==========================================

Inputs are small synthetic files written with this repository's writers (stored in the fixture so the tests can
re-create them byte for byte) plus the reference's own test meshes (``tests/test_data/*``, domain tables only).
"""
import importlib
import io
import json
import sys
import tempfile
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from vasp_b200 import io_dolfin  # noqa: E402
from vasp_b200.h5lite import H5File, H5Writer  # noqa: E402

REF_SRC = Path("/root/reference/src")
REF_DATA = Path("/root/reference/tests/test_data")


def install_shims() -> None:
    h5py = types.ModuleType("h5py")

    class File(H5File):  # h5py.File(path, "r") -> read-only view; datasets answer [...] and np.array()
        def __init__(self, path, mode="r", *a, **k):
            assert mode == "r", "the shim is read-only"
            super().__init__(path)

    h5py.File = File
    sys.modules["h5py"] = h5py
    for name in ("matplotlib", "matplotlib.pyplot", "dolfin"):
        sys.modules[name] = types.ModuleType(name)
    for n in ("TestFunction", "TrialFunction", "inner", "Function", "LocalSolver", "dx", "FunctionSpace"):
        setattr(sys.modules["dolfin"], n, None)  # imported by name at postprocessing_fenics_common.py:7, never called
    # namespace stubs: the package __init__ files import every sibling tool (matplotlib, vmtk, ...)
    for dotted in ("vasp", "vasp.postprocessing", "vasp.postprocessing.postprocessing_fenics",
                   "vasp.postprocessing.postprocessing_h5py"):
        pkg = types.ModuleType(dotted)
        pkg.__path__ = [str(REF_SRC.joinpath(*dotted.split(".")))]
        sys.modules[dotted] = pkg
    sys.path.insert(0, str(REF_SRC))


WSS_CASES = [(0.0, 10.0, 1), (0.0, 10.0, 2), (0.12, 0.33, 1), (0.0, 10.0, 3), (5.0, 6.0, 1)]
ARGV_CASES = [
    ["--folder", "/data/case_1"],
    ["--folder", "rel/case", "--mesh-path", "/m/mesh.h5", "--stride", "4", "-st", "0.25", "-et", "1.5"],
    ["--folder", "x", "--start-time", "2", "--end-time", "3", "--extract-entire-domain", "--log-level", "10"],
    [],
]


def turtle_xdmf(n_steps: int, split_at: int, n_all: int = 40, n_cells: int = 90) -> str:
    """A ``Visualization/velocity.xdmf`` as turtleFSI writes it (first grid carries the mesh, the others include it),
    continued in a second h5 file after a restart; same template as ``tests/helpers.write_turtle_folder``."""
    x = ['<?xml version="1.0"?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>',
         '<Xdmf Version="3.0" xmlns:xi="http://www.w3.org/2001/XInclude">', '  <Domain>',
         '    <Grid Name="TimeSeries_velocity" GridType="Collection" CollectionType="Temporal">']
    for k in range(n_steps):
        fn = "velocity.h5" if k < split_at else "velocity_run_1.h5"
        idx = k if k < split_at else k - split_at
        t = 0.001 * 5 * (k + 1)
        x += ['      <Grid Name="mesh" GridType="Uniform">']
        if k == 0:
            x += [f'        <Topology NumberOfElements="{n_cells}" TopologyType="Tetrahedron" NodesPerElement="4">',
                  f'          <DataItem Dimensions="{n_cells} 4" NumberType="UInt" Format="HDF">{fn}:/Mesh/0/mesh/'
                  'topology</DataItem>', '        </Topology>', '        <Geometry GeometryType="XYZ">',
                  f'          <DataItem Dimensions="{n_all} 3" Format="HDF">{fn}:/Mesh/0/mesh/geometry</DataItem>',
                  '        </Geometry>']
        else:
            x += ['        <xi:include xpointer="xpointer(//Grid[@Name=&quot;TimeSeries_velocity&quot;]/Grid[1]/'
                  '*[self::Topology or self::Geometry])" />']
        x += [f'        <Time Value="{t!r}" />',
              '        <Attribute Name="velocity" AttributeType="Vector" Center="Node">',
              f'          <DataItem Dimensions="{n_all} 3" Format="HDF">{fn}:/VisualisationVector/{idx}</DataItem>',
              '        </Attribute>', '      </Grid>']
    x += ['    </Grid>', '  </Domain>', '</Xdmf>', '']
    return "\n".join(x)


def write_wss_case(folder: Path, vals: np.ndarray, times, btopo, bgeom) -> None:
    """WSS.xdmf/.h5 (vector DG1) and MaxPrincipalStrain.xdmf/.h5 (scalar DG1, needed by the reference at :257-266)."""
    w = io_dolfin.CheckpointWriter(folder, "WSS", btopo, bgeom, True)
    for v, t in zip(vals, times):
        w.write(v, t)
    w.close()
    m = io_dolfin.CheckpointWriter(folder, "MaxPrincipalStrain", btopo, bgeom, False)
    for v, t in zip(vals, times):
        m.write(np.linalg.norm(v, axis=2), t)
    m.close()


def main() -> None:
    install_shims()
    pc = importlib.import_module("vasp.postprocessing.postprocessing_common")
    fc = importlib.import_module("vasp.postprocessing.postprocessing_fenics.postprocessing_fenics_common")
    hc = importlib.import_module("vasp.postprocessing.postprocessing_h5py.postprocessing_h5py_common")
    out = {}
    rng = np.random.default_rng(20261017)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        # ---- ofl: output_file_lists on a turtleFSI series with a restart and on a write_checkpoint series
        txt = turtle_xdmf(9, 6)
        (td / "velocity.xdmf").write_text(txt)
        h5s, ts, idx = pc.output_file_lists(td / "velocity.xdmf")
        out["ofl_turtle_xdmf"] = np.array(txt)
        out["ofl_turtle"] = np.array(json.dumps([h5s, ts, idx]))
        # ---- ctm + ofl (checkpoint flavour)
        nF, n_steps = 6, 8
        btopo = rng.integers(0, 9, size=(nF, 3))
        bgeom = rng.normal(size=(9, 3))
        vals = rng.normal(size=(n_steps, nF, 3, 3))
        times = [0.05 * (k + 1) for k in range(n_steps)]
        case = td / "case" / "Hemodynamic_indices"
        case.mkdir(parents=True)
        write_wss_case(case, vals, times, btopo, bgeom)
        h5s, ts, idx = pc.output_file_lists(case / "WSS.xdmf")
        out["ofl_checkpoint"] = np.array(json.dumps([h5s, ts, idx]))
        out["ctm_btopo"], out["ctm_bgeom"], out["ctm_vals"], out["ctm_times"] = btopo, bgeom, vals, np.array(times)
        for k, (st, et, stride) in enumerate(WSS_CASES):
            dst = td / f"npz{k}"
            with redirect_stdout(io.StringIO()):
                dt_files, dof_info, dof_amp = hc.create_transformed_matrix(case, dst, case, "golden", st, et, "wss", 1, 2,
                                                                           stride)
            out[f"ctm{k}_matrix"] = np.load(dst / "wss_mag.npz")["component"]
            out[f"ctm{k}_dt"] = np.array(dt_files)
            assert sorted(p.name for p in dst.iterdir()) == ["wss_mag.npz"]
            if k == 0:
                for name, arr in dof_info.items():
                    out["ctm_dofinfo_" + name.replace("/", "__")] = arr
                for name, arr in dof_amp.items():
                    out["ctm_dofamp_" + name.replace("/", "__")] = arr
        out["ctm_cases"] = np.array(WSS_CASES, dtype=np.float64)
        # ---- par: read_parameters_from_file (present, absent, broken)
        (td / "p1" / "Checkpoint").mkdir(parents=True)
        params = {"dt": 0.001, "save_step": 5, "save_deg": 2, "mu_f": [0.0035, 0.004], "dx_f_id": 1, "dx_s_id": [2, 1002]}
        (td / "p1" / "Checkpoint" / "default_variables.json").write_text(json.dumps(params))
        (td / "p2" / "Checkpoint").mkdir(parents=True)
        (td / "p2" / "Checkpoint" / "default_variables.json").write_text("{ not json")
        out["par"] = np.array(json.dumps([pc.read_parameters_from_file(td / "p1"), pc.read_parameters_from_file(td / "p2"),
                                          pc.read_parameters_from_file(td / "p3")]))
        out["par_input"] = np.array(json.dumps(params))
    # ---- ids: get_domain_ids on the reference's own FSI test meshes
    for name, rel, fid, sid in (("cylinder", "cylinder/cylinder.h5", 1, 2),
                                ("stenosis", "offset_stenosis/offset_stenosis.h5", 1, 2),
                                ("aneurysm", "aneurysm/small_aneurysm.h5", 1, [2, 1002])):
        f, s, a = pc.get_domain_ids(REF_DATA / rel, fid, sid)
        out[f"ids_{name}_fluid"], out[f"ids_{name}_solid"], out[f"ids_{name}_all"] = f, s, a
        with H5File(REF_DATA / rel) as h:
            out[f"ids_{name}_domains"] = h["domains/values"].read().ravel().astype(np.int32)
            out[f"ids_{name}_topology"] = h["domains/topology"].read().astype(np.int32)
        out[f"ids_{name}_query"] = np.array(json.dumps([fid, sid]))
    # ---- args: the reference's argparse
    got = []
    for argv in ARGV_CASES:
        old = sys.argv
        sys.argv = ["vasp-compute-hemo"] + argv
        try:
            ns = fc.parse_arguments()
        finally:
            sys.argv = old
        got.append({k: (str(v) if isinstance(v, Path) else v) for k, v in vars(ns).items()})
    out["args"] = np.array(json.dumps({"argv": ARGV_CASES, "namespaces": got}))
    np.savez_compressed(HERE / "reference_goldens.npz", **out)
    print("wrote", HERE / "reference_goldens.npz", f"({(HERE / 'reference_goldens.npz').stat().st_size} bytes)")


if __name__ == "__main__":
    main()
