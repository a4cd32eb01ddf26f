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
idg   ``InterpolateDG.__call__`` (compute_hemodynamics.py:65-89): the          ``oracle.hemo_oracle`` ``bcell_local`` (R5), which
      coordinate-matching dof copy from the cell to the boundary space        the CUDA precompute must reproduce bit-exactly
====  =====================================================================  ==========================================

``idg`` runs the method of the real class on an instance whose dolfin-typed attributes are replaced by small numpy-backed
objects (dof coordinates, cell dofs, facet -> cell, sub-space dofmaps): the loop, its ``np.allclose`` test and its
first-match ``break`` are the reference's; the dof numbering handed to it (cell-major DG1, 3 dofs per boundary cell) is
this repository's reading of dolfin's and does not influence *which* cell dof a boundary dof is matched to.

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
    for name in ("matplotlib", "matplotlib.pyplot", "mpi4py", "petsc4py", "vampy", "vampy.automatedPostprocessing",
                 "vampy.automatedPostprocessing.postprocessing_common"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["mpi4py"].MPI = sys.modules["petsc4py"].PETSc = None
    sys.modules["vampy.automatedPostprocessing.postprocessing_common"].get_dataset_names = None

    class _Params(dict):  # dolfin.parameters["form_compiler"]["quadrature_degree"] = ... at import time
        def __missing__(self, key):
            self[key] = _Params()
            return self[key]

    class _Dolfin(types.ModuleType):  # every dolfin name is importable; none of them is called by what runs here
        parameters = _Params()

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return None

    sys.modules["dolfin"] = _Dolfin("dolfin")
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


def run_reference_time_loop(rc) -> dict:
    """Execute ``compute_hemodyanamics`` of the reference (compute_hemodynamics.py:160-372) line by line.

    dolfin is absent, so the names the function uses are bound to small numpy-backed stand-ins: vectors with
    ``[:]`` / ``get_local`` / ``set_local`` / ``axpy`` / ``zero``, function spaces that only know their sizes, an
    ``HDF5File`` that reads the real ``u.h5`` through h5lite, an ``XDMFFile`` that records what is written.  The two
    finite-element pieces -- ``Stress`` (:120-157) and ``project_dg`` (postprocessing_fenics_common.py:31-54) -- are
    replaced by the oracle's restatements (they need an assembler).  Everything else is the reference's own code: which
    snapshots are read, ``dt``, ``tau_prev = 0`` at the first step, the magnitude via ``reshape(local_size, block_size)``
    and ``np.linalg.norm``, the accumulations, the division by ``counter``, RRT / OSI / ECAP, the OSI assertion, the
    order and time stamps of the ``WSS`` checkpoints.  Vector functions use dolfin's node-interleaved layout (dof
    ``3 n + c``), which is what the reference's own reshape assumes."""
    from oracle import hemo_oracle as ho
    from vasp_b200 import synth
    from vasp_b200.io_dolfin import get_dataset_names
    meshes = np.load(HERE / "fluid_meshes.npz")
    xyz, tets = meshes["cylinder_xyz"], meshes["cylinder_tets"].astype(np.int64)
    rx, rt = synth.refine_uniform(xyz, tets, seed=3)
    n = len(rx)
    n_snap, dt, mu, stride = 9, 0.02, 0.0035, 2
    basis = synth.velocity_basis(rx, seed=6)
    times = [dt * (k + 1) for k in range(n_snap)]
    coef = np.array([[1 + 0.6 * np.sin(40 * t), 0.2 * np.sin(70 * t + 1), 0.1 * np.cos(90 * t), 0.3 * np.sin(30 * t + 2)]
                     for t in times])
    vecs = synth.velocity_series(basis, coef)
    node_of_p2 = ho.match_points(ho.p2_node_coordinates(xyz, ho.p2_cell_nodes(tets)[1]), rx, 1e-9)
    S = ho.SurfaceStress(xyz, tets, mu, 2, node_of_p2)
    nF = S.nF
    rec = {"checkpoints": []}

    class Vec:
        def __init__(self, n_):
            self.a = np.zeros(n_)

        def __getitem__(self, idx):
            return self.a[idx].copy()

        def __setitem__(self, idx, val):
            self.a[idx] = val.a if isinstance(val, Vec) else val

        def get_local(self):
            return self.a.copy()

        def set_local(self, v):
            self.a[:] = v

        def apply(self, mode):
            pass

        def axpy(self, alpha, other):
            self.a += alpha * other.a

        def zero(self):
            self.a[:] = 0.0

    class Space:
        def __init__(self, kind, size, vector):
            self.kind, self.size, self.vector = kind, size, vector

        def num_sub_spaces(self):
            return 3 if self.vector else 0

        def dofmap(self):
            return types.SimpleNamespace(block_size=lambda: 3 if self.vector else 1)

        def sub(self, i):
            return types.SimpleNamespace(collapse=lambda: Space(self.kind + "_sub", self.size // 3, False), index=i,
                                         parent=self)

    class Fn:
        def __init__(self, space):
            self.space, self.v = space, Vec(space.size)

        def vector(self):
            return self.v

        def rename(self, a, b):
            self.name = a

        def sub(self, i):
            return ("component", self, i)

    class Assigner:
        def __init__(self, to_space, from_sub):
            pass

        def assign(self, target, source):
            _, f, i = source
            target.v.a[:] = f.v.a[i::3]

    class H5:
        def __init__(self, comm, path, mode):
            self.f = H5File(path)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self.f.close()

        def close(self):
            self.f.close()

        def read(self, target, name, *flags):
            if isinstance(target, Fn):
                target.v.a[:] = self.f[name.lstrip("/")].read().ravel()

        def attributes(self, name):
            return {k: (float(v) if np.ndim(v) == 0 else v) for k, v in self.f[name.lstrip("/")].attrs.items()}

    class Xdmf:
        Encoding = types.SimpleNamespace(HDF5="HDF5")

        def __init__(self, comm, path):
            self.parameters, self.name = {}, Path(path).stem

        def write_checkpoint(self, f, name, t, enc, append=False):
            rec["checkpoints"].append((self.name, name, float(t), bool(append), f.v.get_local()))

        def close(self):
            pass

    class StressStandIn:  # Stress(u=u_p2, ...): the oracle's restatement of :120-157, evaluated on u's current vector
        def __init__(self, u, V_dg, V_sub, mu_f, mesh, boundary_mesh):
            assert mu_f == mu
            self.u, self.out = u, Fn(V_sub)

        def __call__(self):
            self.out.v.a[:] = S(self.u.v.a, (0, n, 2 * n)).reshape(-1)
            return self.out

    class Pow:  # inner(twssg, twssg) ** (1 / 2)
        def __init__(self, f):
            self.f = f

        def __pow__(self, e):
            assert e == 0.5
            return self

    def project_dg_stand_in(expr, V):  # postprocessing_fenics_common.py:31-54 through the oracle's restatement
        g = Fn(V)
        g.v.a[:] = ho.project_dg_norm(expr.f.v.a.reshape(nF, 3, 3), S.area).reshape(-1)
        return g

    def names(f, step=1, vector_filename="/velocity/vector_%d"):  # VaMPy's get_dataset_names: not in the reference tree
        return ["/velocity/" + x for x in get_dataset_names(f.f["velocity"], step=step)]

    spaces = {"CG1": Space("CG1", 3 * n, True), "CG2": Space("CG2", 3 * n, True), "DGb3": Space("DGb3", 9 * nF, True),
              "DGb1": Space("DGb1", 3 * nF, False), "DG3": Space("DG3", 12 * len(tets), True)}

    def vfs(mesh, fam, deg):
        return spaces[{("refined", "CG", 1): "CG1", ("mesh", "CG", 2): "CG2", ("bmesh", "DG", 1): "DGb3",
                       ("mesh", "DG", 1): "DG3"}[(mesh.tag, fam, deg)]]

    mesh_objs = iter([types.SimpleNamespace(tag="mesh"), types.SimpleNamespace(tag="refined")])
    rc.Mesh = lambda: next(mesh_objs)
    rc.BoundaryMesh = lambda mesh, kind: types.SimpleNamespace(tag="bmesh")
    rc.VectorFunctionSpace = vfs
    rc.FunctionSpace = lambda mesh, fam, deg: spaces["DGb1"]
    rc.Function = Fn
    rc.HDF5File = H5
    rc.XDMFFile = Xdmf
    rc.MPI = types.SimpleNamespace(comm_world=None, rank=lambda comm: 0)
    rc.PETScDMCollection = types.SimpleNamespace(
        create_transfer_matrix=lambda a, b: types.SimpleNamespace(__mul__=None))

    class Transfer:  # R2: for nested meshes the transfer matrix is a permutation; the node map lives in the stand-in above
        def __mul__(self, vec):
            return vec.get_local()

    rc.PETScDMCollection = types.SimpleNamespace(create_transfer_matrix=lambda a, b: Transfer())
    rc.FunctionAssigner = Assigner
    rc.Stress = StressStandIn
    rc.project_dg = project_dg_stand_in
    rc.inner = lambda a, b: Pow(a)
    rc.get_dataset_names = names
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "Mesh").mkdir()
        (td / "Visualization_separate_domain").mkdir()
        for nm in ("mesh.h5", "mesh_fluid.h5"):
            io_dolfin.write_mesh(td / "Mesh" / nm, xyz, tets)
        io_dolfin.write_mesh(td / "Mesh" / "mesh_refined_fluid.h5", rx, rt)
        io_dolfin.write_velocity_series(td / "Visualization_separate_domain" / "u.h5", rt, n, vecs, times)
        with redirect_stdout(io.StringIO()) as log:
            rc.compute_hemodyanamics(td / "Visualization_separate_domain", td / "Mesh" / "mesh.h5", mu, stride)
        assert (td / "Hemodynamic_indices").is_dir()
    cps = rec["checkpoints"]
    wss = [c for c in cps if c[0] == "WSS"]
    res = {"loop_seed": np.array(json.dumps({"mesh": "cylinder", "refine_seed": 3, "basis_seed": 6, "n_snap": n_snap,
                                             "dt": dt, "mu": mu, "stride": stride})),
           "loop_coef": coef, "loop_stdout": np.array(log.getvalue()),
           "loop_wss_times": np.array([c[2] for c in wss]), "loop_wss_append": np.array([c[3] for c in wss]),
           "loop_wss": np.stack([c[4] for c in wss])}
    for c in cps:
        if c[0] != "WSS":
            assert c[0] == c[1] and c[2] == 0.0 and c[3] is False
            res["loop_" + c[0]] = c[4]
    assert sorted(k for k in res if k.startswith("loop_") and k[5:].isupper()) == \
        ["loop_ECAP", "loop_OSI", "loop_RRT", "loop_TAWSS", "loop_TWSSG"]
    return res


TURTLE_CASES = [(1, None, None), (2, None, None), (3, 0.01, 0.04), (1, 0.015, 0.03), (4, None, 0.045), (2, 0.035, 0.05)]


def write_raw_turtle_case(folder: Path):
    """A small results folder as turtleFSI + vasp-refine-mesh + vasp-separate-mesh leave it (raw velocity AND
    displacement series, restarted once), with exactly representable values; shared by the generator and the test."""
    from vasp_b200 import synth
    meshes = np.load(HERE / "fluid_meshes.npz")
    xyz, tets = meshes["cylinder_xyz"], meshes["cylinder_tets"].astype(np.int64)
    rx, rt = synth.refine_uniform(xyz, tets, seed=3)
    n_ref, n_all, n_steps, split_at, dt_save = len(rx), len(rx) + 90, 10, 6, 0.05
    fluid_ids = np.sort((np.arange(n_ref) * 1 + (np.arange(n_ref) * 90) // n_ref))        # strictly increasing, gaps
    solid_only = np.setdiff1d(np.arange(n_all), fluid_ids)
    coords = np.zeros((n_all, 3))
    coords[fluid_ids] = rx
    coords[solid_only] = 100.0 + np.arange(len(solid_only))[:, None]
    solid_cells = np.stack([solid_only[:60], solid_only[10:70], fluid_ids[:60], solid_only[20:80]], axis=1)
    topo = np.concatenate([fluid_ids[rt], solid_cells]).astype("<i8")
    domains = np.concatenate([np.full(len(rt), 1), np.full(len(solid_cells), 2)]).astype("<u8")
    for sub in ("Mesh", "Visualization"):
        (folder / sub).mkdir(parents=True)
    io_dolfin.write_mesh(folder / "Mesh" / "mesh_refined_fluid.h5", rx, rt)
    with H5Writer(folder / "Mesh" / "mesh_refined.h5") as w:
        w.create_dataset("/mesh/coordinates", coords.astype("<f8"))
        w.create_dataset("/mesh/topology", topo, attrs={"celltype": "tetrahedron"})
        w.create_dataset("/domains/topology", topo)
        w.create_dataset("/domains/values", domains)
    node, comp = np.arange(n_all)[:, None], np.arange(3)[None, :]
    for quantity in ("velocity", "displacement"):
        files = [H5Writer(folder / "Visualization" / f"{quantity}.h5"),
                 H5Writer(folder / "Visualization" / f"{quantity}_run_1.h5")]
        for k in range(n_steps):
            raw = ((node * 7 + comp * 3 + k * 11 + (5 if quantity == "displacement" else 0)) % 1009).astype(np.float64) \
                + 0.25 * comp
            which, idx = (0, k) if k < split_at else (1, k - split_at)
            files[which].create_dataset(f"/VisualisationVector/{idx}", raw)
        for f in files:
            f.close()
        txt = turtle_xdmf(n_steps, split_at, n_all, len(topo)).replace("velocity", quantity)
        (folder / "Visualization" / f"{quantity}.xdmf").write_text(txt)
    times = [0.001 * 5 * (k + 1) for k in range(n_steps)]
    return {"fluid_ids": fluid_ids, "n_all": n_all, "times": times, "save_time_step": 0.005, "n_ref": n_ref}


def run_reference_create_hdf5() -> dict:
    """``create_hdf5`` of the reference (create_hdf5.py:24-189) executed on a raw turtleFSI folder with emulated dolfin
    objects (``Function.vector().set_local``, an ``HDF5File`` that records every ``write(u, "/velocity", time)``):
    which steps are converted for a given (stride, start, end), from which h5 file and index, the fluid-node slice and
    the component-blocked flattening are the reference's."""
    import hashlib
    cr = importlib.import_module("vasp.postprocessing.postprocessing_fenics.create_hdf5")
    written = []

    class Vec:
        def set_local(self, v):
            self.a = np.array(v)

    class Fn:
        def __init__(self, space):
            self.v = Vec()

        def vector(self):
            return self.v

    class H5:
        def __init__(self, comm, path, mode=None, file_mode=None):
            self.path, self.mode = Path(path), mode or file_mode

        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def read(self, *a):
            pass

        def close(self):
            pass

        def write(self, obj, name, *t):
            if isinstance(obj, Fn):
                written.append((self.path.name, self.mode, name, float(t[0]), obj.v.a.copy()))

    cr.Mesh = lambda comm: object()
    cr.MPI = types.SimpleNamespace(comm_self=None)
    cr.HDF5File = H5
    cr.VectorFunctionSpace = lambda mesh, fam, deg: None
    cr.Function = Fn
    out = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        info = write_raw_turtle_case(td)
        (td / "Visualization_separate_domain").mkdir()
        for k, (stride, st, et) in enumerate(TURTLE_CASES):
            written.clear()
            with redirect_stdout(io.StringIO()):
                cr.create_hdf5(td / "Visualization", td / "Mesh" / "mesh_refined.h5", info["save_time_step"], stride,
                               st, et, False, 1, 2)
            u = [w for w in written if w[0] == "u.h5"]
            assert u and all(w[2] == "/velocity" for w in u) and u[0][1] == "w" and all(w[1] == "a" for w in u[1:])
            out[f"hdf5_{k}_times"] = np.array([w[3] for w in u])
            out[f"hdf5_{k}_sha"] = np.array([hashlib.sha256(np.ascontiguousarray(w[4]).tobytes()).hexdigest() for w in u])
            if k == 0:
                out["hdf5_0_first_vector"] = u[0][4]
    out["hdf5_cases"] = np.array(json.dumps(TURTLE_CASES))
    return out


def run_reference_stress_expression(rc) -> dict:
    """``Stress.__init__`` of the reference (compute_hemodynamics.py:142-150) evaluated numerically: ``grad``, ``sym``,
    ``FacetNormal`` and ``inner`` are bound to tiny numpy-backed tensor / vector / scalar classes that implement the
    operators the UFL expression uses (scalar * tensor, tensor * vector, unary minus, vector - vector, scalar * vector),
    so ``self.Ft`` comes out as a number for a given velocity gradient G and facet normal n: sign conventions, the
    factor 2 mu, sym() and the tangential projection are the reference's lines, not a restatement."""
    class Sc:
        def __init__(self, a):
            self.a = float(a)

        def __mul__(self, o):
            return Vc(self.a * o.a)

    class Vc:
        def __init__(self, a):
            self.a = np.asarray(a, dtype=np.float64)

        def __neg__(self):
            return Vc(-self.a)

        def __sub__(self, o):
            return Vc(self.a - o.a)

    class Tn:
        def __init__(self, a):
            self.a = np.asarray(a, dtype=np.float64)

        def __rmul__(self, s_):
            return Tn(s_ * self.a)

        def __mul__(self, v):
            return Vc(self.a @ v.a)

    state = {}
    rc.SurfaceProjector = lambda V: None
    rc.InterpolateDG = lambda *a: None
    rc.grad = lambda u: Tn(state["G"])
    rc.sym = lambda t: Tn(0.5 * (t.a + t.a.T))
    rc.FacetNormal = lambda mesh: Vc(state["n"])
    rc.inner = lambda a, b: Sc(a.a @ b.a)
    V = types.SimpleNamespace(ufl_element=lambda: types.SimpleNamespace(family=lambda: "Discontinuous Lagrange"))
    rng = np.random.default_rng(11)
    Gs, ns, mus, fts = [], [], [], []
    for _ in range(12):
        state["G"] = rng.normal(size=(3, 3))
        nrm = rng.normal(size=3)
        state["n"] = nrm / np.linalg.norm(nrm)
        mu_ = float(rng.uniform(1e-3, 2.0))
        st = rc.Stress(u=None, V_dg=V, V_sub=V, mu_f=mu_, mesh=None, boundary_mesh=None)
        Gs.append(state["G"]), ns.append(state["n"]), mus.append(mu_), fts.append(st.Ft.a)
    return {"stress_G": np.array(Gs), "stress_n": np.array(ns), "stress_mu": np.array(mus), "stress_Ft": np.array(fts)}


MAIN_SCENARIOS = [
    # (name, files to create under the folder, parameters or None / "broken", argv after --folder F)
    ("separate_domain_present", ["Visualization_separate_domain/", "Mesh/mesh.h5"],
     {"save_deg": 2, "dt": 0.001, "save_step": 5, "mu_f": 0.0035, "dx_f_id": 1, "dx_s_id": 2}, []),
    ("two_viscosities_stride_user_mesh", ["Visualization_separate_domain/", "elsewhere/my_mesh.h5"],
     {"save_deg": 2, "dt": 0.001, "save_step": 5, "mu_f": [0.0035, 0.005], "dx_f_id": 1, "dx_s_id": 2},
     ["--stride", "3", "--mesh-path", "<F>/elsewhere/my_mesh.h5"]),
    ("raw_output_refined_mesh", ["Mesh/mesh.h5", "Mesh/mesh_refined.h5"],
     {"save_deg": 2, "dt": 0.001, "save_step": 5, "mu_f": 0.0035, "dx_f_id": 1, "dx_s_id": [2, 1002]}, ["--stride", "2"]),
    ("raw_output_window_entire_domain", ["Mesh/mesh.h5", "Mesh/mesh_refined.h5", "elsewhere/m.h5"],
     {"save_deg": 2, "dt": 0.002, "save_step": 10, "mu_f": 0.0035, "dx_f_id": [1, 1001], "dx_s_id": 2},
     ["-st", "0.1", "-et", "0.5", "--extract-entire-domain", "--mesh-path", "<F>/elsewhere/m.h5"]),
    ("folder_missing", None, None, []),
    ("parameters_missing", ["Visualization_separate_domain/", "Mesh/mesh.h5"], None, []),
    ("parameters_broken", ["Visualization_separate_domain/", "Mesh/mesh.h5"], "broken", []),
    ("save_deg_1", ["Visualization_separate_domain/", "Mesh/mesh.h5"],
     {"save_deg": 1, "dt": 0.001, "save_step": 5, "mu_f": 0.0035, "dx_f_id": 1, "dx_s_id": 2}, []),
    ("raw_output_refined_mesh_missing", ["Mesh/mesh.h5"],
     {"save_deg": 2, "dt": 0.001, "save_step": 5, "mu_f": 0.0035, "dx_f_id": 1, "dx_s_id": 2}, []),
    ("default_mesh_missing", ["Visualization_separate_domain/"],
     {"save_deg": 2, "dt": 0.001, "save_step": 5, "mu_f": 0.0035, "dx_f_id": 1, "dx_s_id": 2}, []),
    ("user_mesh_missing", ["Visualization_separate_domain/", "Mesh/mesh.h5"],
     {"save_deg": 2, "dt": 0.001, "save_step": 5, "mu_f": 0.0035, "dx_f_id": 1, "dx_s_id": 2},
     ["--mesh-path", "<F>/nowhere.h5"]),
]


def make_scenario(base: Path, files, params) -> Path:
    """Folder of one MAIN_SCENARIOS entry (shared by the generator and the test)."""
    f = base / "case"
    if files is None:
        return f
    f.mkdir(parents=True)
    for rel in files:
        p = f / rel
        if rel.endswith("/"):
            p.mkdir(parents=True, exist_ok=True)
        else:
            p.parent.mkdir(parents=True, exist_ok=True)
            p.write_bytes(b"")
    if params is not None:
        (f / "Checkpoint").mkdir()
        (f / "Checkpoint" / "default_variables.json").write_text("{ not json" if params == "broken" else json.dumps(params))
    return f


def run_reference_main(rc) -> dict:
    """``main()`` of the reference (compute_hemodynamics.py:375-455) with ``create_hdf5`` and ``compute_hemodyanamics``
    replaced by recorders: which checks fire, which messages are printed, what the two stages are called with."""
    results = {}
    rc.MPI = types.SimpleNamespace(comm_world=types.SimpleNamespace(barrier=lambda: None), size=lambda c: 1,
                                   rank=lambda c: 0)
    for name, files, params, argv in MAIN_SCENARIOS:
        calls = []
        rc.create_hdf5 = lambda *a: calls.append(["create_hdf5"] + list(a))
        rc.compute_hemodyanamics = lambda *a: calls.append(["compute_hemodyanamics"] + list(a))
        with tempfile.TemporaryDirectory() as td:
            F = make_scenario(Path(td), files, params)
            old = sys.argv
            sys.argv = ["vasp-compute-hemo", "--folder", str(F)] + [a.replace("<F>", str(F)) for a in argv]
            log, err = io.StringIO(), None
            try:
                with redirect_stdout(log):
                    rc.main()
            except (AssertionError, RuntimeError) as e:
                err = [type(e).__name__, str(e).replace(str(F), "<F>")]
            finally:
                sys.argv = old

            def plain(v):
                return str(v).replace(str(F), "<F>") if isinstance(v, Path) else v
            results[name] = {"stdout": [ln for ln in log.getvalue().replace(str(F), "<F>").splitlines() if ln.strip()],
                             "error": err, "calls": [[plain(v) for v in c] for c in calls]}
    return {"main": np.array(json.dumps(results))}


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
    # ---- idg: the reference's InterpolateDG.__call__ on duck-typed spaces
    rc = importlib.import_module("vasp.postprocessing.postprocessing_fenics.compute_hemodynamics")
    from oracle import hemo_oracle as ho
    meshes = np.load(HERE / "fluid_meshes.npz")
    for name in ("cylinder", "stenosis"):
        xyz, tets = meshes[f"{name}_xyz"], meshes[f"{name}_tets"].astype(np.int64)
        S = ho.SurfaceStress(xyz, tets, 1.0, 1)
        m, cells = S.maps, S.tets                      # cells: vertices ascending, as dolfin orders them
        nF, nc = S.nF, len(cells)

        class _Sub:  # V.sub(k)
            def __init__(self, k):
                self.k = k

            def dofmap(self):
                return self

            def entity_closure_dofs(self, mesh, dim, entities):
                assert dim == 3 and len(entities) == 1
                return np.array([12 * entities[0] + 4 * self.k + v for v in range(4)])

        class _SubDofmap:  # dofmap of a collapsed scalar DG1 space on the boundary mesh
            def cell_dofs(self, i):
                return np.array([3 * i, 3 * i + 1, 3 * i + 2])

        class _Vec:
            def __init__(self, store, k):
                self.store, self.k = store, k

            def vector(self):
                return self

            def set_local(self, v):
                self.store[self.k] = np.array(v)

        got = {}
        idg = rc.InterpolateDG.__new__(rc.InterpolateDG)
        idg.V = types.SimpleNamespace(sub=lambda k: _Sub(k))
        idg.mesh = types.SimpleNamespace(topology=lambda: types.SimpleNamespace(dim=lambda: 3))
        idg.sub_map = np.arange(nF)                                        # boundary cell i <-> facet i
        idg.f_to_c = lambda facet: np.array([m.facet_cell[facet]])
        idg.sub_coords = [xyz[m.bcell_parent].reshape(-1, 3)] * 3          # dof 3 i + j sits on vertex j of cell i
        idg.w_sub_copy = [np.zeros(3 * nF) for _ in range(3)]
        idg.sub_dofmaps = [_SubDofmap()] * 3
        idg.dof_coords = np.repeat(xyz[cells][:, None, :, :], 3, axis=1).reshape(-1, 3)   # dof 12 c + 4 k + v
        idg.ws = [_Vec(got, k) for k in range(3)]
        idg.fa = types.SimpleNamespace(assign=lambda *a: None)
        idg.v_sub = "v_sub"
        u_vec = (np.arange(12 * nc, dtype=np.int64) * 2654435761 % 1000003).astype(np.float64)   # distinct, exact
        assert rc.InterpolateDG.__call__(idg, u_vec) == "v_sub"
        out[f"idg_{name}_boundary"] = np.stack([got[k] for k in range(3)])       # (3 components, 3 nF)
    # ---- stress: the UFL expression of Stress.__init__ evaluated with numeric operands (before rc.Stress is replaced)
    out.update(run_reference_stress_expression(rc))
    # ---- loop: the reference's compute_hemodyanamics() itself, lines 160-372, on emulated dolfin objects
    out.update(run_reference_time_loop(rc))
    # ---- main: the reference's main() with both stages recorded
    out.update(run_reference_main(rc))
    # ---- hdf5: the reference's create_hdf5() on a raw turtleFSI folder
    out.update(run_reference_create_hdf5())
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
