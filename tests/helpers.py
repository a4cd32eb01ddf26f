"""Shared case builders for the parity tests (inputs only; the checker is ``oracle/``)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict

import numpy as np

from oracle import hemo_oracle as ho
from vasp_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"
FIELDS = ("TAWSS", "OSI", "RRT", "ECAP", "TWSSG")


def load_pipe() -> Dict:
    d = np.load(GOLDEN / "pipe_mesh.npz")
    return {"xyz": d["xyz"], "tets": d["tets"].astype(np.int64), "params": json.loads(str(d["params"]))}


def load_fluid(name: str) -> Dict:
    d = np.load(GOLDEN / "fluid_meshes.npz")
    return {"xyz": d[f"{name}_xyz"], "tets": d[f"{name}_tets"].astype(np.int64)}


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    ok = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isfinite(a), np.isfinite(b)), "non-finite entries differ"
    den = np.linalg.norm(b[ok])
    return float(np.linalg.norm(a[ok] - b[ok]) / den) if den > 0 else float(np.linalg.norm(a[ok]))


def make_case(xyz, tets, order: int, n_snap: int, seed: int = 5, shuffle_nodes: bool = True, period: float = 1.0):
    """Velocity series on the (shuffled) refined-mesh vertices (order 2) or the mesh vertices (order 1)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    if order == 2:
        pts, edges, new_id = synth.p2_points(xyz, tets, seed=seed if shuffle_nodes else None)
    else:
        pts = xyz
    basis = synth.velocity_basis(pts, seed=seed)
    t, coef = synth.velocity_coefficients(n_snap, period=period, seed=seed)
    u = synth.velocity_series(basis, coef)
    return {"xyz": xyz, "tets": np.asarray(tets, dtype=np.int64), "order": order, "points": pts, "u": u, "t": t,
            "dt": float(t[1] - t[0]) if n_snap > 1 else 1.0, "n_nodes": len(pts)}


def oracle_stress(case, mu: float) -> ho.SurfaceStress:
    if case["order"] == 2:
        cn, edges = ho.p2_cell_nodes(case["tets"])
        p2 = ho.p2_node_coordinates(case["xyz"], edges)
        tol = 1e-8 * float(np.max(case["points"].max(0) - case["points"].min(0)))
        node_of_p2 = ho.match_points(p2, case["points"], tol)
        return ho.SurfaceStress(case["xyz"], case["tets"], mu, 2, node_of_p2)
    return ho.SurfaceStress(case["xyz"], case["tets"], mu, 1)


def oracle_run(case, mu: float, keep_wss: bool = False):
    S = oracle_stress(case, mu)
    n = case["n_nodes"]
    res = ho.run_time_loop(S, case["u"], case["dt"], (0, n, 2 * n), keep_wss=keep_wss)
    fin = ho.finalize(res["wss_sum"], res["tawss_sum"], res["twssg_sum"], res["count"])
    return S, res, fin


def engine_for(case, mu: float, device: int = 0):
    from vasp_b200.engine import HemoEngine
    eng = HemoEngine(device)
    eng.set_mesh(case["xyz"], case["tets"])
    if case["order"] == 2:
        eng.set_velocity_layout(2, refined_xyz=case["points"])
    else:
        eng.set_velocity_layout(1)
    eng.begin(mu, case["dt"])
    return eng


def write_turtle_folder(tmp: Path, u_of_points, n_snap: int, dt: float, mu, save_step: int = 1, split_at=None,
                        seed: int = 31) -> Dict:
    """A results folder as turtleFSI + vasp-refine-mesh + vasp-separate-mesh leave it (no
    ``Visualization_separate_domain``): raw ``Visualization/velocity.xdmf`` + ``velocity*.h5`` with
    ``VisualisationVector/<i>`` of shape (N_all, 3) on the refined whole-domain mesh, whose fluid nodes are the
    vertices of ``mesh_refined_fluid.h5`` in ascending id order (``separate_mesh.py:79-92``).  Solid rows carry a
    poison value.  ``split_at``: step at which the output continues in a second h5 file (restarted run)."""
    from vasp_b200 import io_dolfin
    from vasp_b200.h5lite import H5Writer
    rng = np.random.default_rng(seed)
    src = load_pipe()
    xyz, tets = src["xyz"], src["tets"]
    for sub in ("Mesh", "Checkpoint", "Visualization"):
        (tmp / sub).mkdir()
    io_dolfin.write_mesh(tmp / "Mesh" / "mesh.h5", xyz, tets)
    io_dolfin.write_mesh(tmp / "Mesh" / "mesh_fluid.h5", xyz, tets)
    rx, rt = synth.refine_uniform(xyz, tets, seed=seed)
    io_dolfin.write_mesh(tmp / "Mesh" / "mesh_refined_fluid.h5", rx, rt)
    n_ref = len(rx)
    n_all = n_ref + 700
    fluid_ids = np.sort(rng.choice(n_all, size=n_ref, replace=False))
    solid_only = np.setdiff1d(np.arange(n_all), fluid_ids)
    coords = rng.normal(size=(n_all, 3)) * 10.0
    coords[fluid_ids] = rx
    solid_cells = np.stack([rng.choice(solid_only, 400), rng.choice(solid_only, 400), rng.choice(fluid_ids, 400),
                            rng.choice(solid_only, 400)], axis=1)
    topo = np.concatenate([fluid_ids[rt], solid_cells]).astype("<i8")
    domains = np.concatenate([np.full(len(rt), 1), np.full(len(solid_cells), 2)]).astype("<u8")
    order = rng.permutation(len(topo))  # cell order of the whole mesh is unrelated to the fluid mesh's
    with H5Writer(tmp / "Mesh" / "mesh_refined.h5") as w:
        w.create_dataset("/mesh/coordinates", coords.astype("<f8"))
        w.create_dataset("/mesh/topology", topo[order], attrs={"celltype": "tetrahedron"})
        w.create_dataset("/domains/topology", topo[order])
        w.create_dataset("/domains/values", domains[order])
    params = dict(src["params"], mu_f=mu, dt=dt, save_step=save_step, save_deg=2, dx_f_id=1, dx_s_id=2)
    (tmp / "Checkpoint" / "default_variables.json").write_text(json.dumps(params))
    times = [dt * save_step * (k + 1) for k in range(n_snap)]
    vecs = []
    files = ["velocity.h5"] + (["velocity_run_1.h5"] if split_at else [])
    writers = [H5Writer(tmp / "Visualization" / f) for f in files]
    # turtleFSI stores the mesh its arrays are indexed by once, with the first step (XDMFFile.write)
    writers[0].create_dataset("/Mesh/0/mesh/geometry", coords.astype("<f8"))
    writers[0].create_dataset("/Mesh/0/mesh/topology", topo[order], attrs={"celltype": "tetrahedron"})
    items = []
    for k, t in enumerate(times):
        uf = np.asarray(u_of_points(rx, t)).reshape(3, n_ref).T        # (n_ref, 3)
        raw = np.full((n_all, 3), 1.0e3)
        raw[fluid_ids] = uf
        which = 1 if (split_at and k >= split_at) else 0
        idx = k - split_at if which else k
        writers[which].create_dataset(f"/VisualisationVector/{idx}", raw.astype("<f8"))
        items.append((files[which], idx, t))
        vecs.append(uf.T.reshape(-1))                                  # what create_hdf5 would have written
    for w in writers:
        w.close()
    x = ['<?xml version="1.0"?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>',
         '<Xdmf Version="3.0" xmlns:xi="http://www.w3.org/2001/XInclude">', '  <Domain>',
         '    <Grid Name="TimeSeries_velocity" GridType="Collection" CollectionType="Temporal">']
    for k, (fn, idx, t) in enumerate(items):
        x += ['      <Grid Name="mesh" GridType="Uniform">']
        if k == 0:
            x += [f'        <Topology NumberOfElements="{len(topo)}" TopologyType="Tetrahedron" NodesPerElement="4">',
                  f'          <DataItem Dimensions="{len(topo)} 4" NumberType="UInt" Format="HDF">{fn}:/Mesh/0/mesh/'
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
    (tmp / "Visualization" / "velocity.xdmf").write_text("\n".join(x))
    return {"xyz": xyz, "tets": tets, "rx": rx, "rt": rt, "fluid_ids": fluid_ids, "n_all": n_all,
            "vecs": np.array(vecs), "times": times, "items": items}
