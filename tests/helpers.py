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
