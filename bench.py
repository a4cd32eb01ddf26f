#!/usr/bin/env python
"""Benchmark of the wall-shear-stress hot path (BASELINE.json metric: wall-facet x snapshots / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--snapshots S]
                    [--wss none|steps|matrix] [--no-cpu-baseline] [--no-io-leg]

One *step* is one pass of the whole hot path over the workload's snapshot batch: staging (K1), per-snapshot traction +
time reductions (K2), fold of the segments + the final TAWSS/OSI/RRT/ECAP/TWSSG formulas (K3; the first fold of a time
loop overwrites the running sums), and -- when N > 1 -- the fused peer-memory reduction + final formulas over NVLink
(or ncclAllReduce + K4).  ``value`` times that with the snapshots already resident in HBM: CUDA events on the compute
stream around EXACTLY K steps, rotating over resident input copies larger than L2, no events between the kernels;
the same K steps then run once more with events around K1 and K2 for the per-launch durations of ``roofline``.
``e2e`` times the same pass through the public Python/C-ABI call with pinned HOST snapshots (H2D inside) and the D2H
read of the five result fields.  ``cpu_baseline`` is the restated reference algorithm (C oracle) on the host cores,
``io`` the HDF5 -> device path of the entry point timed apart (cold / warm page cache) plus the whole entry point.

Under ``torch.distributed.run`` (N > 1) every rank drives one GPU on its own contiguous time range (weak scaling:
each rank processes the workload's snapshot count) and rank 0 prints the single JSON line.  No torch is imported.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from vasp_b200 import synth  # noqa: E402
from vasp_b200.timeshard import env_rank_world  # noqa: E402

METRIC = "wall_facet_snapshots_per_s"
UNIT = "facet*snapshots/s"

# BASELINE.json configs[1..4]; sizes from SURVEY.md §8 (structured vessel stand-ins, no network for real meshes)
WORKLOADS = {
    # offset-stenosis tutorial size: ~14 k fluid tets, P1 velocity, 1000 snapshots
    "stenosis_p1": dict(n=8, m=36, stenosis=0.45, bulge=0.0, order=1, snapshots=1000,
                        desc="offset-stenosis tutorial-size cylinder (13.8k tets), P1 velocity, 1000 snapshots"),
    # same mesh with P2 velocity (the reference-native order, save_deg = 2): the small P2 profiling case
    "stenosis_p2": dict(n=8, m=36, stenosis=0.45, bulge=0.0, order=2, snapshots=1000,
                        desc="offset-stenosis tutorial-size cylinder (13.8k tets), P2 velocity, 1000 snapshots"),
    "aneurysm_p1": dict(n=40, m=208, stenosis=0.0, bulge=0.6, order=1, snapshots=2000,
                        desc="aneurysm-style fluid mesh (~2M tets), P1 velocity, 2000 snapshots"),
    "avf_p2": dict(n=52, m=308, stenosis=0.2, bulge=0.0, order=2, snapshots=4000,
                   desc="AVF-size mesh (~5M tets), P2 velocity, 4000 snapshots"),
    "vessel10m_p2": dict(n=64, m=407, stenosis=0.0, bulge=0.0, order=2, snapshots=2000,
                         desc="synthetic 10M-tet vessel, P2 velocity, 2000 snapshots"),
}
MU = 3.5e-3
PERIOD = 0.951  # s, one cardiac cycle of the offset-stenosis problem (simulations/offset_stenosis.py:38-41)


def algorithmic_bytes_per_unit(order: int, keep_wss: bool = False) -> int:
    """8 B x 3 components x (10 | 4) cell dofs, + 72 B if the per-step WSS is written."""
    return 8 * 3 * (10 if order == 2 else 4) + (72 if keep_wss else 0)


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """NVML samples of SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._period = period_s
        except Exception as e:  # NVML missing: report that, do not guess
            self._nv = None
            self.error = str(e)

    def _loop(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for k, b in names.items():
                    if bits & b:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self._period)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "error", "nvml")}
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def build_workload(name: str, n_snap: int, rank: int):
    w = WORKLOADS[name]
    mesh = synth.vessel_mesh(w["n"], w["m"], radius=2.0e-3, stenosis=w["stenosis"], bulge=w["bulge"], seed=1234)
    xyz, tets = mesh["xyz"], mesh["tets"]
    if w["order"] == 2:
        points, _, _ = synth.p2_points(xyz, tets, seed=1234)
    else:
        points = xyz
    basis = synth.velocity_basis(points, seed=2024)
    # weak scaling: rank r owns snapshots [r*n, (r+1)*n) of one long series, plus the halo snapshot before them
    halo = 1 if rank > 0 else 0
    dt = PERIOD / n_snap
    _, coef = synth.velocity_coefficients(n_snap + halo, period=PERIOD, seed=2024, t0=(rank * n_snap - halo) * dt)
    coef[:, 0] *= 0.3  # m/s scale
    return dict(xyz=xyz, tets=tets, points=points, basis=basis, coef=coef, dt=dt, halo=halo, order=w["order"],
                desc=w["desc"])


def run_reference(args, rank: int, world: int) -> None:
    """CPU arm: the restated reference algorithm (oracle/hemo_oracle.c, all host threads) on the same workload."""
    if rank != 0:
        return
    from oracle import c_oracle, hemo_oracle as ho
    wl = build_workload(args.workload, args.snapshots, 0)
    stress = ho.SurfaceStress(wl["xyz"], wl["tets"], MU, wl["order"],
                              None if wl["order"] == 1 else ho.match_points(
                                  ho.p2_node_coordinates(wl["xyz"], ho.p2_cell_nodes(wl["tets"])[1]), wl["points"],
                                  1e-8 * float(np.ptp(wl["points"], axis=0).max())))
    co = c_oracle.COracle(stress)
    n = len(wl["points"])
    threads = c_oracle.max_threads()
    # bounded sample: at most ~2 s of work per step
    n_s = min(args.snapshots, max(threads, int(2.0e6 * threads / max(stress.nF, 1))))
    u = synth.velocity_series(wl["basis"], wl["coef"][:n_s])
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = co.run(u, wl["dt"], (0, n, 2 * n), threads=threads)
        ho.finalize(res["wss_sum"], res["tawss_sum"], res["twssg_sum"], res["count"])
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    value = stress.nF * n_s / t
    sample = f"{n_s} of {args.snapshots} snapshots x {stress.nF} facets per step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "facets": stress.nF,
                       "snapshots_per_gpu": args.snapshots, "order": wl["order"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = FEniCS path restated in C (oracle/hemo_oracle.c); dolfin itself is not installable"}
    print(json.dumps(line), flush=True)


def run_ours(args, rank: int, local_rank: int, world: int) -> None:
    from vasp_b200.engine import HemoEngine, pinned_empty
    from vasp_b200.timeshard import NcclComm

    wl = build_workload(args.workload, args.snapshots, rank)
    n_snap, halo, order = args.snapshots, wl["halo"], wl["order"]
    eng = HemoEngine(local_rank)
    eng.set_mesh(wl["xyz"], wl["tets"])
    if order == 2:
        eng.set_velocity_layout(2, refined_xyz=wl["points"])
    else:
        eng.set_velocity_layout(1)
    nF, vec_len = eng.nF, eng.vec_len
    comm = NcclComm(eng, rank, world) if world > 1 else None

    # inputs: pinned host copy (e2e) and resident device copies (value).  Timing rule: inputs must not be served from
    # L2 across timed steps.  A workload smaller than 2x the 126 MB L2 gets enough rotating copies to exceed 2.5x L2
    # (consecutive steps read different copies, so the K steps run back to back with no flush kernel and no host
    # synchronisation in the timed region); larger workloads evict themselves.
    u_host = pinned_empty((n_snap + halo, vec_len))
    synth.velocity_series(wl["basis"], wl["coef"], out=u_host)
    L2 = 126e6
    n_copies = 1 if u_host.nbytes >= 2 * L2 else min(8, int(np.ceil(2.5 * L2 / u_host.nbytes)))
    d_copies = []
    for _ in range(n_copies):
        d = eng.device_alloc(u_host.nbytes)
        eng.h2d(d, u_host)
        d_copies.append(d)
    flags = 2 if halo else 1
    stride = vec_len * 8
    n_total = n_snap * world
    step_no = [0]

    # --wss steps|matrix: K2 also writes tau of every snapshot (72 B per facet and snapshot), as one dolfin vector
    # per snapshot (WSS.h5) or as rows of the time-major matrix of the spectral tools
    d_wss = eng.device_alloc(72 * nF * n_snap) if args.wss != "none" else 0

    def step_resident():
        d_u = d_copies[step_no[0] % n_copies]
        step_no[0] += 1
        eng.begin(MU, wl["dt"])
        if args.wss == "matrix":
            eng.set_wss_layout(n_snap, 0)
        eng.push_device(d_u, n_snap + halo, stride, flags, d_wss)
        if comm:
            comm.reduce_finalize(n_total, host=False)  # fused peer reduction + K4 (or ncclAllReduce + K4)
        else:
            eng.finalize_async(n_total)

    def step_e2e():
        eng.begin(MU, wl["dt"])
        eng.push(u_host, flags=flags)
        if comm:
            return comm.reduce_finalize(n_total)
        return eng.finalize(n_total)

    for _ in range(args.warmup):
        step_resident()
    eng.sync()
    with ClockSampler(local_rank) as clk:  # NVML starts before the barrier: its set-up time differs between ranks
        eng.sync()
        if comm:
            # Ranks leave a host barrier tens of us apart, which is not small against K steps of ~60 us.  One untimed
            # step after the barrier lines the GPUs up on the device (its cross-GPU reduction ends on every rank when
            # the last rank arrives); the start event is recorded on the stream right behind it.
            comm.barrier()
            step_resident()
        launches0 = eng.timers()["launches"]
        eng.timer_start()
        for _ in range(args.steps):
            step_resident()
        total_ms = eng.timer_stop()  # CUDA events on the compute stream; stop synchronises
        launches1 = eng.timers()["launches"]
        # Same K steps once more with CUDA events between the kernels (per-launch K1/K2 durations for the roofline).
        # The events sit between dependent launches, so this pass runs without programmatic dependent launch and is a
        # few us per step slower than the pass above; `value` comes from the pass above.
        eng.set_profile(True)
        eng.sync()
        if comm:
            comm.barrier()
            step_resident()
            eng.kernel_profile()  # drop the aligning step's events
        eng.timer_start()
        for _ in range(args.steps):
            step_resident()
        total_ms_events = eng.timer_stop()
    if comm:
        comm.barrier()
    k1_ms, k2_ms, k2_n = eng.kernel_profile()
    eng.set_profile(False)
    launches = launches1 - launches0  # k1 + k2 (+ k2 multi) + k3 + k4 per step (L2 flush not counted)
    if comm:
        total_ms = comm.max(total_ms)
    ms_per_step = total_ms / args.steps
    value = world * nF * n_snap / (ms_per_step * 1e-3)

    # end to end through the public API: pinned host snapshots in, five result fields out
    e2e_steps = max(1, min(args.steps, 10))
    if os.environ.get("VASP_B200_E2E_BATCH"):  # experiments: snapshots per host->device batch (default: auto)
        eng.set_tuning(batch_snapshots=int(os.environ["VASP_B200_E2E_BATCH"]))
    step_e2e()
    eng.sync()
    if comm:
        comm.barrier()
    e2e_h2d_ms = e2e_kernel_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = step_e2e()
        tm = eng.timers()  # per time loop (vh_begin resets them): CUDA-event sums over the batches of the push
        e2e_h2d_ms += tm["h2d_ms"] / e2e_steps        # copy stream
        e2e_kernel_ms += tm["kernel_ms"] / e2e_steps  # compute stream (overlaps the copies)
    eng.sync()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if comm:
        e2e_s = comm.max(e2e_s)
    e2e_value = world * nF * n_snap / e2e_s
    h2d_bytes = int((n_snap + halo) * vec_len * 8)
    d2h_bytes = int(5 * 3 * nF * 8)
    osi = out["OSI"]
    sane = bool(np.isfinite(out["TAWSS"]).all() and np.nanmin(osi) >= -1e-12 and np.nanmax(osi) <= 0.5 + 1e-12)

    reduction = ("none" if not comm else "fused peer-memory reduce+finalize (NVLink, CUDA IPC)" if comm.fused
                 else "ncclAllReduce + finalize")
    if comm:
        comm.close()  # collective teardown while every rank is alive (rank 0 still has the JSON line to assemble)
    for d in d_copies + ([d_wss] if d_wss else []):
        eng.device_free(d)
    if rank != 0:
        return
    peak, peak_src = measured_peak_gbs()
    b_alg = algorithmic_bytes_per_unit(order, args.wss != "none")
    units_per_launch = nF * n_snap
    k2_avg_ms = k2_ms / max(k2_n, 1)
    k1_avg_ms = k1_ms / max(k2_n, 1)
    units_per_launch = units_per_launch * args.steps // max(k2_n, 1)  # a step may split into several column blocks
    achieved = units_per_launch * b_alg / (k2_avg_ms * 1e-3) / 1e9
    # K1 (staging) moves 8 B x 3 components per wall-layer node and snapshot, read once and written once
    k1_bytes = 2 * 24 * eng.n_wall_nodes * (n_snap + halo) * args.steps / max(k2_n, 1)
    # ... but a gather fetches whole 32-byte DRAM sectors: count the distinct sectors the wall-layer nodes occupy in a
    # snapshot vector (the numbering of the synthetic meshes is randomly permuted on purpose, so a wall-layer node
    # rarely shares its sector with another one)
    wall_nodes = np.unique(eng.maps()["facet_nodes"])
    n_all = len(wl["points"])
    sectors = sum(np.unique((c * n_all + wall_nodes) // 4).size for c in range(3))
    k1_sector_bytes = (32 * sectors + 24 * eng.n_wall_nodes) * (n_snap + halo) * args.steps / max(k2_n, 1)
    traffic = None
    tfile = ROOT / "profiles" / "k2_traffic.json"  # written from an `ncu --set full` capture of this command
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl, n_snap)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "ms_per_step_with_kernel_events": total_ms_events / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "facets": nF, "snapshots_per_gpu": n_snap,
                   "order": order, "wss_output": args.wss, "velocity_nodes": int(len(wl["points"])), "wall_layer_nodes": eng.n_wall_nodes,
                   "tets": int(len(wl["tets"])),
                   "parallelism": f"time-shard x{world}",
                   "reduction": reduction, 
                   "l2": (f"{n_copies} rotating resident copies of the input ({n_copies * u_host.nbytes / 1e6:.0f} MB > "
                          f"2.5 x 126 MB L2), steps back to back" if n_copies > 1 else
                          f"input {u_host.nbytes / 1e6:.0f} MB per step, larger than 2 x 126 MB L2"),
                   "results_sane": sane},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps, "h2d_ms_per_step": e2e_h2d_ms,
                "h2d_gbs": h2d_bytes / (e2e_h2d_ms * 1e-3) / 1e9 if e2e_h2d_ms > 0 else None,
                "kernel_ms_per_step": e2e_kernel_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": f"k2_wall<{order}>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_unit": b_alg, "units_per_launch": units_per_launch,
                     "kernel_ms_per_launch": k2_avg_ms, "launches_timed": k2_n,
                     "stage_kernel": {"kernel": "k1_stage", "ms_per_launch": k1_avg_ms,
                                      "bytes_per_launch": k1_bytes,
                                      "achieved": k1_bytes / (k1_avg_ms * 1e-3) / 1e9 if k1_avg_ms > 0 else None,
                                      "frac": k1_bytes / (k1_avg_ms * 1e-3) / 1e9 / peak if k1_avg_ms > 0 else None,
                                      "wall_nodes": eng.n_wall_nodes,
                                      "sector_bytes_per_launch": k1_sector_bytes,
                                      "frac_sectors": (k1_sector_bytes / (k1_avg_ms * 1e-3) / 1e9 / peak
                                                       if k1_avg_ms > 0 else None)}},
        "cpu_baseline": cpu,
    }
    if world == 1 and not args.no_io_leg:
        eng.close()
        try:
            line["io"] = io_leg(wl, n_snap, u_host, local_rank)
        except Exception as e:  # the headline line must not depend on the scratch file system
            line["io"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)


def io_leg(wl, n_snap: int, u_host: np.ndarray, device: int):
    """HDF5 -> device, timed apart from the device-resident compute (north_star; SURVEY.md §8d "Timers").

    A `u.h5` in the layout create_hdf5.py:158-174 writes is produced from the same synthetic series (bounded to
    ~1 GiB), then read back through the path the drop-in CLI uses: raw `pread` into two pinned buffers one block
    ahead of the consumer, H2D on the copy stream, kernels on the compute stream.  Reported: open -> last result for a
    cold page cache (pages dropped with posix_fadvise) and a warm one, and -- third -- the whole entry point
    `compute_hemodyanamics()` including the six XDMF/HDF5 outputs (per-step WSS.h5 included)."""
    import os
    import shutil
    import tempfile
    from vasp_b200 import io_dolfin
    from vasp_b200.compute_hemodynamics import _BlockReader, compute_hemodyanamics, default_block_snapshots
    from vasp_b200.engine import HemoEngine
    from vasp_b200.h5lite import H5Writer
    import contextlib
    import io as _io

    order, vec_len = wl["order"], u_host.shape[1]
    n_io = int(max(3, min(n_snap, (1 << 30) // (vec_len * 8))))
    tmp = Path(tempfile.mkdtemp(prefix="vasp_b200_io_"))
    try:
        (tmp / "Mesh").mkdir()
        vsd = tmp / "Visualization_separate_domain"
        vsd.mkdir()
        io_dolfin.write_mesh(tmp / "Mesh" / "mesh.h5", wl["xyz"], wl["tets"])
        io_dolfin.write_mesh(tmp / "Mesh" / "mesh_fluid.h5", wl["xyz"], wl["tets"])
        if order == 2:  # only the coordinates of the refined mesh are used when u.h5 carries no dof tables
            io_dolfin.write_mesh(tmp / "Mesh" / "mesh_refined_fluid.h5", wl["points"], wl["tets"][:1])
        rows = u_host[wl["halo"]:wl["halo"] + n_io]
        with H5Writer(vsd / "u.h5") as w:
            for k in range(n_io):
                w.create_dataset(f"/velocity/vector_{k}", rows[k],
                                 attrs={"timestamp": float((k + 1) * wl["dt"]), "partition": np.array([0], dtype=np.uint64)})
        path = vsd / "u.h5"
        nbytes = n_io * vec_len * 8

        def drop_cache():
            fd = os.open(path, os.O_RDONLY)
            try:
                os.fsync(fd)
                os.posix_fadvise(fd, 0, 0, os.POSIX_FADV_DONTNEED)
            finally:
                os.close(fd)

        eng = HemoEngine(device)
        eng.set_mesh(wl["xyz"], wl["tets"])
        eng.set_velocity_layout(order, refined_xyz=wl["points"] if order == 2 else None)
        block = default_block_snapshots(vec_len)
        eng.set_tuning(batch_snapshots=block, chunk_snapshots=0)

        def run_once():
            t0 = time.perf_counter()
            series = io_dolfin.VelocitySeries(path, "velocity", 1)
            eng.begin(MU, float(series.timestamps[1] - series.timestamps[0]))
            reader = _BlockReader(series, 0, len(series), block)
            first = True
            for a, b, u in reader:
                eng.push(u, flags=1 if first else 0)
                first = False
            out = eng.finalize(len(series))
            dt_s = time.perf_counter() - t0
            tm = eng.timers()
            series.close()
            return dt_s, reader.io_seconds, tm, out

        res = {}
        for label in ("cold", "warm", "warm"):
            if label == "cold":
                drop_cache()
            dt_s, io_s, tm, out = run_once()
            res[label] = {"open_to_result_s": dt_s, "file_read_s": io_s, "h2d_s": tm["h2d_ms"] * 1e-3,
                          "kernels_s": tm["kernel_ms"] * 1e-3, "gbs": nbytes / dt_s / 1e9,
                          "value": eng.nF * n_io / dt_s}
        nF = eng.nF
        eng.close()
        runs = []
        for _ in range(2):
            t0 = time.perf_counter()
            log = _io.StringIO()
            with contextlib.redirect_stdout(log):
                compute_hemodyanamics(vsd, tmp / "Mesh" / "mesh.h5", MU, 1, velocity_degree=order, device=device)
            runs.append((time.perf_counter() - t0,
                         next((ln for ln in log.getvalue().splitlines() if ln.startswith("--- timing")), "")))
        cli_s, cli_line = min(runs)
        out_bytes = sum(f.stat().st_size for f in (tmp / "Hemodynamic_indices").iterdir())
        return {"file": "u.h5 (create_hdf5 layout), synthetic", "snapshots": n_io, "bytes": nbytes,
                "block_snapshots": block, "unit": UNIT, "hdf5_to_device": res,
                "page_cache": "cold = pages dropped with posix_fadvise(DONTNEED) before the run; warm = second of two "
                              "runs over the file just read",
                "entry_point": {"total_s": cli_s, "runs_s": [r[0] for r in runs], "breakdown": cli_line,
                                "value": nF * n_io / cli_s, "output_bytes": out_bytes,
                                "what": "compute_hemodyanamics(): mesh + u.h5 in, K0, time loop, WSS.h5 per step and "
                                        "the five index files out"},
                "scratch": str(tmp.parent)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline(wl, n_snap: int):
    """Oracle C port on the host cores (bounded sample of the same workload), rank 0 at N = 1 only."""
    from oracle import c_oracle, hemo_oracle as ho
    node_of_p2 = None
    if wl["order"] == 2:
        p2 = ho.p2_node_coordinates(wl["xyz"], ho.p2_cell_nodes(wl["tets"])[1])
        node_of_p2 = ho.match_points(p2, wl["points"], 1e-8 * float(np.ptp(wl["points"], axis=0).max()))
    stress = ho.SurfaceStress(wl["xyz"], wl["tets"], MU, wl["order"], node_of_p2)
    co = c_oracle.COracle(stress)
    threads = c_oracle.max_threads()
    n = len(wl["points"])
    n_s = min(n_snap, max(threads, int(4.0e6 * threads / max(stress.nF, 1))))
    u = synth.velocity_series(wl["basis"], wl["coef"][wl["halo"]:wl["halo"] + n_s])
    co.run(u[:max(1, n_s // 8)], wl["dt"], (0, n, 2 * n), threads=threads)  # warm-up
    reps, t_best = 0, float("inf")
    t_start = time.perf_counter()
    while reps < 5 and time.perf_counter() - t_start < 20.0:
        t0 = time.perf_counter()
        co.run(u, wl["dt"], (0, n, 2 * n), threads=threads)
        t_best = min(t_best, time.perf_counter() - t0)
        reps += 1
    return {"value": stress.nF * n_s / t_best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_s} of {n_snap} snapshots x {stress.nF} facets, best of {reps}",
            "note": "optimistic stand-in: the FEniCS original adds Python dof matching + global LU per snapshot"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="stenosis_p1")
    ap.add_argument("--snapshots", type=int, default=None, help="snapshots per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-io-leg", action="store_true", help="skip the HDF5 -> device and entry-point timings")
    ap.add_argument("--wss", choices=["none", "steps", "matrix"], default="none",
                    help="device-resident pass also writes the per-snapshot WSS (default: the metric's indices only)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rules: at least three warm-up steps
    if args.snapshots is None:
        args.snapshots = WORKLOADS[args.workload]["snapshots"]
    rank, local_rank, world = env_rank_world()
    if world == 1 and args.gpus > 1:
        # launched plainly with --gpus N: re-exec under the launcher the driver would use
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", str(port), str(Path(__file__).resolve()), *sys.argv[1:]])
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
