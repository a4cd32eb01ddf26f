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


def build_geometry(name: str):
    """Mesh, velocity nodes and the spatial modes of the synthetic series."""
    w = WORKLOADS[name]
    mesh = synth.vessel_mesh(w["n"], w["m"], radius=2.0e-3, stenosis=w["stenosis"], bulge=w["bulge"], seed=1234)
    xyz, tets = mesh["xyz"], mesh["tets"]
    new_id = None
    if w["order"] == 2:
        points, _, new_id = synth.p2_points(xyz, tets, seed=1234)
    else:
        points = xyz
    basis = synth.velocity_basis(points, seed=2024)
    return dict(name=name, xyz=xyz, tets=tets, points=points, basis=basis, order=w["order"], desc=w["desc"],
                new_id=new_id)


def series_coefficients(n_snap: int, first: int, period_snaps: int):
    """Mode coefficients of snapshots [first, first + n_snap) of one long series whose period is ``period_snaps``."""
    dt = PERIOD / period_snaps  # same waveform as synth.velocity_coefficients, on the time grid all ranks share
    t = dt * (first + np.arange(1, n_snap + 1))
    rng = np.random.default_rng(2024 + 1)
    phase = rng.uniform(0, 2 * np.pi, size=synth.N_MODES)
    coef = np.empty((n_snap, synth.N_MODES))
    coef[:, 0] = 1.0 + 0.6 * np.sin(2 * np.pi * t / PERIOD) + 0.3 * np.sin(4 * np.pi * t / PERIOD)
    for k in range(1, synth.N_MODES):
        coef[:, k] = 0.2 * np.sin(2 * np.pi * k * t / PERIOD + phase[k])
    coef[:, 0] *= 0.3  # m/s scale
    return coef, dt


def build_workload(name: str, n_snap: int, rank: int):
    """Geometry + the coefficients of rank ``rank``'s snapshots under weak scaling: rank r owns snapshots
    [r*n, (r+1)*n) of one long series, plus the halo snapshot before them."""
    g = build_geometry(name)
    halo = 1 if rank > 0 else 0
    coef, dt = series_coefficients(n_snap + halo, rank * n_snap - halo, n_snap)
    return dict(g, coef=coef, dt=dt, halo=halo)


def oracle_for(wl):
    """The restated reference on this workload (test infrastructure; only the CPU legs of the bench come here)."""
    from oracle import c_oracle, hemo_oracle as ho
    # synth.p2_points numbers vertices, then the midpoints of the edges in lexicographic order -- as the oracle does --
    # and new_id is where each of them went in the shuffled point set (checked on a sample below; a cKDTree match of
    # 13.6 M points would take longer than everything else the bench does)
    node_of_p2 = wl["new_id"] if wl["order"] == 2 else None
    stress = ho.SurfaceStress(wl["xyz"], wl["tets"], MU, wl["order"], node_of_p2)
    if node_of_p2 is not None:
        p2 = ho.p2_node_coordinates(wl["xyz"], stress.edges)
        pick = np.random.default_rng(0).integers(0, len(p2), 4096)
        assert len(p2) == len(wl["points"]) and np.array_equal(wl["points"][node_of_p2[pick]], p2[pick])
    return stress, c_oracle.COracle(stress), c_oracle.max_threads()


def run_reference(args, rank: int, world: int) -> None:
    """CPU arm: the restated reference algorithm (oracle/hemo_oracle.c, all host threads) on the same workload."""
    if rank != 0:
        return
    from oracle import hemo_oracle as ho
    wl = build_workload(args.workload, args.snapshots, 0)
    stress, co, threads = oracle_for(wl)
    n = len(wl["points"])
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
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, wl, stress.nF, args.snapshots, world, args.scaling),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = FEniCS path restated in C (oracle/hemo_oracle.c); dolfin itself is not installable"}
    print(json.dumps(line), flush=True)


def workload_config(name, wl, nF, n_snap, world, scaling):
    """The keys both arms of the bench share (the driver compares them)."""
    return {"workload": f"{name}: {wl['desc']}", "facets": int(nF), "snapshots_per_gpu": int(n_snap),
            "order": int(wl["order"]), "tets": int(len(wl["tets"])), "velocity_nodes": int(len(wl["points"])),
            "parallelism": f"time-shard x{world}", "scaling": scaling}


def kernel_traffic(name: str):
    """DRAM bytes per launch of K1 / K2 from the committed `ncu --set full` captures of this command."""
    tfile = ROOT / "profiles" / "kernel_traffic.json"
    try:
        return json.loads(tfile.read_text()).get(name, {})
    except Exception:
        return {}


def measure(name: str, n_res: int, n_e2e: int, args, rank: int, local_rank: int, world: int, headline: bool):
    """One workload on this rank's GPU: device-resident pass (`value`), per-kernel pass (roofline), end-to-end pass from
    pinned host vectors, CPU baseline.  Returns (entry dict, context) on rank 0, (None, context) elsewhere."""
    from vasp_b200.engine import HemoEngine, pinned_empty
    from vasp_b200.timeshard import NcclComm, plan_shard

    t_setup = time.perf_counter()
    g = build_geometry(name)
    order = g["order"]
    eng = HemoEngine(local_rank)
    eng.set_mesh(g["xyz"], g["tets"])
    eng.set_velocity_layout(order, refined_xyz=g["points"] if order == 2 else None)
    if args.compaction != "auto":
        eng.set_host_compaction(args.compaction)
    nF, vec_len, n_nodes = eng.nF, eng.vec_len, len(g["points"])
    comm = NcclComm(eng, rank, world) if world > 1 else None
    compact = eng.compaction_active

    # this rank's snapshots of one long series: weak scaling gives every rank n_res snapshots (the series grows with
    # the world), strong scaling splits n_res over the ranks; every rank but the first reads one halo snapshot
    if args.scaling == "strong":
        sh = plan_shard(n_res, rank, world)
        first, n_mine, n_total, period = sh.start, sh.count, n_res, n_res
    else:
        first, n_mine, n_total, period = rank * n_res, n_res, n_res * world, n_res
    halo = 1 if rank > 0 else 0
    coef, dt = series_coefficients(n_mine + halo, first - halo, period)

    # Resident input = what the product's host -> device path leaves in HBM for this workload: whole vectors (K1
    # gathers the wall layer) or, when the wall layer is gathered in front of the bus, compact blocks (K1 transposes).
    if compact:
        slots = eng.wall_slots()  # blocked layout: slot == node number
        nwp = eng.compact_len // 3
        idx = np.concatenate([slots, np.full(nwp - len(slots), slots[-1])])
        flat = np.ascontiguousarray(g["basis"][:, :, idx]).reshape(synth.N_MODES, 3 * nwp)
        row_len = 3 * nwp
    else:
        flat = g["basis"].reshape(synth.N_MODES, 3 * n_nodes)
        row_len = vec_len
    resident_bytes = (n_mine + halo) * row_len * 8
    # Timing rule: inputs must not be served from L2 across timed steps.  A workload smaller than 2x the 126 MB L2 gets
    # enough rotating copies to exceed 2.5x L2 (consecutive steps read different copies, so the K steps run back to
    # back with no flush kernel and no host synchronisation in the timed region); larger workloads evict themselves.
    L2 = 126e6
    n_copies = 1 if resident_bytes >= 2 * L2 else min(8, int(np.ceil(2.5 * L2 / resident_bytes)))
    d_copies = [eng.device_alloc(resident_bytes) for _ in range(n_copies)]
    chunk = max(1, int((256 << 20) // (row_len * 8)))
    for a in range(0, n_mine + halo, chunk):  # synthesised and uploaded in pieces: 10 M tets x 256 snapshots = 6 GB
        rows = coef[a:a + chunk] @ flat
        for d in d_copies:
            eng.h2d(d + a * row_len * 8, rows)
    del flat
    flags = 2 if halo else 1
    stride = row_len * 8
    step_no = [0]

    # --wss steps|matrix: K2 also writes tau of every snapshot (72 B per facet and snapshot), as one dolfin vector
    # per snapshot (WSS.h5) or as rows of the time-major matrix of the spectral tools
    d_wss = eng.device_alloc(72 * nF * n_mine) if args.wss != "none" else 0
    push_resident = eng.push_compact_device if compact else eng.push_device

    def step_resident():
        d_u = d_copies[step_no[0] % n_copies]
        step_no[0] += 1
        eng.begin(MU, dt)
        if args.wss == "matrix":
            eng.set_wss_layout(n_mine, 0)
        push_resident(d_u, n_mine + halo, stride, flags, d_wss)
        if comm:
            comm.reduce_finalize(n_total, host=False)  # fused peer reduction + K4 (or ncclAllReduce + K4)
        else:
            eng.finalize_async(n_total)

    steps = args.steps if headline else max(3, min(args.steps, args.other_steps))
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
        for _ in range(steps):
            step_resident()
        total_ms = eng.timer_stop()  # CUDA events on the compute stream; stop synchronises
        launches1 = eng.timers()["launches"]
        # a short timed region gets a second, untimed stretch of ~50 ms so that NVML sees clocks under load (the same
        # number of steps on every rank: each step is a collective)
        extra = 0.0 if total_ms >= 25.0 else 50.0 * steps / max(total_ms, 1e-3)
        if comm:
            extra = comm.max(extra)
        for _ in range(int(min(extra, 5000))):
            step_resident()
        eng.sync()
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
        for _ in range(steps):
            step_resident()
        total_ms_events = eng.timer_stop()
    if comm:
        comm.barrier()
    k1_ms, k2_ms, k_n = eng.kernel_profile()
    eng.set_profile(False)
    launches = launches1 - launches0  # k1 + k2 + k3 (+ k4 / the peer reduction) per step
    if comm:
        total_ms = comm.max(total_ms)
    ms_per_step = total_ms / steps
    value = n_total * nF / (ms_per_step * 1e-3) if args.scaling == "strong" else world * nF * n_mine / (ms_per_step * 1e-3)

    # parity of the reduced fields against the restated reference over the whole series (N > 1: SCALE lines carry it)
    parity = None
    if comm and args.parity and n_total * nF <= 2.0e9:
        eng.begin(MU, dt)
        push_resident(d_copies[0], n_mine + halo, stride, flags, 0)
        fields = comm.reduce_finalize(n_total)
        if rank == 0:
            parity = parity_against_oracle(g, eng, fields, n_total, period, compact)
    for d in d_copies + ([d_wss] if d_wss else []):
        eng.device_free(d)

    # end to end through the public API: pinned host snapshot VECTORS in, five result fields out
    n_e2e = int(min(n_e2e, n_mine, max(4, 12e9 // (vec_len * 8))))  # at most ~12 GB of pinned host vectors
    if args.e2e_snapshots:
        n_e2e = max(2, min(n_e2e, args.e2e_snapshots))
    e2e_steps = max(1, min(steps, 10 if headline else 3))
    if os.environ.get("VASP_B200_E2E_BATCH"):  # experiments: snapshots per host->device batch (default: auto)
        eng.set_tuning(batch_snapshots=int(os.environ["VASP_B200_E2E_BATCH"]))
    u_host = pinned_empty((n_e2e + halo, vec_len))
    synth.velocity_series(g["basis"], coef[:n_e2e + halo], out=u_host)

    def step_e2e():
        eng.begin(MU, dt)
        eng.push(u_host, flags=flags)
        if comm:
            return comm.reduce_finalize(n_e2e * world)
        return eng.finalize(n_e2e)

    step_e2e()
    eng.sync()
    if comm:
        comm.barrier()
    e2e_h2d_ms = e2e_kernel_ms = e2e_gather_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = step_e2e()
        tm, st = eng.timers(), eng.io_stats()  # per time loop (vh_begin resets them)
        e2e_h2d_ms += tm["h2d_ms"] / e2e_steps        # copy stream (includes waiting for gathered pieces)
        e2e_kernel_ms += tm["kernel_ms"] / e2e_steps  # compute stream (overlaps the copies)
        e2e_gather_ms += st["gather_ms"] / e2e_steps  # host threads (overlap both)
        h2d_bytes = st["h2d_bytes"]
    eng.sync()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if comm:
        e2e_s = comm.max(e2e_s)
    e2e_value = world * nF * n_e2e / e2e_s
    d2h_bytes = int(5 * 3 * nF * 8)
    osi = out["OSI"]
    sane = bool(np.isfinite(out["TAWSS"]).all() and np.nanmin(osi) >= -1e-12 and np.nanmax(osi) <= 0.5 + 1e-12)
    reduction = ("none" if not comm else "fused peer-memory reduce+finalize (NVLink, CUDA IPC)" if comm.fused
                 else "ncclAllReduce + finalize")
    if comm:
        comm.close()  # collective teardown while every rank is alive (rank 0 still has the JSON line to assemble)
    ctx = dict(g=g, coef=coef, dt=dt, halo=halo, u_host=u_host, eng=eng, n_mine=n_mine)
    if rank != 0:
        return None, ctx

    peak, peak_src = measured_peak_gbs()
    b_alg = algorithmic_bytes_per_unit(order, args.wss != "none")
    launches_per_step = max(k_n, 1) / steps  # a step may split into several column blocks
    cols = (n_mine + halo) / launches_per_step
    units_per_launch = nF * n_mine / launches_per_step
    k2_avg_ms, k1_avg_ms = k2_ms / max(k_n, 1), k1_ms / max(k_n, 1)
    nW = eng.n_wall_nodes
    traffic = kernel_traffic(name)
    kernels = {}
    # K2: the §8d figure -- 8 B x 3 components x (4 | 10) cell dofs per facet and snapshot (+72 B with the WSS series)
    k2_bytes = units_per_launch * b_alg
    # K1 moves 8 B x 3 components per wall-layer node and snapshot, read once and written once
    k1_bytes = 2 * 24 * nW * cols
    for kname, ms, nbytes in ((f"k2_wall<{order}>", k2_avg_ms, k2_bytes),
                              ("k1_stage<dense>" if compact else "k1_stage<gather>", k1_avg_ms, k1_bytes)):
        tr = traffic.get(kname.split("<")[0], {})
        dram = tr.get("dram_bytes_per_launch") if abs(tr.get("columns", -1) - cols) < 0.5 else None
        kernels[kname] = {"ms_per_launch": ms, "algorithmic_bytes_per_launch": nbytes,
                          "achieved": nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None,
                          "frac": nbytes / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
                          "dram_bytes_per_launch": dram,
                          "frac_dram": dram / (ms * 1e-3) / 1e9 / peak if dram and ms > 0 else None,
                          "traffic_source": tr.get("source") if dram else None}
    if not compact:
        # a gather fetches whole 32-byte DRAM sectors: count the distinct sectors the wall-layer nodes occupy in a
        # snapshot vector (the synthetic meshes are randomly renumbered on purpose)
        wall_nodes = eng.wall_slots()
        sectors = sum(np.unique((c * n_nodes + wall_nodes) // 4).size for c in range(3))
        sb = (32 * sectors + 24 * nW) * cols
        kernels["k1_stage<gather>"].update(sector_bytes_per_launch=sb,
                                           frac_sectors=sb / (k1_avg_ms * 1e-3) / 1e9 / peak if k1_avg_ms > 0 else None)
    dominant = max(kernels, key=lambda k: kernels[k]["ms_per_launch"])
    dk = kernels[dominant]
    cpu = cpu_baseline(dict(g, coef=coef, dt=dt, halo=halo), n_mine) if world == 1 and not args.no_cpu_baseline else None
    entry = {
        "workload": name, "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps,
        "ms_per_step_with_kernel_events": total_ms_events / steps,
        "config": dict(workload_config(name, g, nF, n_mine, world, args.scaling), wss_output=args.wss,
                       wall_layer_nodes=nW, reduction=reduction,
                       resident_input=("compact blocks (wall layer gathered in front of the bus): "
                                       f"{row_len * 8 / 1e6:.2f} MB per snapshot" if compact else
                                       f"whole snapshot vectors: {row_len * 8 / 1e6:.2f} MB per snapshot"),
                       l2=(f"{n_copies} rotating resident copies of the input ({n_copies * resident_bytes / 1e6:.0f} MB "
                           f"> 2.5 x 126 MB L2), steps back to back" if n_copies > 1 else
                           f"input {resident_bytes / 1e6:.0f} MB per step, larger than 2 x 126 MB L2"),
                       results_sane=sane, setup_s=round(time.perf_counter() - t_setup, 1)),
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps, "snapshots": n_e2e,
                "host_vector_bytes_per_step": int((n_e2e + halo) * vec_len * 8),
                "bus": "wall layer gathered on the host, compact blocks copied" if compact else "whole vectors copied",
                "copy_stream_ms_per_step": e2e_h2d_ms, "host_gather_ms_per_step": e2e_gather_ms,
                "h2d_gbs": h2d_bytes / (e2e_h2d_ms * 1e-3) / 1e9 if e2e_h2d_ms > 0 else None,
                "kernel_ms_per_step": e2e_kernel_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": dk["achieved"], "peak": peak, "unit": "GB/s",
                     "frac": dk["frac"], "traffic": dk["dram_bytes_per_launch"], "peak_source": peak_src,
                     "frac_dram": dk["frac_dram"],
                     # whole step (K1 + K2 + K3 ...) on the §8d algorithmic bytes: the figure to read first
                     "frac_step": nF * n_mine * b_alg / (ms_per_step * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_unit": b_alg, "units_per_launch": units_per_launch,
                     "columns_per_launch": cols, "launches_timed": k_n, "kernels": kernels},
        "cpu_baseline": cpu,
    }
    if parity is not None:
        entry["parity_rel_l2"] = parity
    return entry, ctx


def parity_against_oracle(g, eng, fields, n_total: int, period: int, compact: bool):
    """max over the five fields of the relative L2 distance between the cross-GPU result and the restated reference
    run over the WHOLE series on the host (chunked; the oracle reads the same container the GPUs were given)."""
    from oracle import c_oracle, hemo_oracle as ho
    wl = dict(g)
    stress, co, threads = oracle_for(wl)
    n_nodes = len(g["points"])
    if compact:
        slots = eng.wall_slots()
        nwp = eng.compact_len // 3
        flat = np.ascontiguousarray(g["basis"][:, :, slots]).reshape(synth.N_MODES, 3, len(slots))
        pad = np.zeros((synth.N_MODES, 3, nwp))
        pad[:, :, :len(slots)] = flat
        flat = pad.reshape(synth.N_MODES, 3 * nwp)
        remap = np.searchsorted(slots, stress.maps.cell_nodes[stress.maps.wall_cells])
        co._keep["wall_nodes"] = np.ascontiguousarray(remap, dtype=np.int64)
        co._maps.wall_nodes = co._keep["wall_nodes"].ctypes.data
        off = (0, nwp, 2 * nwp)
    else:
        flat = g["basis"].reshape(synth.N_MODES, 3 * n_nodes)
        off = (0, n_nodes, 2 * n_nodes)
    chunk = max(2, int((512 << 20) // (flat.shape[1] * 8)))
    sums, prev, dt = None, None, None
    for a in range(0, n_total, chunk):
        coef, dt = series_coefficients(min(chunk, n_total - a), a, period)
        r = co.run(coef @ flat, dt, off, tau_prev=prev, threads=threads)
        prev = r["tau_last"]
        sums = r if sums is None else {k: (sums[k] + r[k] if k != "tau_last" else r[k]) for k in r}
    fin = ho.finalize(sums["wss_sum"], sums["tawss_sum"], sums["twssg_sum"], n_total)
    worst = 0.0
    for k in ("TAWSS", "OSI", "RRT", "ECAP", "TWSSG"):
        a, b = np.asarray(fields[k]).ravel(), np.asarray(fin[k]).ravel()
        ok = np.isfinite(a) & np.isfinite(b)
        worst = max(worst, float(np.linalg.norm(a[ok] - b[ok]) / np.linalg.norm(b[ok])))
    return worst


# resident / end-to-end snapshot counts of the workloads a default run measures besides the headline one: >= 256
# resident snapshots (full lanes in K2), end-to-end on as many whole host vectors as stay within ~10 GB of pinned memory
OTHER_WORKLOADS = {"aneurysm_p1": (512, 512), "avf_p2": (256, 48), "vessel10m_p2": (256, 32)}


def run_ours(args, rank: int, local_rank: int, world: int) -> None:
    entry, ctx = measure(args.workload, args.snapshots, args.snapshots, args, rank, local_rank, world, True)
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": entry["value"], "unit": UNIT, "n_gpus": world, "steps": entry["steps"],
                "warmup": args.warmup, "ms_per_step": entry["ms_per_step"],
                "ms_per_step_with_kernel_events": entry["ms_per_step_with_kernel_events"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": entry["config"], "clocks": entry["clocks"], "e2e": entry["e2e"],
                "gpu_launches": entry["gpu_launches"], "roofline": entry["roofline"],
                "cpu_baseline": entry["cpu_baseline"]}
        if "parity_rel_l2" in entry:
            line["parity_rel_l2"] = entry["parity_rel_l2"]
    eng = ctx["eng"]
    if rank == 0 and world == 1 and not args.no_io_leg:
        eng.close()
        try:
            wl = dict(ctx["g"], coef=ctx["coef"], dt=ctx["dt"], halo=ctx["halo"])
            line["io"] = io_leg(wl, ctx["n_mine"], ctx["u_host"], local_rank, int(args.io_gib * (1 << 30)))
        except Exception as e:  # the headline line must not depend on the scratch file system
            line["io"] = {"error": f"{type(e).__name__}: {e}"}
    eng.close()
    del ctx, eng
    if world == 1 and not args.no_other_workloads:
        # BASELINE.json configs[2..4] ride in the same line (the driver's run must carry them)
        line["other_workloads"] = []
        import gc
        for name, (n_res, n_e2e) in OTHER_WORKLOADS.items():
            if name == args.workload:
                continue
            gc.collect()
            try:
                e, c = measure(name, n_res, n_e2e, args, rank, local_rank, world, False)
                c["eng"].close()
                del c
                line["other_workloads"].append(e)
            except Exception as ex:  # a workload that does not fit this box must not take the headline down
                line["other_workloads"].append({"workload": name, "error": f"{type(ex).__name__}: {ex}"})
    if rank == 0:
        print(json.dumps(line), flush=True)


def io_leg(wl, n_snap: int, u_host: np.ndarray, device: int, io_bytes: int = 1 << 30):
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
    n_io = int(max(3, min(n_snap, io_bytes // (vec_len * 8), len(u_host) - wl["halo"])))
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
        block = default_block_snapshots(vec_len, eng.compact_len if eng.compaction_active else 0, eng.nF)
        eng.set_tuning(batch_snapshots=block, chunk_snapshots=0)

        def run_once():
            t0 = time.perf_counter()
            series = io_dolfin.VelocitySeries(path, "velocity", 1)
            eng.begin(MU, float(series.timestamps[1] - series.timestamps[0]))
            reader = _BlockReader(series, 0, len(series), block, eng)
            push = eng.push_compact if reader.compact else eng.push
            first = True
            for a, b, u in reader:
                push(u, flags=1 if first else 0)
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
    stress, co, threads = oracle_for(wl)
    n = len(wl["points"])
    n_s = min(n_snap, max(threads, int(4.0e6 * threads / max(stress.nF, 1))))
    n_s = max(2, min(n_s, int((4 << 30) // (3 * n * 8))))  # at most ~4 GB of host vectors
    u = synth.velocity_series(wl["basis"], wl["coef"][wl["halo"]:wl["halo"] + n_s])
    co.run(u[:max(1, n_s // 8)], wl["dt"], (0, n, 2 * n), threads=threads)  # warm-up
    reps, t_best = 0, float("inf")
    t_start = time.perf_counter()
    while reps < 5 and time.perf_counter() - t_start < (20.0 if n_s * stress.nF < 5e7 else 8.0):
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
    ap.add_argument("--io-gib", type=float, default=1.0, help="size bound of the u.h5 written for the HDF5 -> device leg")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="measure the headline workload only (default: BASELINE configs[2..4] ride in the same line)")
    ap.add_argument("--other-steps", type=int, default=5, help="timed steps per non-headline workload")
    ap.add_argument("--e2e-snapshots", type=int, default=0, help="bound of the end-to-end sample (host vectors per GPU)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: every GPU processes --snapshots; strong: --snapshots are split over the GPUs")
    ap.add_argument("--no-parity", dest="parity", action="store_false",
                    help="N > 1: skip the comparison of the reduced fields with the restated reference")
    ap.add_argument("--compaction", choices=["auto", "on", "off"], default="auto",
                    help="wall-layer gather in front of the bus (default: the library's rule)")
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
