#!/usr/bin/env python
"""How should a snapshot cross the bus?  Times one pass of the hot path from pinned HOST vectors to the five result
fields on the host, for the ways of moving the wall layer that VERDICT r1 item 3 asks to compare:

  full      whole vectors by cudaMemcpyAsync, K1 gathers on the device            (round 1)
  gather:T  wall layer gathered on the host by T threads into a pinned ring, K1 transposes (csrc/compact.cu)
  zerocopy  K1 gathers straight out of the mapped pinned host vectors over PCIe, nothing is staged
  compact   the caller already holds compact blocks (what the u.h5 reader produces from the page cache): bus only

    python tools/e2e_variants.py --workload avf_p2 --snapshots 48 [--reps 3]

Prints one JSON line per variant (facet x snapshots / s, ms per snapshot, bytes over the bus).  Builder-side tool: its
output is kept under profiles/.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from vasp_b200 import synth  # noqa: E402
from vasp_b200.engine import HemoEngine, pinned_empty  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="avf_p2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--snapshots", type=int, default=48)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--threads", default="4,8,16,32")
    args = ap.parse_args()
    t0 = time.perf_counter()
    wl = bench.build_workload(args.workload, args.snapshots, 0)
    eng = HemoEngine(0)
    eng.set_mesh(wl["xyz"], wl["tets"])
    eng.set_velocity_layout(wl["order"], refined_xyz=wl["points"] if wl["order"] == 2 else None)
    n_snap, nF, vec_len = args.snapshots, eng.nF, eng.vec_len
    u = pinned_empty((n_snap, vec_len))
    synth.velocity_series(wl["basis"], wl["coef"], out=u)
    setup_s = time.perf_counter() - t0
    base = {"workload": args.workload, "facets": nF, "snapshots": n_snap, "velocity_nodes": eng.n_nodes,
            "wall_layer_nodes": eng.n_wall_nodes, "vector_MB": vec_len * 8 / 1e6,
            "compact_MB": eng.compact_len * 8 / 1e6, "setup_s": round(setup_s, 1)}
    ref = {}

    def timed(label, fn, extra=None):
        fn()  # warm-up (allocations)
        eng.sync()
        best, out = float("inf"), None
        for _ in range(args.reps):
            t = time.perf_counter()
            out = fn()
            best = min(best, time.perf_counter() - t)
        st, tm = eng.io_stats(), eng.timers()
        if not ref:
            ref.update(out)
        same = all(np.array_equal(out[k], ref[k], equal_nan=True) for k in ref)
        line = dict(base, variant=label, value=nF * n_snap / best, ms_per_snapshot=1e3 * best / n_snap,
                    bus_MB_per_snapshot=st["h2d_bytes"] / n_snap / 1e6, gather_ms_per_snapshot=st["gather_ms"] / n_snap,
                    copy_stream_ms_per_snapshot=tm["h2d_ms"] / n_snap, kernel_ms_per_snapshot=tm["kernel_ms"] / n_snap,
                    bitwise_equal_to_full=bool(same))
        if extra:
            line.update(extra)
        print(json.dumps(line), flush=True)

    def run_push(mode, threads=0):
        def fn():
            eng.set_host_compaction(mode, threads)
            eng.begin(bench.MU, wl["dt"])
            eng.push(u, flags=1)
            return eng.finalize(n_snap)
        return fn

    timed("full", run_push("off"))
    for t in [int(x) for x in args.threads.split(",") if x]:
        timed(f"gather:{t}", run_push("on", t))

    def zerocopy():
        eng.begin(bench.MU, wl["dt"])
        eng.push_device(u.ctypes.data, n_snap, u.strides[0], 1)  # UVA: pinned host memory is device-addressable
        return eng.finalize(n_snap)
    timed("zerocopy", zerocopy)

    eng.set_host_compaction("on", 0)
    c = eng.compact(u)

    def compact():
        eng.begin(bench.MU, wl["dt"])
        eng.push_compact(c, flags=1)
        return eng.finalize(n_snap)
    timed("compact", compact)

    # the gather alone (no device work), per thread count
    for t in [int(x) for x in args.threads.split(",") if x]:
        eng.set_host_compaction("on", t)
        eng.compact(u[:2], out=c[:2])
        tt = time.perf_counter()
        eng.compact(u, out=c)
        dt = time.perf_counter() - tt
        print(json.dumps(dict(base, variant=f"gather-only:{t}", ms_per_snapshot=1e3 * dt / n_snap,
                              vector_GBps=u.nbytes / dt / 1e9, compact_GBps=c.nbytes / dt / 1e9)), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
