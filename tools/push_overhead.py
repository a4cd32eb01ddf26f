#!/usr/bin/env python
"""What does one push cost when it carries only a few snapshots?  (builder-side timing; aneurysm_p1 compact rows)"""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from vasp_b200 import synth
from vasp_b200.engine import HemoEngine, pinned_empty

g = bench.build_geometry("aneurysm_p1")
eng = HemoEngine(0)
eng.set_mesh(g["xyz"], g["tets"])
eng.set_velocity_layout(1)
n = 128
coef, dt = bench.series_coefficients(n, 0, n)
slots = eng.wall_slots(); nwp = eng.compact_len // 3
idx = np.concatenate([slots, np.full(nwp - len(slots), slots[-1])])
flat = np.ascontiguousarray(g["basis"][:, :, idx]).reshape(synth.N_MODES, 3 * nwp)
c = pinned_empty((n, 3 * nwp)); c[:] = coef @ flat
for block in (2, 4, 16, 64):
    eng.set_tuning(batch_snapshots=block)
    for rep in range(2):
        eng.begin(bench.MU, dt)
        t0 = time.perf_counter(); calls = []
        for a in range(0, n, block):
            t1 = time.perf_counter()
            eng.push_compact(c[a:a + block], flags=1 if a == 0 else 0)
            calls.append(time.perf_counter() - t1)
        out = eng.finalize(n)
        tot = time.perf_counter() - t0
    tm = eng.timers()
    print(json.dumps({"block": block, "pushes": len(calls), "total_ms": 1e3 * tot, "ms_per_push": 1e3 * float(np.mean(calls)),
                      "ms_per_push_min": 1e3 * float(np.min(calls)), "copy_stream_ms_per_push": tm["h2d_ms"] / len(calls),
                      "kernel_ms_per_push": tm["kernel_ms"] / len(calls)}), flush=True)
