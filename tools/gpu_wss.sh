#!/bin/bash
# WSS output layouts: parity tests, then K2 timing with no / per-snapshot / time-major WSS output.  Usage: bash tools/gpu_wss.sh <tag>
set -u
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
for spec in "stenosis_p1 1000" "stenosis_p2 1000" "aneurysm_p1 186"; do
 for wss in none steps matrix; do
  WL=${spec% *}; NS=${spec#* }
  timeout 600 python bench.py --workload $WL --snapshots $NS --steps 50 --wss $wss --no-cpu-baseline > $OUT/b_${WL}_$wss.json 2> $OUT/b_${WL}_$wss.err
  python - $OUT/b_${WL}_$wss.json "$WL wss=$wss" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]; s=r["stage_kernel"]
    print(f'{sys.argv[2]:28s} step {d["ms_per_step"]*1e3:7.1f} us  k2 {r["kernel_ms_per_launch"]*1e3:6.1f} us  alg {r["achieved"]:7.0f} GB/s frac {r["frac"]:.2f}  k1 {s["ms_per_launch"]*1e3:6.1f} us')
except Exception as e:
    print("no line:", sys.argv[2], e)
PY
 done
done
