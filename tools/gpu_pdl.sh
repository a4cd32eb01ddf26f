#!/bin/bash
# PDL mask comparison.  Usage: bash tools/gpu_pdl.sh <tag> "<masks>"
set -u
OUT=gpurun_out/$1; mkdir -p $OUT
for pdl in $2; do
 for spec in "stenosis_p1 1000" "stenosis_p2 1000" "aneurysm_p1 186"; do
  WL=${spec% *}; NS=${spec#* }
  VASP_B200_PDL=$pdl timeout 600 python bench.py --workload $WL --snapshots $NS --steps 20 --no-cpu-baseline > $OUT/b_${WL}_pdl$pdl.json 2> $OUT/b_${WL}_pdl$pdl.err
  python - $OUT/b_${WL}_pdl$pdl.json "$WL pdl=$pdl" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]; s=r["stage_kernel"]
    print(f'{sys.argv[2]:24s} step {d["ms_per_step"]*1e3:7.1f} us (with events {d["ms_per_step_with_kernel_events"]*1e3:7.1f})  k2 {r["kernel_ms_per_launch"]*1e3:6.1f} us  k1 {s["ms_per_launch"]*1e3:6.1f} us  e2e {d["e2e"]["ms_per_step"]:.3f} ms')
except Exception as e:
    print("no line:", sys.argv[2], e)
PY
 done
done
