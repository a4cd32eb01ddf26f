#!/bin/bash
# Kernel timing across workloads and K2 variants (after the parity tests):
#   bash tools/gpu_kbench.sh <tag> "<variants>" ["<workload snapshots>" ...]
set -u
TAG=${1:-kb}; VARS=${2:-0}; shift; shift
if [ $# -eq 0 ]; then set -- "stenosis_p1 1000" "stenosis_p2 1000" "aneurysm_p1 186" "avf_p2 62"; fi
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in $VARS; do
  export VASP_B200_K2_VARIANT=$v VASP_B200_K2_FLAT=$v
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_v$v.log 2>&1; echo "variant $v: pytest rc=$? $(tail -1 $OUT/pytest_v$v.log)"
  for spec in "$@"; do
    WL=${spec% *}; NS=${spec#* }
    timeout 900 python bench.py --workload $WL --snapshots $NS --steps 10 --no-cpu-baseline > $OUT/bench_${WL}_v$v.json 2> $OUT/bench_${WL}_v$v.err
    python - $OUT/bench_${WL}_v$v.json $WL <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]; s=r["stage_kernel"]
    print(f'  {sys.argv[2]:12s} step {d["ms_per_step"]*1e3:7.1f} us  value {d["value"]/1e9:6.2f} G/s  k2 {r["kernel_ms_per_launch"]*1e3:6.1f} us frac {r["frac"]:.3f}  k1 {s["ms_per_launch"]*1e3:6.1f} us frac {s["frac"]:.3f}  e2e {d["e2e"]["ms_per_step"]:.3f} ms')
except Exception as e:
    print("  no line:", sys.argv[2], e)
PY
  done
done
