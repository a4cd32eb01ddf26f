#!/bin/bash
# Kernel timing across workloads (after the parity tests): bash tools/gpu_kbench.sh <tag> [extra bench args]
set -u
TAG=${1:-kb}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for spec in "stenosis_p1 1000" "stenosis_p2 1000" "aneurysm_p1 186" "avf_p2 62"; do
  set -- $spec
  timeout 900 python bench.py --workload $1 --snapshots $2 --steps 10 --no-cpu-baseline > $OUT/bench_$1.json 2> $OUT/bench_$1.err
  echo "$1 rc=$?"
  python - $OUT/bench_$1.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]; s=r["stage_kernel"]
    print(f'  step {d["ms_per_step"]*1e3:.1f} us  value {d["value"]/1e9:.2f} G/s  k2 {r["kernel_ms_per_launch"]*1e3:.1f} us frac {r["frac"]:.3f}  k1 {s["ms_per_launch"]*1e3:.1f} us frac {s["frac"]:.3f}  e2e {d["e2e"]["ms_per_step"]:.3f} ms  launches {d["gpu_launches"]}')
except Exception as e:
    print("  no line:", e)
PY
done
