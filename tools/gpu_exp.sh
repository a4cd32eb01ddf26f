#!/bin/bash
# ad-hoc timing experiments: each line "ENV=VAL ... | workload snapshots"
OUT=gpurun_out/$1; mkdir -p $OUT; shift
i=0
for spec in "$@"; do
  envs=${spec%%|*}; rest=${spec##*|}; set -- $rest; i=$((i+1))
  env $envs timeout 600 python bench.py --workload $1 --snapshots $2 --steps 10 --no-cpu-baseline > $OUT/exp_$i.json 2> $OUT/exp_$i.err
  python - $OUT/exp_$i.json "$spec" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]; s=r["stage_kernel"]
    print(f'{sys.argv[2]:60s} step {d["ms_per_step"]*1e3:7.1f} us  k2 {r["kernel_ms_per_launch"]*1e3:6.1f} us  k1 {s["ms_per_launch"]*1e3:6.1f} us')
except Exception as e:
    print("no line:", sys.argv[2], e)
PY
done
