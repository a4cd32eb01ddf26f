#!/bin/bash
# K2 direct-load shape for P2, then the ncu evidence: launch list of the default command, --set full of K1/K2 on four workloads
set -u
OUT=gpurun_out/r2f; mkdir -p $OUT
B="--steps 20 --no-cpu-baseline --no-io-leg --no-other-workloads"
for OCC in 3 6; do
  for spec in "stenosis_p2 1000" "avf_p2 256"; do
    WL=${spec% *}; NS=${spec#* }
    VASP_B200_K2_OCC=$OCC timeout 600 python bench.py --workload $WL --snapshots $NS $B > $OUT/occ${OCC}_$WL.json 2> $OUT/occ${OCC}_$WL.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/occ${OCC}_$WL.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("occ $OCC $WL value %.4g step %.4f ms"%(d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"],4) for k,v in r["kernels"].items()}, "sane", d["config"]["results_sane"])
except Exception as e: print("occ $OCC $WL failed", e)
PY
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_stenosis_p1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-io-leg --no-other-workloads > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
for spec in "stenosis_p1 1000" "stenosis_p2 1000" "avf_p2 256" "vessel10m_p2 256"; do
  WL=${spec% *}; NS=${spec#* }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_wall|k1_stage' -s 6 -c 2 \
      -f -o $OUT/prof_$WL python bench.py --workload $WL --snapshots $NS --steps 2 --warmup 3 --no-cpu-baseline --no-io-leg --no-other-workloads > $OUT/ncu_$WL.log 2>&1; echo "ncu $WL rc=$?"
done
ls -la $OUT/*.ncu-rep
