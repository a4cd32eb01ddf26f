#!/bin/bash
# K2 launch shapes (blocks per SM the kernel is compiled for) on four workloads + the big-mesh parity tests
set -u
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bigmesh.py tests/test_gpu_parity.py -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
for OCC in 3 4 5; do
  for spec in "stenosis_p1 1000" "stenosis_p2 1000" "aneurysm_p1 512" "avf_p2 256"; do
    WL=${spec% *}; NS=${spec#* }
    VASP_B200_K2_OCC=$OCC timeout 600 python bench.py --workload $WL --snapshots $NS --steps 20 --no-cpu-baseline --no-io-leg --no-other-workloads > $OUT/occ${OCC}_$WL.json 2> $OUT/occ${OCC}_$WL.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/occ${OCC}_$WL.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("occ $OCC $WL value %.4g step %.4f ms"%(d["value"], d["ms_per_step"]), {k: round(v["ms_per_launch"],4) for k,v in r["kernels"].items()}, "sane", d["config"]["results_sane"])
except Exception as e: print("occ $OCC $WL failed", e)
PY
  done
done
