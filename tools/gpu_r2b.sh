#!/bin/bash
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
timeout 600 python tools/e2e_variants.py --workload stenosis_p1 --snapshots 1000 --reps 5 --threads 2,4,8,16 > $OUT/variants_stenosis_p1.jsonl 2> $OUT/variants_stenosis_p1.err; echo "stenosis_p1 rc=$?"
timeout 600 python tools/e2e_variants.py --workload stenosis_p2 --snapshots 1000 --reps 5 --threads 2,4,8,16 > $OUT/variants_stenosis_p2.jsonl 2> $OUT/variants_stenosis_p2.err; echo "stenosis_p2 rc=$?"
grep -h -o '"variant": "[a-z:0-9-]*", "value": [0-9.]*, "ms_per_snapshot": [0-9.e-]*' $OUT/variants_*.jsonl
