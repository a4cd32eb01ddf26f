#!/bin/bash
# One GPU-box pass: parity tests, the default bench line (+ reference arm), the launch list and full ncu captures
# (cold-cache replay = ncu default, and --cache-control none) of the hot kernels.  Usage: bash tools/gpu_check.sh <tag>
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
timeout 600 python bench.py > $OUT/bench_stenosis_p1.json 2> $OUT/bench_stenosis_p1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; echo "ref rc=$?"
for spec in "stenosis_p2 1000" "aneurysm_p1 186" "avf_p2 62"; do
  WL=${spec% *}; NS=${spec#* }
  timeout 900 python bench.py --workload $WL --snapshots $NS --steps 10 --no-cpu-baseline > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "$WL rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
for WL in stenosis_p1 stenosis_p2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_wall|k1_stage|k3_fold|k4_ind' -s 12 -c 4 \
      -f -o $OUT/prof_$WL python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_$WL.log 2>&1; echo "ncu $WL rc=$?"
  timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k2_wall|k1_stage' -s 12 -c 2 \
      -f -o $OUT/prof_warm_$WL python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_warm_$WL.log 2>&1; echo "ncu warm $WL rc=$?"
done
timeout 900 ncu --set full --clock-control none -k regex:'k2_wall|k1_stage' -s 12 -c 2 \
    -f -o $OUT/prof_aneurysm_p1 python bench.py --workload aneurysm_p1 --snapshots 186 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_aneurysm_p1.log 2>&1; echo "ncu aneurysm rc=$?"
