#!/bin/bash
# One GPU-box pass: parity tests, the default bench line, the launch list and a full ncu capture of the hot kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag>
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 600 python bench.py > $OUT/bench_stenosis_p1.json 2> $OUT/bench_stenosis_p1.err; echo "bench rc=$?"
cat $OUT/bench_stenosis_p1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_wall|k1_stage' -s 6 -c 3 \
    -o $OUT/prof_stenosis_p1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
