#!/bin/bash
# Slim GPU-box pass for a checkpoint: parity tests, default bench line + reference arm, launch list, one full ncu
# capture of the hot kernels on the headline workload.  Usage: bash tools/gpu_head.sh <tag>
set -u
TAG=${1:-head}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi.csv 2>&1
nproc > $OUT/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(grep -E 'passed|failed|error' $OUT/pytest.log | tail -1)"
timeout 600 python bench.py > $OUT/bench_stenosis_p1.json 2> $OUT/bench_stenosis_p1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; echo "ref rc=$?"
timeout 600 python bench.py --workload stenosis_p2 --no-cpu-baseline > $OUT/bench_stenosis_p2.json 2> $OUT/bench_stenosis_p2.err; echo "p2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-io-leg > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_wall|k1_stage|k3_fold' -s 12 -c 3 \
    -f -o $OUT/prof_stenosis_p1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-io-leg > $OUT/ncu_stenosis_p1.log 2>&1; echo "ncu rc=$?"
cat $OUT/bench_stenosis_p1.json
