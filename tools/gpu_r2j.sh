#!/bin/bash
# ncu evidence at the round's final code: launch list of the default command (headline only), --set full of K1/K2 on two workloads
set -u
OUT=gpurun_out/r2j; mkdir -p $OUT
B="--steps 2 --warmup 3 --no-cpu-baseline --no-io-leg --no-other-workloads"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_stenosis_p1.csv python bench.py $B > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'k2_wall|k1_stage' -s 6 -c 2 -f -o $OUT/prof_stenosis_p1 python bench.py $B > $OUT/ncu_stenosis_p1.log 2>&1; echo "ncu p1 rc=$?"
timeout 160 ncu --set full --clock-control none --import-source on -k regex:'k2_wall|k1_stage' -s 6 -c 2 -f -o $OUT/prof_avf_p2 python bench.py --workload avf_p2 --snapshots 256 $B > $OUT/ncu_avf_p2.log 2>&1; echo "ncu avf rc=$?"
ls -la $OUT
