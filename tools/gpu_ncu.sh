#!/bin/bash
# Full ncu capture of the hot kernels for one workload:
#   bash tools/gpu_ncu.sh <tag> <workload> <snapshots> [kernel regex] [extra ncu args, e.g. "--cache-control none"]
set -u
TAG=$1; WL=$2; NS=$3; RE=${4:-k2_wall|k1_stage}; EXTRA=${5:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on $EXTRA -k regex:"$RE" -s 9 -c 3 \
    -o $OUT/prof_$WL -f python bench.py --workload $WL --snapshots $NS --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_$WL.log 2>&1
echo "ncu $WL rc=$?"
