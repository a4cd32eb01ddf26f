#!/bin/bash
# Bench at N ranks launched as the driver does.  Usage: bash tools/gpu_n.sh <tag> <N> [bench args]
set -u
OUT=gpurun_out/$1; mkdir -p $OUT; N=$2; shift; shift
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
   bench.py --gpus $N --steps 200 --warmup 3 "$@" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "n$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
   bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err; echo "ref n$N rc=$?"
grep -h '^{' $OUT/bench_n$N.json $OUT/ref_n$N.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d.get('impl','ours'), d['n_gpus'], 'step us', d['ms_per_step']*1e3, 'value G/s', d['value']/1e9, 'e2e G/s', d['e2e']['value']/1e9, d['config'].get('reduction'))
"
