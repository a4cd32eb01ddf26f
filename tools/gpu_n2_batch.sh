#!/bin/bash
# e2e of two concurrent ranks: geometric H2D batches (auto) against eight equal ones.  Usage: bash tools/gpu_n2_batch.sh <tag>
OUT=gpurun_out/$1; mkdir -p $OUT
for rep in 1 2; do for b in 0 125; do
  VASP_B200_E2E_BATCH=$b timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700+rep*10+b%7)) \
     bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/b${b}_$rep.json 2> $OUT/b${b}_$rep.err
  grep -h '^{' $OUT/b${b}_$rep.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; print('batch $b rep $rep', 'e2e ms', round(e['ms_per_step'],4), 'h2d GB/s', round(e['h2d_gbs'],2), 'e2e G/s', round(e['value']/1e9,3))"
done; done
