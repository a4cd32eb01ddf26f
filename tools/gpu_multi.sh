#!/bin/bash
# Multi-GPU pass: 2-rank parity tests, then the bench at N = 2 .. $2 launched as the driver does.  Usage: bash tools/gpu_multi.sh <tag> <max N>
set -u
OUT=gpurun_out/$1; mkdir -p $OUT; NMAX=${2:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$? $(tail -1 $OUT/pytest_multi.log)"
timeout 300 python bench.py --steps 20 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  [ $N -le $NMAX ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
     bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "n$N rc=$?"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
     bench.py --gpus $N --steps 10 --warmup 3 --workload stenosis_p2 > $OUT/bench_p2_n$N.json 2> $OUT/bench_p2_n$N.err; echo "p2 n$N rc=$?"
done
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/bench*_n*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], d["n_gpus"], f'step {d["ms_per_step"]*1e3:.1f} us value {d["value"]/1e9:.2f} G/s e2e {d["e2e"]["value"]/1e9:.3f} G/s', d["config"]["reduction"])
    except Exception as e:
        print(f, "no line", e)
PY
