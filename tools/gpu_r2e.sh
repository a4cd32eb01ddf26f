#!/bin/bash
# 2 GPUs: parity tests of the time-shard path (fused reduction now on its own stream), bench at N = 1, 2
set -u
OUT=gpurun_out/r2e; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -q > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$? $(tail -1 $OUT/pytest_multi.log)"
timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-io-leg --no-other-workloads > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "n1 rc=$?"
for N in 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
     bench.py --gpus $N --steps 50 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "n$N rc=$?"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
     bench.py --gpus $N --steps 10 --warmup 3 --workload stenosis_p2 > $OUT/bench_p2_n$N.json 2> $OUT/bench_p2_n$N.err; echo "p2 n$N rc=$?"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) \
     bench.py --gpus $N --steps 5 --warmup 3 --workload aneurysm_p1 --snapshots 1000 --scaling strong > $OUT/bench_an_strong_n$N.json 2> $OUT/bench_an_strong_n$N.err; echo "an strong n$N rc=$?"
done
python - $OUT <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/bench*_n*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], d["n_gpus"], f'step {d["ms_per_step"]*1e3:.1f} us value {d["value"]/1e9:.2f} G/s e2e {d["e2e"]["value"]/1e9:.3f} G/s', d["config"]["reduction"], "parity", d.get("parity_rel_l2"))
    except Exception as e:
        print(f, "no line", e)
PY
tail -3 $OUT/*.err | tail -30
