#!/bin/bash
# 8 GPUs, the north-star case only: 10 M tets, P2, 2000 snapshots split over the GPUs (strong scaling), compact route
set -u
OUT=gpurun_out/r2scale; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --workload vessel10m_p2 --snapshots 2000 --scaling strong --steps 3 --warmup 3 --no-parity \
  --no-cpu-baseline --no-io-leg --no-other-workloads --e2e-snapshots 4 > $OUT/strong_vessel10m_p2_n8_compact.json 2> $OUT/strong_vessel10m_p2_n8_compact.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2scale/strong_vessel10m_p2_n8_compact.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["scaling"], f'step {d["ms_per_step"]:.4f} ms value {d["value"]/1e9:.2f} G/s e2e {d["e2e"]["value"]/1e9:.3f}', {k: round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()}, d["config"]["resident_input"])
PY
