#!/bin/bash
# Scaling evidence on N GPUs of one box (builder side).  Usage: bash tools/gpu_scale.sh <tag> <N> [with10m]
set -u
OUT=gpurun_out/$1; mkdir -p $OUT; N=$2; BIG=${3:-}
run() {  # name, bench args...
  local name=$1; shift
  if [ $N -eq 1 ]; then
    timeout 1200 python bench.py --gpus 1 --no-cpu-baseline --no-io-leg --no-other-workloads "$@" > $OUT/${name}_n$N.json 2> $OUT/${name}_n$N.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --no-cpu-baseline --no-io-leg --no-other-workloads "$@" > $OUT/${name}_n$N.json 2> $OUT/${name}_n$N.err
  fi
  echo "$name n$N rc=$?"
}
nvidia-smi topo -m > $OUT/topo_n$N.txt 2>&1; nproc >> $OUT/topo_n$N.txt
run weak_stenosis_p1 --steps 200 --warmup 3
run weak_stenosis_p2 --workload stenosis_p2 --steps 50 --warmup 3
run strong_aneurysm_p1 --workload aneurysm_p1 --snapshots 2000 --scaling strong --steps 5 --warmup 3
if [ -n "$BIG" ]; then
  if [ "$BIG" = "--no-parity" ]; then PAR="--no-parity"; else PAR=""; fi
  run strong_vessel10m_p2 --workload vessel10m_p2 --snapshots 2000 --scaling strong --steps 3 --warmup 3 $PAR
fi
python - $OUT $N <<'PY'
import json,sys,glob
for f in sorted(glob.glob(sys.argv[1]+"/*_n"+sys.argv[2]+".json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], d["n_gpus"], d["scaling"], f'step {d["ms_per_step"]:.4f} ms value {d["value"]/1e9:.2f} G/s e2e {d["e2e"]["value"]/1e9:.3f} G/s ({d["e2e"]["snapshots"]} snaps)', d["config"]["reduction"][:12], "parity", d.get("parity_rel_l2"))
    except Exception as e:
        print(f, "no line", e)
PY
