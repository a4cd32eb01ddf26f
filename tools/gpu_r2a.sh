#!/bin/bash
# round 2, first GPU pass: parity suite, then how a snapshot should cross the bus (tools/e2e_variants.py)
set -u
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.csv 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt; lscpu | head -30 >> $OUT/host.txt; numactl -H >> $OUT/host.txt 2>&1
nvidia-smi topo -m >> $OUT/host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
timeout 600 python tools/e2e_variants.py --workload aneurysm_p1 --snapshots 128 > $OUT/variants_aneurysm_p1.jsonl 2> $OUT/variants_aneurysm_p1.err; echo "aneurysm rc=$?"
timeout 900 python tools/e2e_variants.py --workload avf_p2 --snapshots 48 > $OUT/variants_avf_p2.jsonl 2> $OUT/variants_avf_p2.err; echo "avf rc=$?"
timeout 900 python tools/e2e_variants.py --workload vessel10m_p2 --snapshots 32 > $OUT/variants_vessel10m_p2.jsonl 2> $OUT/variants_vessel10m_p2.err; echo "10m rc=$?"
tail -n 3 $OUT/variants_*.jsonl
