#!/bin/bash
set -u
OUT=gpurun_out/r2c
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
( time timeout 900 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -3 $OUT/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; echo "ref rc=$?"
