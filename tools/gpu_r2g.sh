#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the parity tests; full GPU suite; default bench + reference arm
set -u
OUT=gpurun_out/r2g; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
( time timeout 900 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"; tail -4 $OUT/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; echo "ref rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_compact.py tests/test_gpu_tiny_meshes.py -m gpu -q -x > $OUT/memcheck_pytest.log 2>&1; echo "memcheck rc=$? $(tail -1 $OUT/memcheck_pytest.log)"; tail -3 $OUT/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_tiny_meshes.py "tests/test_gpu_parity.py::test_segments_passes_and_column_blocks" -m gpu -q -x > $OUT/racecheck_pytest.log 2>&1; echo "racecheck rc=$? $(tail -1 $OUT/racecheck_pytest.log)"; tail -3 $OUT/racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file $OUT/synccheck.log \
    python -m pytest tests/test_gpu_tiny_meshes.py -m gpu -q -x > $OUT/synccheck_pytest.log 2>&1; echo "synccheck rc=$? $(tail -1 $OUT/synccheck_pytest.log)"; tail -3 $OUT/synccheck.log
