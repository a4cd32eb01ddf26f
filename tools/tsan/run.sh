#!/bin/bash
# ThreadSanitizer over the host thread pool + gather of csrc/compact.cu (CPU only).  Writes profiles/r2_tsan_host_gather.txt
set -eu
cd "$(dirname "$0")/../.."
T=$(mktemp -d)
nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=thread -o $T/host_gather_tsan tools/tsan/host_gather_tsan.cu -lpthread 2> $T/build.log || { cat $T/build.log; exit 1; }
{ echo "# nvcc -O1 -g -Xcompiler -fsanitize=thread tools/tsan/host_gather_tsan.cu  (gcc $(gcc -dumpversion), $(nproc) cores)"; TSAN_OPTIONS="halt_on_error=0 exitcode=66" $T/host_gather_tsan 2>&1; echo "exit code $?"; } | tee profiles/r2_tsan_host_gather.txt
rm -rf $T
