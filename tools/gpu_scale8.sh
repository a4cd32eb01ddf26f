#!/bin/bash
# 8-GPU box: N = 8 and N = 4 of every scaling line (the north-star 10 M-tet case with the parity check at N = 8)
set -u
bash tools/gpu_scale.sh r2scale 8 --parity-default
bash tools/gpu_scale.sh r2scale 4 --no-parity
