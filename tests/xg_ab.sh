#!/bin/bash
# A/B of 1024-point passes on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/xg_ab.log; : > $L
run() { tag=$1; shift; KB_TAG="$tag" "$@" 2>&1 | tail -1 >> $L; }
run "1024x128x256 default" python tests/kbench.py 1024 128 256
run "128x1024x256 default" python tests/kbench.py 128 1024 256
run "128x1024x256 y p3" env FGB_S1024_P3=1 python tests/kbench.py 128 1024 256
run "128x256x1024 default" python tests/kbench.py 128 256 1024
run "128x256x1024 zh" env FGB_ZH=1 python tests/kbench.py 128 256 1024
cat $L
