#!/bin/bash
# A/B of the marching sweep on one GPU: slab shapes, halo variant (FGB_MARCH_FAKE_HALO: timing only), segment lengths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/march_ab.log; : > $L
run() { tag=$1; shift; KB_TAG="$tag" "$@" 2>&1 | tail -1 >> $L; }
python -m pytest tests/test_gpu_schemes.py -q -x -k "cg_staggered" 2>&1 | tail -3 >> $L
run "256^3" python tests/kbench.py 256
run "64x256x1024" python tests/kbench.py 64 256 1024
run "64x256x1024 nt256" env FGB_MARCH_NT256=1 python tests/kbench.py 64 256 1024
run "128x1024x1024 halo" env FGB_MARCH_FAKE_HALO=1 python tests/kbench.py 128 1024 1024
cat $L
