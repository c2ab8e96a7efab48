#!/bin/bash
# A/B of the marching sweep on one GPU: slab shapes, halo variant (FGB_MARCH_FAKE_HALO: timing only), segment lengths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/march_ab.log; : > $L
run() { tag=$1; shift; KB_TAG="$tag" "$@" 2>&1 | tail -1 >> $L; }
run "64x256x1024 kfast" python tests/kbench.py 64 256 1024
run "64x256x1024 kslow" env FGB_MARCH_KSLOW=1 python tests/kbench.py 64 256 1024
run "64x256x1024 kfast nt256" env FGB_MARCH_NT256=1 python tests/kbench.py 64 256 1024
run "128x1024x1024 halo kfast" env FGB_MARCH_FAKE_HALO=1 python tests/kbench.py 128 1024 1024
for g in "64 256 1024" "64 512 512"; do
  t=$(echo $g | tr ' ' x)
  timeout 300 ncu --set full --clock-control none -k regex:k_dsd_march -s 6 -c 1 -f -o /tmp/m_$t python tests/kbench.py $g > /dev/null 2>&1
  ncu -i /tmp/m_$t.ncu-rep --page raw --csv > gpurun_out/ncu_march_$t.csv 2>/dev/null
done
cat $L
