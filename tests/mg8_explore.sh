#!/bin/bash
# 8-GPU exploration (run under gpurun --gpus 8): weak scaling with the default and the copy-engine transposes, the 1024^3 target, parity
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 240 $TR --master-port 29551 bench.py --gpus 8 --steps 15 --warmup 3 --no-e2e > gpurun_out/r2_w8_default.json 2> gpurun_out/r2_w8_default.err
FGB_P2P_MEMCPY=1 timeout 240 $TR --master-port 29552 bench.py --gpus 8 --steps 15 --warmup 3 --no-e2e > gpurun_out/r2_w8_memcpy.json 2> gpurun_out/r2_w8_memcpy.err
timeout 300 $TR --master-port 29553 bench.py --gpus 8 --scaling strong --grid 1024 --warmup 3 > gpurun_out/r2_s8_1024.json 2> gpurun_out/r2_s8_1024.err
timeout 300 $TR --master-port 29554 tests/mgpu_check.py > gpurun_out/r2_mgpu8.log 2>&1
for f in r2_w8_default r2_w8_memcpy r2_s8_1024; do
python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    txt = open("gpurun_out/%s.json" % f).read()
    d = json.loads(txt[txt.index('{"metric'):])
    print(f, "ms/iter %.3f" % d["ms_per_step"], "value %.3e" % d["value"], "hbm %.3f" % d["iteration_hbm"]["frac_of_peak"], d["nvlink"] and round(d["nvlink"]["achieved_gbs_per_direction"]), d["strong_scaling"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches"])[:11]:
        print("    %-26s %4d %8.4f ms %s" % (k, v["launches"], v["avg_ms"], v["gbs"] and round(v["gbs"])))
except Exception as e:
    print(f, "FAILED", e)
    print(open("gpurun_out/%s.err" % f).read()[-1500:])
PY
done
grep -v "^$" gpurun_out/r2_mgpu8.log | grep -v "\*\*\*\|OMP_NUM" | tail -22
