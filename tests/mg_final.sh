#!/bin/bash
# final multi-GPU measurements of round 2 (run under gpurun --gpus N): usage mg_final.sh N [weak] [strong] [strongcl] [check] [checkcl]
cd "$(dirname "$0")/.."
N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
for mode in "$@"; do
  if [ "$mode" = "weak" ]; then
    timeout 300 $TR --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu_weak.json 2> gpurun_out/bench_r02_${N}gpu_weak.err
  elif [ "$mode" = "strong" ]; then
    timeout 400 $TR --master-port 29562 bench.py --gpus $N --scaling strong --grid 1024 --warmup 3 > gpurun_out/bench_r02_${N}gpu_strong1024.json 2> gpurun_out/bench_r02_${N}gpu_strong1024.err
  elif [ "$mode" = "strongcl" ]; then          # A/B: CTA-pair x pass (128-byte peer segments at nx = 1024)
    FGB_XG_CLUSTER=1 timeout 400 $TR --master-port 29564 bench.py --gpus $N --scaling strong --grid 1024 --warmup 3 > gpurun_out/bench_r02_${N}gpu_strong1024_cluster.json 2> gpurun_out/bench_r02_${N}gpu_strong1024_cluster.err
  elif [ "$mode" = "checkcl" ]; then
    FGB_XG_CLUSTER=1 timeout 400 $TR --master-port 29565 tests/mgpu_check.py > gpurun_out/mgpu_check_${N}gpu_cluster.log 2>&1
    grep -c " OK" gpurun_out/mgpu_check_${N}gpu_cluster.log; grep "MISMATCH" gpurun_out/mgpu_check_${N}gpu_cluster.log
  elif [ "$mode" = "check" ]; then
    timeout 400 $TR --master-port 29563 tests/mgpu_check.py > gpurun_out/mgpu_check_${N}gpu.log 2>&1
    grep -c " OK" gpurun_out/mgpu_check_${N}gpu.log; grep "MISMATCH" gpurun_out/mgpu_check_${N}gpu.log
  fi
done
for f in gpurun_out/bench_r02_${N}gpu_*.json; do
python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    txt = open(f).read()
    d = json.loads(txt[txt.index('{"metric'):])
    open(f, "w").write(json.dumps(d) + "\n")          # drop the NCCL banner line
    print(f, "ms/iter %.3f" % d["ms_per_step"], "value %.3e" % d["value"], "hbm %.3f" % d["iteration_hbm"]["frac_of_peak"],
          d["nvlink"] and round(d["nvlink"]["achieved_gbs_per_direction"]), d["strong_scaling"], d["e2e"] and "e2e %.3e" % d["e2e"]["value"],
          d["e2e"] and d["e2e"]["parts"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches"])[:10]:
        print("    %-26s %4d %8.4f ms %s" % (k, v["launches"], v["avg_ms"], v["gbs"] and round(v["gbs"])))
except Exception as e:
    print(f, "FAILED", e)
    print(open(f.replace(".json", ".err")).read()[-1500:])
PY
done
