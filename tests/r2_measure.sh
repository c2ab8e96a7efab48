#!/bin/bash
# single-GPU measurement pass of round 2 (run under gpurun): tests, the four bench configurations, the CPU arm, ncu launch list and captures
# usage: r2_measure.sh bench | ncu   (two calls: gpurun copies back at most 64 MiB)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" = "ncu" ]; then
# the reports stay on the box (they exceed the 64 MiB that are copied back); their raw / source pages are exported as CSV
T=/tmp/fgb_prof; mkdir -p $T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cg_update_u6|k_dsd_march|k_fftx_green|k_fftz_p2|k_ffts_p2" -s 16 -c 8 -o $T/prof_r02 -f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_heat_march|k_heat_cg_u" -s 4 -c 3 -o $T/prof_r02_c3 -f python bench.py --config c3 --grid 256 --steps 4 --warmup 3 --no-e2e > gpurun_out/ncu_f3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_nh_dir_tangent|k_hyper_cg_u" -s 4 -c 3 -o $T/prof_r02_c4 -f python bench.py --config c4 --steps 4 --warmup 3 --no-e2e > gpurun_out/ncu_f4.log 2>&1
for r in prof_r02 prof_r02_c3 prof_r02_c4; do ncu -i $T/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
ncu -i $T/prof_r02.ncu-rep --page source --csv --kernel-name regex:k_fftx_green > gpurun_out/prof_r02_src_fftx_green.csv 2>/dev/null
ncu -i $T/prof_r02.ncu-rep --page source --csv --kernel-name regex:k_dsd_march > gpurun_out/prof_r02_src_dsd_march.csv 2>/dev/null
ls -la gpurun_out/*.csv
exit 0
fi
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -15 | cut -c1-250 > gpurun_out/r2_tests_final.log
tail -3 gpurun_out/r2_tests_final.log
for c in c2 c1 c3 c4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_r02_$c.json 2> gpurun_out/bench_r02_$c.err
done
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python - <<'PY'
import json
for c in ("c2", "c1", "c3", "c4"):
    try:
        d = json.load(open("gpurun_out/bench_r02_%s.json" % c))
        print(c, "ms %.4f value %.4e hbm %.3f e2e %s parity %s" % (d["ms_per_step"], d["value"], d["iteration_hbm"]["frac_of_peak"], d["e2e"] and "%.3e" % d["e2e"]["value"], d.get("parity_256") and d["parity_256"]["max_rel"]), "cpu", d.get("cpu_baseline") and "%.3e x%d" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["launches"])[:9]:
            print("    %-26s %4d %8.4f ms %s" % (k, v["launches"], v["avg_ms"], v["gbs"] and round(v["gbs"])))
    except Exception as e:
        print(c, "FAILED", e, open("gpurun_out/bench_r02_%s.err" % c).read()[-800:])
try:
    d = json.load(open("gpurun_out/bench_r02_reference_arm.json"))
    print("reference arm", d["value"], d["cpu_baseline"]["cores"], d["config"]["grid"])
except Exception as e:
    print("reference arm FAILED", e, open("gpurun_out/bench_r02_reference_arm.err").read()[-800:])
PY
