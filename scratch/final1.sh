timeout 200 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/bench_r1_v8.json; tail -3 gpurun_out/bench_err.log
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-maxiter 2 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_dsd_march|k_fftx_green_p2|k_eps_dot6|k_cg_update|k_fftz_p2|k_ffts_p2" -s 12 -c 9 -o gpurun_out/prof_r01b python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-maxiter 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
