cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/bench_r1_v11.json; tail -3 gpurun_out/bench_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-maxiter 2 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dsd_march|k_fftx_green_p2|k_cg_update_u6|k_fftz_p2|k_ffts_p2" -s 16 -c 8 -o gpurun_out/prof_r01c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-maxiter 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
