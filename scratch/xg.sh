for T in 8 328 324; do FGB_XG_T=$T timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-maxiter 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('T=$T', d['ms_per_step'], d['kernels']['fft_x_green']['avg_ms'])"; done
