cd /root/repo
timeout 900 python -m pytest tests/test_gpu_schemes.py tests/test_gpu_operators.py -m gpu -x -q -k "cg or fused or march or iso" 2>&1 | tail -3
for v in "" "FGB_MARCH_SEG=16" "FGB_MARCH_SEG=8"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-maxiter 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], {k:round(x['avg_ms'],3) for k,x in d['kernels'].items()})"
done
