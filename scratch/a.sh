cd /root/repo
timeout 600 python -m pytest tests/test_gpu_operators.py -m gpu -x -q -k "implicit" 2>&1 | tail -2
for v in "" "FGB_XG_T=4" "FGB_XG_T=16"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-maxiter 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v',d['ms_per_step'], d['value'], {k:round(x['avg_ms'],3) for k,x in d['kernels'].items()})"
done
