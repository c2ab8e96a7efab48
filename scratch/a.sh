cd /root/repo
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-maxiter 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], {k:round(x['avg_ms'],3) for k,x in d['kernels'].items()})"
