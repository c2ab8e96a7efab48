cd /root/repo
timeout 900 python -m pytest tests/test_gpu_schemes.py -m gpu -x -q -k "cg_staggered" 2>&1 | tail -2
timeout 200 python scratch/marchtime.py new
FGB_MARCH_NT256=1 FGB_NO_ZH=1 timeout 200 python scratch/marchtime.py old
