cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_schemes.py -m gpu -x -q --tb=short 2>&1 | tail -15
