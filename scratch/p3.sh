cd /root/repo
timeout 900 python -m pytest tests/test_gpu_operators.py -m gpu -x -q -k "pow2" 2>&1 | tail -5
timeout 120 python scratch/p3time2.py default
FGB_XG_P3=1 timeout 120 python scratch/p3time2.py XG_P3
FGB_NO_P3=1 timeout 120 python scratch/p3time2.py NO_P3
