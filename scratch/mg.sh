run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --e2e-maxiter 3 2>/dev/null | grep '^{"metric' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('$2', round(d['ms_per_step'],3), {n:round(v['avg_ms'],3) for n,v in k.items() if n in ('fft_x_green','p2p_push_bwd','fft_y_fwd_p2p','fft_y_bwd')})"; }
run 29551 memcpy_push
FGB_P2P_XSTORE=1 run 29552 xstore
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tests/mgpu_check.py 2>&1 | grep "ranks="
