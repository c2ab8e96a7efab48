N=8
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tests/mgpu_check.py 2>&1 | grep "ranks=\|Error\|error" | head -20
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b$N.log 2>&1; echo EXIT $?; grep -c metric gpurun_out/b$N.log
