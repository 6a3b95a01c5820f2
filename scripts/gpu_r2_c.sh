#!/bin/bash
# round 2, call C (2 GPUs): multi-process sharded parity, exchange roofline, sharded bench lines
mkdir -p gpurun_out
N=2
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -3 gpurun_out/r2c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/exchange_bench.py > gpurun_out/r2c_exchange_${N}gpu.json 2> gpurun_out/r2c_exchange_${N}gpu.err
echo "exchange rc=$?"; cat gpurun_out/r2c_exchange_${N}gpu.json | head -c 1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --scale 2560 --plan osdi16_plan --no-secondary > gpurun_out/r2c_bench_sharded2560_${N}gpu.json 2> gpurun_out/r2c_bench_sharded2560_${N}gpu.err
echo "bench2560 rc=$?"; tail -c 400 gpurun_out/r2c_bench_sharded2560_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c_bench_default_${N}gpu.json 2> gpurun_out/r2c_bench_default_${N}gpu.err
echo "bench default rc=$?"; tail -c 400 gpurun_out/r2c_bench_default_${N}gpu.err
