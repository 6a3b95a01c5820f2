# 2-GPU run (gpurun --gpus 2): sharded parity tests, exchange roofline, sharded bench with both exchange paths
mkdir -p gpurun_out
N=2
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/exchange_bench.py > gpurun_out/exchange_bench_${N}gpu.json 2> gpurun_out/exchange_bench_${N}gpu.err
for EX in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 --mode sharded --exchange $EX > gpurun_out/bench_sharded_${EX}_${N}gpu.json 2> gpurun_out/bench_sharded_${EX}_${N}gpu.err
done
