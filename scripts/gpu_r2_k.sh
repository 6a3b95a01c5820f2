#!/bin/bash
# round 2, call K (N GPUs, default 2): sharded parity (resident sharded servers), default bench line with the launch path as
# A/B (WK_RESIDENT=0), exchange roofline, R-MAT scan with chunked second hop
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q -x > gpurun_out/r2k_pytest_${N}gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest_${N}gpu.log; tail -3 gpurun_out/r2k_pytest_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2k_bench_default_${N}gpu.json 2> gpurun_out/r2k_bench_default_${N}gpu.err
echo "bench default rc=$?"; tail -c 300 gpurun_out/r2k_bench_default_${N}gpu.err; head -c 600 gpurun_out/r2k_bench_default_${N}gpu.json; echo
WK_RESIDENT=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/r2k_bench_default_launch_${N}gpu.json 2> gpurun_out/r2k_bench_default_launch_${N}gpu.err
echo "bench (launch path) rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/exchange_bench.py > gpurun_out/r2k_exchange_${N}gpu.json 2> gpurun_out/r2k_exchange_${N}gpu.err
echo "exchange rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 scripts/rmat_scan_sharded.py > gpurun_out/r2k_rmat_${N}gpu.jsonl 2> gpurun_out/r2k_rmat_${N}gpu.err
echo "rmat rc=$?"; tail -2 gpurun_out/r2k_rmat_${N}gpu.jsonl | cut -c1-600
