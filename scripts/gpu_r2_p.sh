#!/bin/bash
# round 2, call P (N GPUs, default 2): the default bench line (with the secondary runs) on the final tree
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2p_bench_default_${N}gpu.json 2> gpurun_out/r2p_bench_default_${N}gpu.err
echo "bench default rc=$?"; tail -c 300 gpurun_out/r2p_bench_default_${N}gpu.err; head -c 300 gpurun_out/r2p_bench_default_${N}gpu.json; echo
