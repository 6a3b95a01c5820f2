#!/bin/bash
# round 2, call M (N GPUs, default 2): where the exchange's time goes (reservation granularity, CTAs per SM, timing-only debug
# modes), then the default bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/exchange_bench.py --variants --sizes 4194304,16777216,67108864 > gpurun_out/r2m_exchange_variants_${N}gpu.json 2> gpurun_out/r2m_exchange_variants_${N}gpu.err
echo "exchange variants rc=$?"; tail -c 300 gpurun_out/r2m_exchange_variants_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2m_bench_default_${N}gpu.json 2> gpurun_out/r2m_bench_default_${N}gpu.err
echo "bench default rc=$?"; tail -c 300 gpurun_out/r2m_bench_default_${N}gpu.err; head -c 400 gpurun_out/r2m_bench_default_${N}gpu.json; echo
