#!/bin/bash
# round 2, call G (N GPUs, N = $1, default 8): the default bench line at N GPUs (LUBM-10240 sharded + secondary runs), the exchange
# roofline, the R-MAT scan
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_default_${N}gpu.json 2> gpurun_out/r2g_bench_default_${N}gpu.err
echo "bench default N=$N rc=$?"; tail -c 600 gpurun_out/r2g_bench_default_${N}gpu.err; head -c 1500 gpurun_out/r2g_bench_default_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 scripts/exchange_bench.py > gpurun_out/r2g_exchange_${N}gpu.json 2> gpurun_out/r2g_exchange_${N}gpu.err
echo "exchange rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 scripts/rmat_scan_sharded.py > gpurun_out/r2g_rmat_${N}gpu.jsonl 2> gpurun_out/r2g_rmat_${N}gpu.err
echo "rmat rc=$?"; tail -2 gpurun_out/r2g_rmat_${N}gpu.jsonl | cut -c1-500
