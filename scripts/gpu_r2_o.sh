#!/bin/bash
# round 2, call O (N GPUs): sharded parity, the exchange roofline, the default bench line (no secondary runs), the R-MAT scan
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q -x > gpurun_out/r2o_pytest_${N}gpu.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2o_pytest_${N}gpu.log; tail -5 gpurun_out/r2o_pytest_${N}gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/exchange_bench.py --sizes 1048576,4194304,16777216,67108864 > gpurun_out/r2o_exchange_${N}gpu.json 2> gpurun_out/r2o_exchange_${N}gpu.err
echo "exchange rc=$?"; tail -c 200 gpurun_out/r2o_exchange_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/r2o_bench_default_${N}gpu.json 2> gpurun_out/r2o_bench_default_${N}gpu.err
echo "bench default rc=$?"; tail -c 300 gpurun_out/r2o_bench_default_${N}gpu.err; head -c 300 gpurun_out/r2o_bench_default_${N}gpu.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 scripts/rmat_scan_sharded.py --max-chunks 4 > gpurun_out/r2o_rmat_${N}gpu.jsonl 2> gpurun_out/r2o_rmat_${N}gpu.err
echo "rmat rc=$?"; tail -2 gpurun_out/r2o_rmat_${N}gpu.jsonl | cut -c1-400
