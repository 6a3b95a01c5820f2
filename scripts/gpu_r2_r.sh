#!/bin/bash
# round 2, call R (N GPUs, default 2): parity, exchange roofline and bench line of the push kernel with match.any ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q -x > gpurun_out/r2r_pytest_${N}gpu.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2r_pytest_${N}gpu.log; tail -5 gpurun_out/r2r_pytest_${N}gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/exchange_bench.py --variants --sizes 1048576,4194304,16777216,67108864 > gpurun_out/r2r_exchange_variants_${N}gpu.json 2> gpurun_out/r2r_exchange_variants_${N}gpu.err
echo "exchange rc=$?"; tail -c 200 gpurun_out/r2r_exchange_variants_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/r2r_bench_default_${N}gpu.json 2> gpurun_out/r2r_bench_default_${N}gpu.err
echo "bench default rc=$?"; tail -c 300 gpurun_out/r2r_bench_default_${N}gpu.err; head -c 300 gpurun_out/r2r_bench_default_${N}gpu.json; echo
