mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/exchange_bench.py > gpurun_out/exchange_bench_${N}gpu.json 2> gpurun_out/exchange_bench_${N}gpu.err; echo rc=$?; tail -2 gpurun_out/exchange_bench_${N}gpu.err; tail -1 gpurun_out/exchange_bench_${N}gpu.json
