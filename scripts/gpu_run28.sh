mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -4
SCALE=2560 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/inplace_trace.py > gpurun_out/inplace_trace.json 2> gpurun_out/inplace_trace.err; echo rc=$?; tail -1 gpurun_out/inplace_trace.json
N=2
for EX in p2p nccl; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 --mode sharded --exchange $EX ) > gpurun_out/bench_sharded_${EX}_${N}gpu.json 2> gpurun_out/bench_sharded_${EX}_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_sharded_${EX}_2gpu.json"))
print("$EX", d["value"], d["rows"], d["latency_us"]["device"])
PY
done
