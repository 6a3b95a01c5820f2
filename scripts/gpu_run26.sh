mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -15
N=2
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 --mode sharded --exchange p2p ) > gpurun_out/bench_sharded_p2p_${N}gpu.json 2> gpurun_out/bench_sharded_p2p_${N}gpu.err; echo "rc=$?"; tail -6 gpurun_out/bench_sharded_p2p_${N}gpu.err | grep -v "^$" | tail -4
python - <<PY
import json
d = json.load(open("gpurun_out/bench_sharded_p2p_2gpu.json"))
print(d["value"], d["rows"], d["latency_us"])
PY
