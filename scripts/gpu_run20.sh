set -x
mkdir -p gpurun_out
N=4
for EX in p2p nccl; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --scale 10240 --plan optimal10240_plan --steps 5 --warmup 3 --mode sharded --exchange $EX ) > gpurun_out/bench_lubm10240_sharded_${EX}_${N}gpu.json 2> gpurun_out/bench_lubm10240_sharded_${EX}_${N}gpu.err; echo "rc=$?"; tail -4 gpurun_out/bench_lubm10240_sharded_${EX}_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_lubm10240_sharded_${EX}_${N}gpu.json"))
    print({k: d.get(k) for k in ("value", "rows", "dataset", "latency_us", "exchange")})
except Exception as e:
    print("no json", e)
PY
done
