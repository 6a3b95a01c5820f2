set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sharded.log; tail -25 gpurun_out/pytest_sharded.log
for ex in p2p nccl; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --scale 640 --mode sharded --exchange $ex > gpurun_out/bench_sharded2_640_$ex.json 2> gpurun_out/bench_sharded2_640_$ex.err; echo "sharded $ex rc=$?"; tail -3 gpurun_out/bench_sharded2_640_$ex.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_sharded2_640_$ex.json').read().strip().splitlines()[-1])
print('$ex', round(d['value'],1), round(d['e2e']['value'],1), d['latency_us']['device'], d['rows'])
PY
done
