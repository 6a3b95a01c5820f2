set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sharded.log; tail -25 gpurun_out/pytest_sharded.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --scale 640 --mode sharded > gpurun_out/bench_sharded2_640.json 2> gpurun_out/bench_sharded2_640.err; echo "sharded rc=$?"; tail -5 gpurun_out/bench_sharded2_640.err; tail -c 1500 gpurun_out/bench_sharded2_640.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --scale 640 > gpurun_out/bench_replicas2_640.json 2> gpurun_out/bench_replicas2_640.err; echo "replicas rc=$?"; tail -3 gpurun_out/bench_replicas2_640.err; tail -c 600 gpurun_out/bench_replicas2_640.json
