set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -6; free -g | head -2; nproc
timeout 600 python -m pytest tests/test_sharded.py tests/test_gpu_store_build.py -m gpu -x -q > gpurun_out/pytest_4gpu.log 2>&1; tail -5 gpurun_out/pytest_4gpu.log
for N in 4; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/bench_replicas_${N}gpu.json 2> gpurun_out/bench_replicas_${N}gpu.err; echo "rc=$?"; tail -3 gpurun_out/bench_replicas_${N}gpu.err; cut -c1-900 gpurun_out/bench_replicas_${N}gpu.json
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --mode sharded --exchange p2p ) > gpurun_out/bench_sharded_p2p_${N}gpu.json 2> gpurun_out/bench_sharded_p2p_${N}gpu.err; echo "rc=$?"; tail -3 gpurun_out/bench_sharded_p2p_${N}gpu.err; cut -c1-1200 gpurun_out/bench_sharded_p2p_${N}gpu.json
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --mode sharded --exchange nccl ) > gpurun_out/bench_sharded_nccl_${N}gpu.json 2> gpurun_out/bench_sharded_nccl_${N}gpu.err; echo "rc=$?"; tail -3 gpurun_out/bench_sharded_nccl_${N}gpu.err; cut -c1-1200 gpurun_out/bench_sharded_nccl_${N}gpu.json
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_ref_${N}gpu.json
done
