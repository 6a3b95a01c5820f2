mkdir -p gpurun_out
SCALE=2560 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/inplace_trace.py > gpurun_out/inplace_trace.json 2> gpurun_out/inplace_trace.err; echo rc=$?; tail -3 gpurun_out/inplace_trace.err; tail -1 gpurun_out/inplace_trace.json
