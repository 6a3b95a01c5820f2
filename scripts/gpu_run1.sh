set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/env.txt 2>&1
nproc >> gpurun_out/env.txt; free -g >> gpurun_out/env.txt; lscpu | head -20 >> gpurun_out/env.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python __graft_entry__.py --smoke > gpurun_out/memcheck_run.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_run.log
tail -15 gpurun_out/memcheck.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
