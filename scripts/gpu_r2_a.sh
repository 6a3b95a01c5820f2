#!/bin/bash
# round 2, call A (1 GPU): the whole -m gpu suite, then the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_env.txt 2>&1
nproc >> gpurun_out/r2a_env.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2a_env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.json
