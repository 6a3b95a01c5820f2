#!/bin/bash
# round 2, call E (1 GPU): whole -m gpu suite (warp-mode light interpreter, 1024-thread server), light trace, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -4 gpurun_out/r2e_pytest.log
timeout 600 python scripts/light_trace.py > gpurun_out/r2e_light_trace.json 2> gpurun_out/r2e_light_trace.err
echo "trace rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2e_bench.json
