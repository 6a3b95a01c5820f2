#!/bin/bash
# round 2, call J (1 GPU): light-path parity (resident / launch / batch / sharded in-place on one GPU) + A/B + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -4 gpurun_out/r2j_pytest.log
timeout 600 python scripts/light_ab.py --rounds 4 > gpurun_out/r2j_light_ab.json 2> gpurun_out/r2j_light_ab.err
echo "ab rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2j_bench.json
