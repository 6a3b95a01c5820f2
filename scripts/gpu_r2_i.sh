#!/bin/bash
# round 2, call I (1 GPU): whole -m gpu suite (star rounds of the light interpreter), light A/B, bench, R-MAT scan at 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -4 gpurun_out/r2i_pytest.log
timeout 600 python scripts/light_ab.py --rounds 4 > gpurun_out/r2i_light_ab.json 2> gpurun_out/r2i_light_ab.err
echo "ab rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2i_bench.json
timeout 900 python scripts/rmat_scan_sharded.py > gpurun_out/r2i_rmat_1gpu.jsonl 2> gpurun_out/r2i_rmat_1gpu.err
echo "rmat rc=$?"; tail -3 gpurun_out/r2i_rmat_1gpu.jsonl | cut -c1-400
