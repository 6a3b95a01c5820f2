#!/bin/bash
# round 2, call S (1 GPU): final validation of the tree: GPU suite, bench line, reference arm, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log; tail -4 gpurun_out/r2s_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; tail -c 200 gpurun_out/r2s_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err
echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2s_smoke.log
timeout 120 python scripts/push_ncu.py > gpurun_out/r2s_push_plain.log 2>&1; tail -2 gpurun_out/r2s_push_plain.log
