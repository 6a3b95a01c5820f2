#!/bin/bash
# round 2, call L (1 GPU): the whole GPU suite, the bench line and the reference arm on the current tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log; tail -4 gpurun_out/r2l_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2l_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err
echo "ref rc=$?"; tail -c 300 gpurun_out/r2l_bench_ref.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2l_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2l_smoke.log
