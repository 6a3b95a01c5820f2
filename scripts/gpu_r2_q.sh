#!/bin/bash
# round 2, call Q (1 GPU): final validation of the tree (GPU suite, bench line, reference arm, smoke) + an ncu capture of the
# exchange's push kernel on a one-rank group
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -4 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
echo "bench rc=$?"; tail -c 200 gpurun_out/r2q_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err
echo "ref rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2q_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2q_smoke.log
timeout 120 python scripts/push_ncu.py > gpurun_out/r2q_push_plain.log 2>&1; tail -3 gpurun_out/r2q_push_plain.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:p2p_push -c 2 -o gpurun_out/r2q_push -f python scripts/push_ncu.py > gpurun_out/r2q_push_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2q_push_ncu.log
