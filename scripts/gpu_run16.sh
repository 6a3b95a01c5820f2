set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time timeout 1800 python bench.py ) > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
