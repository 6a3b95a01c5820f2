set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time timeout 1800 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo "ref rc=$?"
( time timeout 1800 python bench.py ) > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_r1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_r1.log 2>&1; echo "ncu rc=$?"
timeout 600 python scripts/emu_bench.py > gpurun_out/emu_bench.json 2> gpurun_out/emu_bench.err; tail -1 gpurun_out/emu_bench.json | cut -c1-400
