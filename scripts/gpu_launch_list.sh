set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
( time timeout 1800 python bench.py ) > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_r1.err
( time timeout 1800 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo "ref rc=$?"; tail -4 gpurun_out/bench_r1_ref.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_r1.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 4 -o gpurun_out/prof_r1_v5 python scripts/expand_bench.py --scale 2560 --reps 2 > gpurun_out/ncu_r1_v5.log 2>&1
