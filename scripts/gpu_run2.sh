set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_l40.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l40.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --scale 320 > gpurun_out/bench_320.json 2> gpurun_out/bench_320.err; echo "bench320 rc=$?"
( time timeout 2400 python bench.py --steps 10 --warmup 3 --scale 2560 ) > gpurun_out/bench_2560.json 2> gpurun_out/bench_2560.err; echo "bench2560 rc=$?"
tail -3 gpurun_out/bench_2560.err
