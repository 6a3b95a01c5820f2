set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 1,2,3 > gpurun_out/variants_q1_v3.log 2>&1
grep -E "total_us|CTAs" gpurun_out/variants_q1_v3.log
grep '"variant": "2"' gpurun_out/variants_q1_v3.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 4 -o gpurun_out/prof_r1_step_v3 python scripts/expand_bench.py --scale 2560 --reps 2 > gpurun_out/ncu_full_v3.log 2>&1
tail -3 gpurun_out/ncu_full_v3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_l40_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l40_v3.log 2>&1
