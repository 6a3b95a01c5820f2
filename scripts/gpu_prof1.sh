set -x
mkdir -p gpurun_out
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 > gpurun_out/expand_2560.log 2>&1
cat gpurun_out/expand_2560.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 3 -o gpurun_out/prof_r1_step python scripts/expand_bench.py --scale 640 --reps 3 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
