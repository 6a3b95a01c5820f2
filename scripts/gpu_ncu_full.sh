set -x
mkdir -p gpurun_out
WK_VARIANT=${1:-4} timeout 1500 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 4 -o gpurun_out/prof_${2:-r1_v4} python scripts/expand_bench.py --scale 2560 --reps 2 > gpurun_out/ncu_${2:-r1_v4}.log 2>&1
tail -3 gpurun_out/ncu_${2:-r1_v4}.log
