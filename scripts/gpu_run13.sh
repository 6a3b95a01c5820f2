set -x
mkdir -p gpurun_out
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 6,8,9 > gpurun_out/variants_q1_v6.log 2>&1
grep -E "total_us|CTAs" gpurun_out/variants_q1_v6.log
grep -E '"variant": "(8)"' gpurun_out/variants_q1_v6.log | cut -c1-60,150-260
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 6,8 --query 7 > gpurun_out/variants_q7_v6.log 2>&1
grep -E "total_us" gpurun_out/variants_q7_v6.log
