set -x
mkdir -p gpurun_out
WK_VARIANT=6 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v6.log 2>&1; echo "pytest v6 rc=$?" >> gpurun_out/pytest_gpu_v6.log; tail -15 gpurun_out/pytest_gpu_v6.log
WK_VARIANT=7 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v7.log 2>&1; echo "pytest v7 rc=$?" >> gpurun_out/pytest_gpu_v7.log; tail -3 gpurun_out/pytest_gpu_v7.log
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 4,6,7 > gpurun_out/variants_q1_v5.log 2>&1
grep -E "total_us|CTAs" gpurun_out/variants_q1_v5.log
grep -E '"variant": "(6|7)"' gpurun_out/variants_q1_v5.log | cut -c1-60,150-260
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 4,6,7 --query 7 > gpurun_out/variants_q7_v5.log 2>&1
grep -E "total_us" gpurun_out/variants_q7_v5.log
grep -E '"variant": "(6)"' gpurun_out/variants_q7_v5.log | cut -c1-60,150-260
