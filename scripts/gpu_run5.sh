set -x
mkdir -p gpurun_out
WK_VARIANT=5 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v5.log 2>&1; echo "pytest v5 rc=$?" >> gpurun_out/pytest_gpu_v5.log; tail -15 gpurun_out/pytest_gpu_v5.log
WK_VARIANT=4 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v4.log 2>&1; echo "pytest v4 rc=$?" >> gpurun_out/pytest_gpu_v4.log; tail -3 gpurun_out/pytest_gpu_v4.log
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 2,4,5 > gpurun_out/variants_q1_v4.log 2>&1
grep -E "total_us|CTAs" gpurun_out/variants_q1_v4.log
grep '"variant": "5"' gpurun_out/variants_q1_v4.log
