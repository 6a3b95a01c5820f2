set -x
mkdir -p gpurun_out
WK_VARIANT=8 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_rmat.py -m gpu -x -q > gpurun_out/pytest_gpu_v8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v8.log; tail -5 gpurun_out/pytest_gpu_v8.log
WK_VARIANT=9 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_rmat.py -m gpu -x -q > gpurun_out/pytest_gpu_v9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v9.log; tail -3 gpurun_out/pytest_gpu_v9.log
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 6,8,9 > gpurun_out/variants_q1_v6.log 2>&1
grep -E "total_us|CTAs" gpurun_out/variants_q1_v6.log
grep -E '"variant": "(8)"' gpurun_out/variants_q1_v6.log | cut -c1-60,150-260
