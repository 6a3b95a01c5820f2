set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_store_build.py -m gpu -x -q > gpurun_out/pytest_build.log 2>&1; tail -30 gpurun_out/pytest_build.log
timeout 900 python scripts/build_bench.py --scale 2560 > gpurun_out/build_bench.log 2>&1; tail -12 gpurun_out/build_bench.log
