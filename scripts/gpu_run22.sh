mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_surface.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12
