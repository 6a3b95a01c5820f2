mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/light_cold.py > gpurun_out/light_cold.json 2> gpurun_out/light_cold.err; tail -3 gpurun_out/light_cold.err; cat gpurun_out/light_cold.json
