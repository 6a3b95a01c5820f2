set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python scripts/emu_bench.py --scale 2560 > gpurun_out/emu_2560.json 2> gpurun_out/emu_2560.err; echo "emu rc=$?"; tail -3 gpurun_out/emu_2560.err; cat gpurun_out/emu_2560.json
