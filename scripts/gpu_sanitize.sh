# memcheck over the single-GPU parity tests (small stores; the sanitizer slows kernels down 10-50x)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/sanitizer_parity.log 2>&1; echo "parity rc=$?"; tail -6 gpurun_out/sanitizer_parity.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_store_build.py -m gpu -x -q -k "same_store or queries_on" > gpurun_out/sanitizer_build.log 2>&1; echo "build rc=$?"; tail -6 gpurun_out/sanitizer_build.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/sanitizer_race.log 2>&1; echo "race rc=$?"; tail -12 gpurun_out/sanitizer_race.log
