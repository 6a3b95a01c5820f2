set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
( time timeout 1500 python scripts/rmat_scan.py --scale 24 --edges 268435456 --rbuf-gb 16 ) > gpurun_out/rmat_scan_s24.log 2>&1; tail -30 gpurun_out/rmat_scan_s24.log
