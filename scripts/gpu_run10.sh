set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
( time timeout 1500 python scripts/rmat_scan.py --scale 24 --edges 268435456 --rbuf-gb 16 ) > gpurun_out/rmat_scan_s24.log 2>&1; tail -12 gpurun_out/rmat_scan_s24.log
( time timeout 2400 python scripts/rmat_scan.py --scale 26 --edges 1000000000 --rbuf-gb 40 ) > gpurun_out/rmat_scan_s26.log 2>&1; tail -24 gpurun_out/rmat_scan_s26.log
