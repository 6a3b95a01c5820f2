set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
( time timeout 1500 python scripts/rmat_scan.py --scale 24 --edges 268435456 --rbuf-gb 16 ) > gpurun_out/rmat_scan_s24.log 2>&1; tail -12 gpurun_out/rmat_scan_s24.log
timeout 900 python bench.py --steps 20 --warmup 3 --scale 640 --no-cpu-baseline > gpurun_out/bench_640.json 2> gpurun_out/bench_640.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_640.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['latency_us'])
PY
