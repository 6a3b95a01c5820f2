set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python scripts/expand_bench.py --scale 2560 --reps 5 --variants 0,1,2,3 > gpurun_out/variants_q1.log 2>&1
grep -E "total_us|CTAs" gpurun_out/variants_q1.log
timeout 600 python bench.py --steps 20 --warmup 3 --scale 160 > gpurun_out/bench_160.json 2> gpurun_out/bench_160.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_160.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['latency_us'], d['cpu_baseline']['latency_us'])
PY
