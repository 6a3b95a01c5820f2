mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print(d["value"], d["e2e"]["value"], d["gpu_launches"]); print(d["latency_us"]); print(d["dataset"])
PY
