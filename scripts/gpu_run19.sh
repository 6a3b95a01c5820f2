set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py --scale 10240 --steps 5 --warmup 3 --no-cpu-baseline --plan optimal10240_plan ) > gpurun_out/bench_lubm10240_1gpu.json 2> gpurun_out/bench_lubm10240_1gpu.err; echo "rc=$?"; tail -5 gpurun_out/bench_lubm10240_1gpu.err; cut -c1-700 gpurun_out/bench_lubm10240_1gpu.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_lubm10240_1gpu.json"))
    print({k: d[k] for k in ("value", "rows", "dataset")}); print(d["latency_us"]); print(d["roofline_expand"])
except Exception as e:
    print("no json", e)
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
