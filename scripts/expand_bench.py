#!/usr/bin/env python
"""Micro-benchmark of single pattern steps on a LUBM store (for ncu captures and kernel tuning).

  python scripts/expand_bench.py --scale 2560 --reps 5
Runs, through the C ABI primitives: i2u(GraduateStudent) -> k2u(memberOf) -> k2u(undergraduateDegreeFrom)
 -> k2c(type University) -> k2k(subOrganizationOf IN)   (= LUBM Q1, osdi16 plan) and prints per-step
CUDA-event time, algorithmic bytes and GB/s.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=640)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--query", type=int, default=1)
ap.add_argument("--plan", default="osdi16_plan")
ap.add_argument("--variants", default="")
args = ap.parse_args()

from wukong_b200 import capi, datagen, host  # noqa: E402
from conftest import load_query  # noqa: E402

t0 = time.time()
tr = datagen.lubm(args.scale, seed=1)
gst = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS)      # device-side store build, as in bench.py
print("dataset ready in %.1fs: %d triples" % (time.time() - t0, tr.shape[0]), file=sys.stderr)
pats, nvars, req, _ = load_query(args.query, args.plan)
variants = [v for v in args.variants.split(",") if v != ""] or [os.environ.get("WK_VARIANT", "")]
for var in variants:
    if var != "":
        os.environ["WK_VARIANT"] = var
    os.environ["WK_VERBOSE"] = "1"
    eng = capi.Engine(gst, rbuf_bytes=max(256 << 20, tr.shape[0] * 8))
    eng.set_profiling(2)
    agg = {}
    for rep in range(args.reps):
        eng.flush_l2()
        rc, rows, cols, _ = eng.query(pats, nvars, req, blind=True)
        assert rc == 0
        for i, st in enumerate(eng.step_stats()):
            agg.setdefault(i, []).append(st)
    tot = 0.0
    for i, lst in sorted(agg.items()):
        us = sorted(x["device_us"] for x in lst)[len(lst) // 2]
        tot += us
        st = lst[0]
        print(json.dumps({"variant": var, "step": i, "kind": st["kind"], "in_rows": st["in_rows"], "in_cols": st["in_cols"],
                          "out_rows": st["out_rows"], "buckets": st["buckets_visited"], "edges": st["edges_touched"],
                          "algo_bytes": st["algo_bytes"], "median_us": round(us, 2),
                          "gbs": round(st["algo_bytes"] / us / 1e3, 1) if us else None}))
    print(json.dumps({"variant": var, "total_us": round(tot, 1)}))
    eng.close()
