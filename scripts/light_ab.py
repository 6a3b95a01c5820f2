"""A/B of the resident light-query server's shape on one box, one process: for every variant (WK_OPT_RESIDENT_VARIANT) and the
launch-per-query kernel, Q4-Q6 with the L2 flushed before every query (and unflushed), interleaved so that clock and thermal
drift hit all variants alike.  Prints median wall clock and in-kernel span per variant."""
import argparse
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import load_query  # noqa: E402
from wukong_b200 import capi, datagen, host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=2560)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
tr = datagen.lubm(a.scale, seed=1)
gst = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS)
del tr
eng = capi.Engine(gst, rbuf_bytes=256 << 20)
plans = {q: load_query(q, "osdi16_plan")[:3] for q in (4, 5, 6)}
NAMES = {0: "1024thr+warp", 1: "256thr+warp", 2: "256thr block-only (round-2 first cut)", 3: "512thr+warp", -1: "launch per query"}
acc = {}
for rnd in range(a.rounds):
    for var in (0, 1, 2, 3, -1):
        if var >= 0:
            eng.set_resident(True)
            eng.set_option(capi.WK_OPT_RESIDENT_VARIANT, var)
        else:
            eng.set_resident(False)
        for q, (pats, nvars, req) in plans.items():
            for mode in ("cold", "warm"):
                w, _, rows, _ = host.time_query(eng, pats, nvars, req, a.reps, blind=True, flush=(mode == "cold"))
                ns = eng.get_option(capi.WK_INFO_LAST_RESIDENT_NS) if var >= 0 else 0
                acc.setdefault((var, q, mode), []).append((float(np.median(w[1:])), ns / 1e3))
out = {}
for (var, q, mode), v in sorted(acc.items()):
    out.setdefault(NAMES[var], {})["q%d_%s" % (q, mode)] = {"wall_us": round(float(np.median([x[0] for x in v])), 2),
                                                             "in_kernel_us_last": round(float(np.median([x[1] for x in v])), 2)}
print(json.dumps(out, indent=1))
