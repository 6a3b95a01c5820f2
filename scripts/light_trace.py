"""Where do the light queries' microseconds go?  Q4-Q6 through the resident server (and the launch-per-query kernel) with
the L2 flushed before every query: host wall clock, in-kernel span, and the phase clocks of the interpreter
(profiling level 3; layout in wk_light.cuh)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from conftest import load_query  # noqa: E402
from wukong_b200 import capi, datagen, host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=2560)
ap.add_argument("--reps", type=int, default=15)
ap.add_argument("--mhz", type=float, default=1965.0)
a = ap.parse_args()
tr = datagen.lubm(a.scale, seed=1)
gst = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS)
del tr
eng = capi.Engine(gst, rbuf_bytes=256 << 20)
out = {}
for resident in (True, False):
    eng.set_resident(resident)
    for q in (4, 5, 6):
        pats, nvars, req, _ = load_query(q, "osdi16_plan")
        n = len(pats)
        for mode in ("cold", "warm"):
            eng.set_profiling(0)
            w, _, rows, _ = host.time_query(eng, pats, nvars, req, a.reps, blind=True, flush=(mode == "cold"))
            res = {"rows": rows, "wall_us": round(float(np.median(w)), 2)}
            if resident:
                ns = []
                for _ in range(a.reps):
                    if mode == "cold":
                        eng.flush_l2()
                    eng.query(pats, nvars, req, blind=True)
                    ns.append(eng.get_option(capi.WK_INFO_LAST_RESIDENT_NS))
                res["in_kernel_us"] = round(float(np.median(ns)) / 1e3, 2)
            eng.set_profiling(3)
            acc, fine = [], []
            for _ in range(a.reps):
                if mode == "cold":
                    eng.flush_l2()
                eng.query(pats, nvars, req, blind=True)
                if not resident:
                    eng.sync()
                t = eng.light_trace().astype(np.float64)
                t0 = t[28] if resident else t[0]
                marks = [t0] + ([t[29]] if resident else [t[1]]) + [t[2 + s] for s in range(n)] + [t[26], t[27]]
                acc.append(np.diff(np.array(marks)) / a.mhz)
                fine.append([[(t[32 + 4 * s + k] - (t[2 + s - 1] if s else (t[29] if resident else t[1]))) / a.mhz for k in range(3)] for s in range(1, n)])
            res["phases_us"] = {"names": ["acquire/entry->decoded"] + ["step%d" % s for s in range(n)] + ["project", "record"],
                                "median": [round(float(x), 2) for x in np.median(np.array(acc), axis=0)]}
            res["inside_steps_us(after_probe,after_scan,after_write)"] = [[round(float(x), 2) for x in row] for row in np.median(np.array(fine), axis=0)] if n > 1 else []
            out["q%d_%s_%s" % (q, "resident" if resident else "launch", mode)] = res
eng.set_profiling(0)
print(json.dumps(out))
