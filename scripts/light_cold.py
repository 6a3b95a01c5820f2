"""Where do the light queries' microseconds go?  Device time of Q4-Q6 (a) after a full L2 flush, (b) after a flush
followed by the same plan on ANOTHER constant (code + kernel parameters warm, data lines cold), (c) fully warm."""
import argparse
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import sparql_mini as M  # noqa: E402
from conftest import load_query  # noqa: E402
from wukong_b200 import capi, datagen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=2560)
ap.add_argument("--reps", type=int, default=15)
a = ap.parse_args()
tr = datagen.lubm(a.scale, seed=1)
gst = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS)
eng = capi.Engine(gst, rbuf_bytes=256 << 20)
eng.set_profiling(1)
d0 = M.lubm_str2id("<http://www.Department0.University0.edu>")
u0 = M.lubm_str2id("<http://www.University0.edu>")
d_other = M.lubm_str2id("<http://www.Department3.University77.edu>")
u_other = M.lubm_str2id("<http://www.University77.edu>")
out = {}
for q in (4, 5, 6):
    pats, nvars, req, _ = load_query(q, "osdi16_plan")
    other = [tuple(d_other if x == d0 else (u_other if x == u0 else x) for x in p) for p in pats]
    assert other != [tuple(p) for p in pats]
    res = {}
    for mode in ("cold", "code_warm", "warm"):
        us = []
        for _ in range(a.reps):
            if mode != "warm":
                eng.flush_l2()
                eng.sync()
            if mode == "code_warm":
                eng.query(other, nvars, req, blind=True)
            rc, rows, _, _ = eng.query(pats, nvars, req, blind=True)
            assert rc == 0
            us.append(eng.last_query_device_us())
        res[mode] = round(float(np.median(us)), 2)
    res["rows"] = rows
    out["q%d" % q] = res
print(json.dumps(out))
# phase clocks of the fused kernel (profiling level 3), warm and cold
eng.set_profiling(3)
tr_out = {}
for q in (4, 5, 6):
    pats, nvars, req, _ = load_query(q, "osdi16_plan")
    for mode in ("warm", "cold"):
        acc = []
        for _ in range(10):
            if mode == "cold":
                eng.flush_l2(); eng.sync()
            eng.query(pats, nvars, req, blind=True)
            t = eng.light_trace()
            n = len(pats)
            marks = [t[0], t[1]] + [t[2 + s] for s in range(n)] + [t[26], t[27]]
            acc.append(np.diff(np.array(marks, dtype=np.float64)) / 1.965e3)   # us at 1965 MHz
        tr_out["q%d_%s" % (q, mode)] = [round(float(x), 2) for x in np.median(np.array(acc), axis=0)]
print(json.dumps(tr_out))
