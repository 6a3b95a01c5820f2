"""Time the device-side store build (wk_store_build) against the host builder on a LUBM-shaped dataset."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from wukong_b200 import capi, datagen, host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=2560)
ap.add_argument("--host", action="store_true", help="also time the host builder")
a = ap.parse_args()
t0 = time.time()
tr = datagen.lubm(a.scale, seed=1)
t_gen = time.time() - t0
out = {"scale": a.scale, "triples": int(tr.shape[0]), "gen_s": round(t_gen, 2)}
for rep in range(2):
    t0 = time.time()
    gs = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS)
    out["device_build_s_rep%d" % rep] = round(time.time() - t0, 3)
    out["device_stats_rep%d" % rep] = {k: (round(v, 1) if isinstance(v, float) else v) for k, v in gs.build_stats.items()}
    if rep == 0:
        gs.close()
eng = capi.Engine(gs, rbuf_bytes=1 << 30)
import glob  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_query  # noqa: E402
rows = {}
for q in range(1, 8):
    pats, nvars, req, _ = load_query(q, "osdi16_plan")
    rc, n, c, _ = eng.query(pats, nvars, req, blind=True)
    rows["q%d" % q] = [rc, n]
out["rows_on_device_built_store"] = rows
if a.host:
    t0 = time.time()
    hs = host.HostStore(tr)
    out["host_build_s"] = round(time.time() - t0, 2)
    out["host_threads"] = os.cpu_count()
print(json.dumps(out))
