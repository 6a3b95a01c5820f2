#!/usr/bin/env python
"""The exchange's push kernel alone on ONE GPU, for ncu: a group of one rank (every row is its own), 16 M rows x 3 columns.
The kernels of a real group wait for each other's flags and cannot run under a profiler's serialisation; with one rank the
whole tile pipeline (cp.async double buffer, ranks, staging, 16-byte stores, per-chunk reservation) still runs, only the
stores stay on the device.  Usage: ncu --set full -k regex:p2p_push -c 2 python scripts/push_ncu.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wukong_b200 import capi, datagen  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
gst = capi.Store.build(datagen.lubm(1, seed=1), datagen.LUBM_NUM_NORMAL_PREDS, device=0)
eng = capi.Engine(gst, rbuf_bytes=1 << 30)
eng.set_resident(False)
capi.local_group([eng])
tbl = np.random.default_rng(3).integers(1 << 17, 1 << 31, (rows, 3), dtype=np.uint32)
eng.set_profiling(2)
for rep in range(3):
    eng.upload(tbl)
    eng.flush_l2()
    n = eng.exchange_p2p(1)
    assert n == rows
    st = [x for x in eng.step_stats() if x["kind"] == "exchange"][-1]
    print("exchange of %d rows x 3 on one rank: %.1f us, %.0f GB/s read + written" % (rows, st["device_us"], 2 * 12 * rows / st["device_us"] / 1e3))
eng.close()
gst.close()
