#!/usr/bin/env python
"""Light-query throughput (reference: sparql-emu with emulator/mix_config, A1-A6 templates,
published 62-73 K q/s on one 24-core node at LUBM-2560): the same mix answered by wk_query_execute_batch
(one launch per batch, one CTA per query) vs the CPU oracle's closed loop on all host threads."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=2560)
ap.add_argument("--queries", type=int, default=1 << 18)
ap.add_argument("--batch", type=int, default=1 << 14)
ap.add_argument("--cpu-queries", type=int, default=1 << 18)
args = ap.parse_args()
import emu_util  # noqa: E402
from oracle import oracle as O  # noqa: E402
from wukong_b200 import capi, datagen, host  # noqa: E402

tr = datagen.lubm(args.scale, seed=1)
gst = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS)      # device-side store build
del tr
eng = capi.Engine(gst, rbuf_bytes=256 << 20)
_v, _e = gst.download()                                         # the same arrays for the CPU oracle arm
ost = O.Store.wrap(_v, _e, gst.segs())
tpl = emu_util.load_templates()
cands = {t[4]: ost.get_edges(0, t[4], 0) for t in tpl}
pats, off, nv, pick = emu_util.instantiate(tpl, cands, args.queries, seed=11)


def batches():
    for b0 in range(0, args.queries, args.batch):
        b1 = min(args.queries, b0 + args.batch)
        yield b0, b1, np.ascontiguousarray(pats[off[b0]:off[b1]]), np.ascontiguousarray(off[b0:b1 + 1] - off[b0]), nv[b0:b1]


# warm-up + correctness of a sample against the oracle
rows_gpu = np.zeros(args.queries, dtype=np.uint64)
for b0, b1, p, o, n in batches():
    r, st = eng.query_batch_raw(p, o, n)
    assert (st == 0).all()
    rows_gpu[b0:b1] = r
nchk = min(args.queries, 1 << 14)
_, want = O.emu_run(ost, np.ascontiguousarray(pats[: off[nchk]]), np.ascontiguousarray(off[: nchk + 1]), nv[:nchk], os.cpu_count())
assert np.array_equal(rows_gpu[:nchk], want), "batched GPU rows differ from the oracle"
best = None
for rep in range(3):
    t0 = time.perf_counter()
    for b0, b1, p, o, n in batches():
        eng.query_batch_raw(p, o, n)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
ncpu = min(args.cpu_queries, args.queries)
threads = os.cpu_count()
cpu_p, cpu_o, cpu_n = np.ascontiguousarray(pats[: off[ncpu]]), np.ascontiguousarray(off[: ncpu + 1]), nv[:ncpu]
O.emu_run(ost, cpu_p, cpu_o, cpu_n, threads)
cpu_best = min(O.emu_run(ost, cpu_p, cpu_o, cpu_n, threads)[0] for _ in range(3))
print(json.dumps({"workload": "LUBM-%d emulator mix A1-A6 (25/25/3/6/25/2), blind" % args.scale, "queries": args.queries,
                  "batch": args.batch, "gpu_qps": round(args.queries / best), "gpu_us_per_batch": round(best / ((args.queries + args.batch - 1) // args.batch) * 1e6, 1),
                  "cpu_oracle_qps": round(ncpu / cpu_best), "cpu_threads": threads, "rows_total": int(rows_gpu.sum()),
                  "includes": "per batch: host plan resolution, H2D of plans (720 B/query), one launch, D2H of row counts"}))
