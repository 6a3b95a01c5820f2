#!/usr/bin/env python
"""bench.py — LUBM Q1-Q7 queries/sec (geomean) on B200, with roofline and CPU baseline.

Contract:
  step     = one pass over the query mix Q1..Q7 (each query executed once)
  value    = geomean over Q1..Q7 of 1 / mean device latency (CUDA events on the engine's stream, store
             and plan resident in HBM, blind mode = the reference's global_silent=1 protocol)
  e2e      = same metric through the public C-ABI call wk_query_execute with HOST buffers: the plan
             goes host->device inside the call, the projected result table comes back device->host
             into pinned memory inside the timed region (non-blind)
  roofline = the dominant kernel (largest share of device time): algorithmic bytes
             / CUDA-event duration vs the measured HBM peak in MEASURED_PEAKS.json
  --impl reference : the CPU oracle (faithful restatement of the reference engine; the reference
             itself cannot be built here) timed on the host cores with the same store arrays.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

HEAVY = (1, 2, 3, 7)
QUERIES = (1, 2, 3, 4, 5, 6, 7)


def geomean(xs):
    xs = [max(float(x), 1e-12) for x in xs]
    return math.exp(sum(math.log(x) for x in xs) / len(xs))


def load_plans(plan):
    from conftest import load_query
    return {q: load_query(q, plan)[:3] for q in QUERIES}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None
        self.t_mark = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons, timed = [], 0, set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            try:
                if self.t_mark is not None and ts >= self.t_mark:
                    timed += 1
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for i, nm in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": timed,
                "note": "sampled every 20 ms from the start of warm-up to the end of the timed region"}


LOAD_FACTORS = (55, 45, 35, 25)   # Global::est_load_factor (global.hpp:99-104) and its fallbacks for small datasets


def build_dataset(args):
    """host arm: triples -> store arrays with the host builder (csrc/store/host_builder.cpp)"""
    from wukong_b200 import datagen, host
    t0 = time.time()
    tr = datagen.lubm(args.scale, seed=args.seed)
    t1 = time.time()
    hs, lf = None, None
    for lf in LOAD_FACTORS:      # a segment that outgrows its single 15 % ext extent (meta.hpp:38-40) needs a sparser header
        try:
            hs = host.HostStore(tr, est_load_factor=lf)
            break
        except RuntimeError:
            if lf == LOAD_FACTORS[-1]:
                raise
    t2 = time.time()
    info = {"triples": int(tr.shape[0]), "keys": int(hs.num_keys), "gen_s": round(t1 - t0, 2), "build_s": round(t2 - t1, 2),
            "store_build": "host", "est_load_factor": lf, "header_mb": round(hs.num_slots * 16 / 1e6, 1), "edges_mb": round(hs.num_edges * 4 / 1e6, 1)}
    return tr, hs, info


class DeviceBuiltStore:
    """the arrays of a device-built store copied back for the CPU baseline (same accessors as host.HostStore)"""

    def __init__(self, gst):
        self._v, self._e = gst.download()
        self._s = gst.segs()

    def vertices(self): return self._v
    def edges(self): return self._e
    def segs(self): return self._s


def build_dataset_device(args, device, num_servers=1, sid=0):
    """triples -> store directly in HBM (wk_store_build: sort / dedup / partition / insert on the GPU)"""
    from wukong_b200 import capi, datagen
    t0 = time.time()
    if num_servers > 1:
        tr = datagen.lubm_shard(args.scale, num_servers, sid, seed=args.seed)
    else:
        tr = datagen.lubm(args.scale, seed=args.seed)
    t1 = time.time()
    gst, lf = None, None
    for lf in LOAD_FACTORS:
        try:
            gst = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS, num_servers=num_servers, sid=sid, est_load_factor=lf, device=device)
            break
        except capi.WukongError as ex:
            if ex.code != capi.WK_ERR_STORE_FULL or lf == LOAD_FACTORS[-1]:
                raise
    t2 = time.time()
    st = gst.build_stats
    info = {"triples": int(tr.shape[0]), "keys": int(st["num_keys"]), "gen_s": round(t1 - t0, 2), "build_s": round(t2 - t1, 2),
            "store_build": "device", "est_load_factor": lf, "build_ms": {k: round(st[k], 1) for k in ("ms_upload", "ms_sort", "ms_insert", "ms_total")},
            "header_mb": round(st["num_slots"] * 16 / 1e6, 1), "edges_mb": round(st["num_edges"] * 4 / 1e6, 1)}
    return tr, gst, info


def cpu_engine_kind(args):
    """which CPU engine times the path: "reference" = the reference's OWN SPARQLEngine (core/engine/sparql.hpp), compiled in
    oracle/_ref behind C shims (oracle/Makefile `ref`), probing the same store arrays through its own GStore code;
    "port" = the oracle restatement (when oracle/_ref was never built, i.e. no reference tree was available)."""
    from oracle import ref as REF
    want = getattr(args, "cpu_engine", "auto")
    if want == "port":
        return "port"
    if REF.available():
        return "reference"
    if want == "reference":
        raise RuntimeError("oracle/_ref/libwukong_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
    return "port"


class CpuEngine:
    """the CPU arm over one set of store arrays (adopted once): per-query mean latency (us); heavy queries as `threads` index
    slices on host threads (the reference's mt_factor replicas, sparql.hpp:1064-1089), light queries single-threaded; the
    timed region is the pattern phase + merge + final_process (projection), non-blind."""

    def __init__(self, hs, kind):
        self.kind = kind
        if kind == "reference":
            from oracle import ref as REF
            self.rs = REF.RefStore.adopt(hs.vertices(), hs.edges(), hs.segs(), num_normal_preds=31)
        else:
            from oracle import oracle as O
            self.ost = O.Store.wrap(hs.vertices(), hs.edges(), hs.segs())

    def times(self, plans, threads, heavy_reps, light_reps):
        out = {}
        for q in QUERIES:
            pats, nvars, req = plans[q]
            heavy = q in HEAVY
            reps = heavy_reps if heavy else light_reps
            mt = threads if heavy else 1
            if self.kind == "reference":
                rc, us, rows = self.rs.time_query(pats, nvars, req, reps=1 + reps, mt_factor=mt, threaded=heavy)
                assert rc == 0, rc
                out[q] = (float(np.mean(us[1:])), int(rows), mt)      # first repetition = warm-up
                continue
            from oracle import oracle as O
            O.run_query([self.ost], pats, nvars, req, mt_factor=mt, blind=False, threaded=heavy)   # warm
            us = []
            for _ in range(reps):
                r = O.run_query([self.ost], pats, nvars, req, mt_factor=mt, blind=False, threaded=heavy)
                assert r.status == 0
                us.append(r.usec)
            out[q] = (float(np.mean(us)), int(r.rows), mt)
        return out


def cpu_engine_times(hs, plans, threads, heavy_reps, light_reps, kind):
    return CpuEngine(hs, kind).times(plans, threads, heavy_reps, light_reps)


CPU_ENGINE_NOTE = {"reference": "the reference's own SPARQLEngine + GStore probe (core/engine/sparql.hpp, core/store/gstore.hpp) compiled "
                                "in oracle/_ref with std-based stand-ins for Boost/TBB/ZeroMQ, over the product-built store arrays",
                   "port": "oracle restatement of the reference engine (oracle/_ref not built)"}


def published_baseline(args):
    """the reference's own published geomean for this exact metric and config (LUBM-2560, Q1-Q7, OSDI16
    fixed plans, 1 node, 2 x 12-core Xeon E5-2650 v4; docs/performance/S1C24-LUBM2560-20181203.md:417-425): 4 253 us -> 235 q/s.
    Other scales / plan sets have no published counterpart."""
    if args.scale == 2560 and args.plan == "osdi16_plan":
        return 235.0
    return None


def vs_published(value, args):
    b = published_baseline(args)
    return (value / b) if b else None


def peak_hbm():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, rank, world):
    if rank != 0:
        return
    tr, hs, info = build_dataset(args)
    plans = load_plans(args.plan)
    threads = args.cpu_threads or os.cpu_count()
    kind = cpu_engine_kind(args)
    cpu = CpuEngine(hs, kind)
    # each step = one bounded pass: heavy queries once with all host threads, light queries 50x
    lat = {q: [] for q in QUERIES}
    t_start = time.time()
    for it in range(args.warmup + args.steps):
        res = cpu.times(plans, threads, 1, 50)
        if it >= args.warmup:
            for q in QUERIES:
                lat[q].append(res[q][0])
    mean = {q: float(np.mean(lat[q])) for q in QUERIES}
    qps = [1e6 / mean[q] for q in QUERIES]
    value = geomean(qps)
    line = {"impl": "reference", "metric": "lubm_q1_q7_geomean_queries_per_sec", "value": value, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sum(mean.values()) / 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": vs_published(value, args),
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "LUBM-%d Q1-Q7 (%s), seeded LUBM-shaped generator" % (args.scale, args.plan),
                       "triples": info["triples"], "non_blind": True},
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": kind,
                             "sample": "per step: Q1,Q2,Q3,Q7 once with mt_factor=%d threads, Q4-Q6 50x single thread; "
                                       "pattern phase + final_process only; engine: %s" % (threads, CPU_ENGINE_NOTE[kind])},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "latency_us": {"q%d" % q: mean[q] for q in QUERIES},
            "wall_s": round(time.time() - t_start, 1)}
    emit(line)


def run_sharded(args, rank, world, local_rank, dist):
    """vid % N sharded store, one exchange (NCCL all-to-all(v)) before every step whose start variable is not local.
    Strong scaling: the dataset is fixed, every query is answered by all ranks together."""
    import torch
    from wukong_b200 import capi, datagen, host
    t0 = time.time()
    if args.store_build == "device":
        tr, gst, _ = build_dataset_device(args, local_rank, num_servers=world, sid=rank)
    else:
        tr = datagen.lubm_shard(args.scale, world, rank, seed=args.seed)
        gst = host.HostStore(tr, num_servers=world, sid=rank).upload(local_rank)
    t1 = time.time()
    rbuf = (args.rbuf_mb << 20) if args.rbuf_mb else max(256 << 20, min(8 << 30, int(tr.shape[0]) * 8))
    eng = capi.Engine(gst, rbuf_bytes=rbuf)
    if args.exchange == "p2p":
        allh = [None] * world
        dist.all_gather_object(allh, eng.p2p_export(world, rank))
        eng.p2p_import(b"".join(allh))
        # peers' store arrays: const-start (light) plans are answered in place by the constant's owner over NVLink loads
        blobs = [None] * world
        dist.all_gather_object(blobs, eng.p2p_export_store())
        eng.p2p_import_store(blobs)
        dist.barrier()
    else:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(world, rank, uid[0])
    plans = load_plans(args.plan)
    sampler = ClockSampler(local_rank)
    sampler.start()
    rows = {}
    for _ in range(args.warmup):
        for q in QUERIES:
            pats, nvars, req = plans[q]
            rc, r, c, _ = eng.query_sharded(pats, nvars, req, blind=True)
            assert rc == 0, rc
            rows[q] = r
    launches0 = eng.launch_count()
    eng.sync(); dist.barrier(); torch.cuda.synchronize()
    t_region0 = time.time()
    sampler.t_mark = t_region0
    eng.set_profiling(1)
    dev_us = {q: [] for q in QUERIES}
    wall_us = {q: [] for q in QUERIES}
    for _ in range(args.steps):
        for q in QUERIES:
            pats, nvars, req = plans[q]
            eng.flush_l2()
            eng.sync()        # the flush must be over on EVERY rank before anyone starts: a rank still flushing would
            dist.barrier()    # make its peers' exchange waits (and their device times) absorb its flush
            w0 = time.perf_counter_ns()
            rc, r, c, _ = eng.query_sharded(pats, nvars, req, blind=True)
            w1 = time.perf_counter_ns()
            assert rc == 0, rc
            dev_us[q].append(eng.last_query_device_us())
            wall_us[q].append((w1 - w0) / 1e3)
    eng.sync(); dist.barrier(); torch.cuda.synchronize()
    t_region = time.time() - t_region0
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    stats = eng.comm_stats()
    t = torch.tensor([np.mean(dev_us[q]) for q in QUERIES] + [np.mean(wall_us[q]) for q in QUERIES], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lat = t.cpu().numpy()
    rr = torch.tensor([rows[q] for q in QUERIES] + [stats["rows_sent"], launches], device="cuda", dtype=torch.int64)
    dist.all_reduce(rr, op=dist.ReduceOp.SUM)
    rr = rr.cpu().numpy()
    if rank == 0:
        dev_mean, wall_mean = lat[:7], lat[7:]
        line = {"metric": "lubm_q1_q7_geomean_queries_per_sec", "value": geomean(1e6 / dev_mean), "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(dev_mean.sum() / 1e3), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": vs_published(geomean(1e6 / dev_mean), args), "dtype": "u32", "data": "synthetic",
                "config": {"workload": "LUBM-%d Q1-Q7 (%s), store sharded by vid %% %d" % (args.scale, args.plan, world),
                           "parallelism": "sharded x%d, %s before non-local steps" % (world, "fused bucketise + peer-memory push over NVLink (CUDA IPC); light plans in place on the owner through peer loads" if args.exchange == "p2p" else "NCCL all-to-all(v)"),
                           "l2": "flushed before every timed query (384 MB memset + 256 MB read-back, outside the timed region)", "value_mode": "blind, device-resident"},
                "e2e": {"value": geomean(1e6 / wall_mean), "unit": "queries/s", "h2d_bytes_per_step": 584, "d2h_bytes_per_step": 56},
                "gpu_launches": int(rr[8]), "clocks": clocks,
                "latency_us": {"device": {"q%d" % q: round(float(dev_mean[i]), 2) for i, q in enumerate(QUERIES)},
                               "wall": {"q%d" % q: round(float(wall_mean[i]), 2) for i, q in enumerate(QUERIES)}},
                "rows": {"q%d" % q: int(rr[i]) for i, q in enumerate(QUERIES)},
                "exchange": {"rows_sent_all_ranks": int(rr[7]), "exchanges_per_rank": stats["exchanges"]},
                "dataset": {"shard_triples_rank0": int(tr.shape[0]), "build_s": round(t1 - t0, 1)}, "timed_region_s": round(t_region, 2)}
        emit(line)
    dist.barrier()
    dist.destroy_process_group()


_RESULT_FD = None


def claim_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the whole run and keep the real stdout for the result line."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="wukong_b200")
    ap.add_argument("--scale", type=int, default=2560, help="number of universities (LUBM-<scale>)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--plan", default="osdi16_plan")
    ap.add_argument("--rbuf-mb", type=int, default=0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-engine", default="auto", choices=["auto", "reference", "port"],
                    help="CPU arm: the reference's own engine compiled in oracle/_ref, or the oracle port (auto: reference when built)")
    ap.add_argument("--store-build", default="device", choices=["device", "host"],
                    help="build the graph store on the GPU (wk_store_build) or with the host builder + upload")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="sharded mode: peer-memory push or NCCL all-to-all(v)")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="N>1: replicas (whole store per GPU, weak scaling) or sharded (vid %% N + NCCL all-to-all)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:   # torchrun pins OMP_NUM_THREADS=1; the host-side store build is OpenMP-parallel
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    from wukong_b200 import capi, host
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        if rank != 0:
            ge.build()   # libraries exist by now; loads them
    if capi.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")

    if world > 1 and args.mode == "sharded":
        run_sharded(args, rank, world, local_rank, dist)
        return
    hs = None
    if args.store_build == "device":
        tr, gst, info = build_dataset_device(args, local_rank)
    else:
        tr, hs, info = build_dataset(args)
        gst = hs.upload(local_rank)
    plans = load_plans(args.plan)
    rbuf = (args.rbuf_mb << 20) if args.rbuf_mb else max(256 << 20, min(8 << 30, int(info["triples"]) * 8))
    eng = capi.Engine(gst, rbuf_bytes=rbuf)
    out_tbl, _keep = capi.pinned_array(min(rbuf // 4, 1 << 28))

    def barrier():
        eng.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ---- warm-up ---------------------------------------------------------------------------------
    rows = {}
    for _ in range(args.warmup):
        for q in QUERIES:
            pats, nvars, req = plans[q]
            _, _, r, c = host.time_query(eng, pats, nvars, req, 1, blind=False, table=out_tbl, flush=True)
            rows[q] = (r, c)
    launches0 = eng.launch_count()
    barrier()
    t_region0 = time.time()
    sampler.t_mark = t_region0
    # ---- timed: device-resident (value) and end-to-end (e2e), K steps, L2 flushed between queries ----
    dev_us = {q: [] for q in QUERIES}
    e2e_us = {q: [] for q in QUERIES}
    eng.set_profiling(1)
    for _ in range(args.steps):
        for q in QUERIES:
            pats, nvars, req = plans[q]
            _, d, _, _ = host.time_query(eng, pats, nvars, req, 1, blind=True, flush=True, device_times=True)
            dev_us[q].append(float(d[0]))
    eng.set_profiling(0)
    for _ in range(args.steps):
        for q in QUERIES:
            pats, nvars, req = plans[q]
            w, _, _, _ = host.time_query(eng, pats, nvars, req, 1, blind=False, table=out_tbl, flush=True)
            e2e_us[q].append(float(w[0]))
    barrier()
    t_region = time.time() - t_region0
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0

    # ---- roofline of the dominant kernel (per-step CUDA events; separate pass) ---------------------------
    eng.set_profiling(2)
    agg = {}
    for _ in range(max(3, min(args.steps, 10))):
        for q in QUERIES:
            pats, nvars, req = plans[q]
            eng.flush_l2()
            rc, _, _, _ = eng.query(pats, nvars, req, blind=True)
            assert rc == 0
            for i, s in enumerate(eng.step_stats()):
                a = agg.setdefault((q, i, s["kind"]), {"us": [], "bytes": s["algo_bytes"], "in_rows": s["in_rows"],
                                                       "out_rows": s["out_rows"], "in_cols": s["in_cols"]})
                a["us"].append(s["device_us"])
    eng.set_profiling(0)
    kern = [(k, v) for k, v in agg.items() if k[2] in ("k2u", "k2k", "k2c") and np.mean(v["us"]) > 0]
    roof = None
    roof_expand = None
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from this round's ncu --set full capture
    # of the same launches (profiles/README.md, LUBM-2560, osdi16 plan); only valid for that workload
    NCU_TRAFFIC = {(1, 1): 881.1e6, (1, 2): 779.7e6, (1, 3): 111.9e6, (1, 4): 80.9e6} if (args.scale == 2560 and args.plan == "osdi16_plan") else {}
    steps_table = []
    for k, v in sorted(agg.items()):
        us = float(np.mean(v["us"])) if v["us"] else 0.0
        steps_table.append({"q": k[0], "step": k[1], "kind": k[2], "in_rows": int(v["in_rows"]), "out_rows": int(v["out_rows"]),
                            "algo_bytes": int(v["bytes"]), "device_us": round(us, 2),
                            "gbs": round(v["bytes"] / us / 1e3, 1) if us > 0 else None})
    if kern:
        (kq, ki, kk), v = max(kern, key=lambda kv: float(np.mean(kv[1]["us"])))
        peak, peak_src = peak_hbm()

        def roof_of(kq, ki, kk, v):
            us = float(np.mean(v["us"]))
            ach = v["bytes"] / us / 1e3   # GB/s
            tr_ = NCU_TRAFFIC.get((kq, ki))
            return {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                    "traffic": int(tr_) if tr_ else None,
                    "kernel": "step_kernel_v5<%s> (q%d step %d: %d rows x %d cols -> %d rows)" % (kk, kq, ki, v["in_rows"], v["in_cols"], v["out_rows"]),
                    "algo_bytes_per_launch": int(v["bytes"]), "us_per_launch": round(us, 2), "peak_source": peak_src}
        roof = roof_of(kq, ki, kk, v)
        if roof["traffic"] is not None and roof["traffic"] < 0.5 * roof["algo_bytes_per_launch"]:
            roof["note"] = "this step probes a few thousand hot keys: most algorithmic bytes are L2 hits, not DRAM traffic"
        # the expand (known_to_unknown) launch with the largest device time: north_star's "expand-kernel HBM GB/s"
        k2u = [(k, v2) for k, v2 in kern if k[2] == "k2u"]
        if k2u:
            (eq, ei, ek), ev = max(k2u, key=lambda kv: float(np.mean(kv[1]["us"])))
            roof_expand = roof_of(eq, ei, ek, ev)

    # ---- reduce over ranks (max latency), compute the metric ----------------------------------------------------
    dev_mean = np.array([np.mean(dev_us[q]) for q in QUERIES])
    e2e_mean = np.array([np.mean(e2e_us[q]) for q in QUERIES])
    if dist is not None:
        import torch
        t = torch.tensor(np.concatenate([dev_mean, e2e_mean]), device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        arr = t.cpu().numpy()
        dev_mean, e2e_mean = arr[:7], arr[7:]
    # replicas: every rank answers its own stream of queries => whole-job rate = world x per-replica rate
    value = geomean(world * 1e6 / dev_mean)
    e2e = geomean(world * 1e6 / e2e_mean)
    d2h = sum(rows[q][0] * rows[q][1] * 4 + 8 for q in QUERIES)
    h2d = sum(len(plans[q][0]) * 16 + len(plans[q][2]) * 4 for q in QUERIES)

    line = {"metric": "lubm_q1_q7_geomean_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(dev_mean.sum() / 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": vs_published(value, args), "dtype": "u32", "data": "synthetic",
            "config": {"workload": "LUBM-%d Q1-Q7 (%s), seeded LUBM-shaped generator" % (args.scale, args.plan),
                       "triples": info["triples"], "keys": info["keys"], "store_mb": info["header_mb"] + info["edges_mb"],
                       "parallelism": "replicas x%d" % world if world > 1 else "single GPU",
                       "l2": "flushed before every timed query (384 MB memset + 256 MB read-back, outside the timed region)",
                       "value_mode": "blind (row count only), device-resident", "e2e_mode": "non-blind, table D2H into pinned memory"},
            "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_expand": roof_expand,
            "latency_us": {"device": {"q%d" % q: round(float(dev_mean[i]), 2) for i, q in enumerate(QUERIES)},
                           "e2e": {"q%d" % q: round(float(e2e_mean[i]), 2) for i, q in enumerate(QUERIES)}},
            "rows": {"q%d" % q: int(rows[q][0]) for q in QUERIES}, "steps_table": steps_table,
            "dataset": info, "timed_region_s": round(t_region, 2)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = args.cpu_threads or os.cpu_count()
        kind = cpu_engine_kind(args)
        res = cpu_engine_times(hs if hs is not None else DeviceBuiltStore(gst), plans, threads, 3, 200, kind)
        cq = [1e6 / res[q][0] for q in QUERIES]
        for q in QUERIES:
            assert res[q][1] == rows[q][0], "GPU and CPU engine disagree on q%d rows" % q
        line["cpu_baseline"] = {"value": geomean(cq), "unit": "queries/s", "cores": threads, "kind": kind,
                                "sample": "same store arrays: Q1,Q2,Q3,Q7 3x with mt_factor=%d threads, Q4-Q6 200x single thread; "
                                          "pattern phase + final_process; engine: %s" % (threads, CPU_ENGINE_NOTE[kind]),
                                "latency_us": {"q%d" % q: round(res[q][0], 2) for q in QUERIES}}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
