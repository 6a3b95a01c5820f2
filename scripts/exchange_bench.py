"""NVLink roofline of the fork-join exchange (SURVEY.md §8d "shard exchange" row): R random rows x C columns per rank are
bucketised by row[col] % n and moved to their owners, through the peer-memory push (wk_exchange_p2p) and through NCCL
(wk_exchange).  Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/exchange_bench.py
Achieved GB/s = bytes a GPU sends to its peers (4 * C * rows not kept) / max-over-ranks wall time of the call (the call
ends with the row-count synchronisation, microseconds against milliseconds here); peak = 900 GB/s per direction per GPU."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from wukong_b200 import capi, datagen  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
C_COLS = 3
gst = None
for lf in (55, 45, 35, 25, 15):     # any store will do (the exchange never probes it); a 1/8 shard of LUBM-1 needs a lower load factor
    try:
        gst = capi.Store.build(datagen.lubm_shard(2, world, rank, seed=1), datagen.LUBM_NUM_NORMAL_PREDS, num_servers=world, sid=rank,
                               est_load_factor=lf, device=local)
        break
    except capi.WukongError as ex:
        if ex.code != capi.WK_ERR_STORE_FULL or lf == 15:
            raise
eng = capi.Engine(gst, rbuf_bytes=2 << 30)
allh = [None] * world
dist.all_gather_object(allh, eng.p2p_export(world, rank))
eng.p2p_import(b"".join(allh))
uid = [capi.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(world, rank, uid[0])
dist.barrier()
import argparse  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="1048576,4194304,16777216")
ap.add_argument("--variants", action="store_true", help="also time the push kernel's experiment knobs (WK_P2P_G / _CTAS / _DEBUG)")
args = ap.parse_args()
# (label, environment): the debug modes deliver garbage tables, they only time parts of the kernel
VARIANTS = [("p2p", {}), ("nccl", {})]
if args.variants:
    VARIANTS += [("p2p G=1 (a reservation per tile)", {"WK_P2P_G": "1"}), ("p2p G<=4", {"WK_P2P_G": "4"}), ("p2p G<=64", {"WK_P2P_G": "64"}),
                 ("p2p 3 CTAs/SM", {"WK_P2P_CTAS": "3"}), ("p2p 2 CTAs/SM", {"WK_P2P_CTAS": "2"}),
                 ("p2p no reservations [timing only]", {"WK_P2P_DEBUG": "2", "WK_P2P_G": "1"}),
                 ("p2p no remote stores [timing only]", {"WK_P2P_DEBUG": "1"}),
                 ("p2p neither [timing only]", {"WK_P2P_DEBUG": "3", "WK_P2P_G": "1"})]
out = []
for rows in [int(x) for x in args.sizes.split(",")]:
    rng = np.random.default_rng(1000 + rank)
    tbl = rng.integers(1 << 17, 1 << 31, (rows, C_COLS), dtype=np.uint32)
    kept = int((tbl[:, 1] % world == rank).sum())
    sent_bytes = 4 * C_COLS * (rows - kept)
    for how, env in VARIANTS:
        for k in ("WK_P2P_G", "WK_P2P_CTAS", "WK_P2P_DEBUG"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ts, dev = [], []
        eng.set_profiling(2)
        for rep in range(6):
            eng.upload(tbl)
            eng.flush_l2()
            eng.sync()
            dist.barrier()
            t0 = time.perf_counter()
            n = eng.exchange(1) if how == "nccl" else eng.exchange_p2p(1)
            t1 = time.perf_counter()
            st = [x for x in eng.step_stats() if x["kind"] == "exchange"]
            t = torch.tensor([t1 - t0, (st[-1]["device_us"] if st else 0.0) * 1e-6], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep:
                ts.append(float(t[0].item()))
                dev.append(float(t[1].item()))
        eng.set_profiling(0)
        tot = torch.tensor([n, sent_bytes], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        if "WK_P2P_DEBUG" not in env:
            assert int(tot[0]) == rows * world          # nothing lost, nothing duplicated
        sec, dsec = float(np.median(ts)), float(np.median(dev))
        out.append({"rows_per_gpu": rows, "cols": C_COLS, "exchange": how, "ms": round(sec * 1e3, 3), "device_ms": round(dsec * 1e3, 3),
                    "sent_gb_per_gpu": round(sent_bytes / 1e9, 4), "achieved_gbs_per_gpu": round(sent_bytes / sec / 1e9, 1),
                    "device_gbs_per_gpu": round(sent_bytes / dsec / 1e9, 1) if dsec > 0 else None,
                    "nvlink_peak_gbs": 900.0, "frac": round(sent_bytes / sec / 900e9, 3),
                    "hbm_algo_gbs": round((2 * 4 * C_COLS * rows + 4 * C_COLS * n) / sec / 1e9, 1)})
for k in ("WK_P2P_G", "WK_P2P_CTAS", "WK_P2P_DEBUG"):
    os.environ.pop(k, None)
if rank == 0:
    print(json.dumps({"n_gpus": world, "results": out}))
dist.barrier()
dist.destroy_process_group()
