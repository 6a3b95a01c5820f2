"""per-step clocks of in-place light queries on a sharded store (owner rank), cold and warm"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist
from conftest import load_query
from wukong_b200 import capi, datagen
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scale = int(os.environ.get("SCALE", "640"))
gst = capi.Store.build(datagen.lubm_shard(scale, world, rank, seed=1), datagen.LUBM_NUM_NORMAL_PREDS, num_servers=world, sid=rank, device=local)
eng = capi.Engine(gst, rbuf_bytes=1 << 30)
allh = [None] * world
dist.all_gather_object(allh, eng.p2p_export(world, rank)); eng.p2p_import(b"".join(allh))
blobs = [None] * world
dist.all_gather_object(blobs, eng.p2p_export_store()); eng.p2p_import_store(blobs)
dist.barrier()
eng.set_profiling(3)
out = {}
for q in (4, 5, 6):
    pats, nvars, req, _ = load_query(q, "osdi16_plan")
    owner = pats[0][0] % world
    for mode in ("warm", "cold"):
        acc, dev = [], []
        for _ in range(8):
            if mode == "cold":
                eng.flush_l2(); eng.sync()
            dist.barrier()
            rc, rows, _, _ = eng.query_sharded(pats, nvars, req, blind=True)
            assert rc == 0
            dev.append(eng.last_query_device_us())
            if rank == owner:
                t = eng.light_trace()
                n = len(pats)
                marks = [t[0], t[1]] + [t[2 + s] for s in range(n)] + [t[26], t[27]]
                acc.append(np.diff(np.array(marks, dtype=np.float64)) / 1.965e3)
        if rank == owner:
            out["q%d_%s" % (q, mode)] = {"steps_us": [round(float(x), 2) for x in np.median(np.array(acc), axis=0)],
                                         "device_us": round(float(np.median(dev)), 2), "rows": rows}
        elif rank == (owner + 1) % world:
            out["q%d_%s_peer" % (q, mode)] = {"device_us": round(float(np.median(dev)), 2)}
allo = [None] * world
dist.all_gather_object(allo, out)
if rank == 0:
    m = {}
    for o in allo: m.update(o)
    print(json.dumps(m))
dist.barrier(); dist.destroy_process_group()
