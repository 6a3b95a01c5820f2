"""One rank of a sharded query run (spawned by tests/test_sharded.py and usable by hand).

  mode cpu : product shard store (host builder) probed by the ORACLE primitives, rows exchanged with
             torch.distributed all_to_all over gloo, following the product's exchange plan (wk_plan_exchanges)
  mode gpu : the real thing - shard on cuda:<rank>, wk_query_execute_sharded (NCCL all-to-all(v) inside)
Each rank writes its share of every query result to <out>/rank<r>.npz.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run_cpu(rank, world, dist, hs, queries):
    import torch
    from oracle import oracle as O
    from wukong_b200 import capi
    st = O.Store.wrap(hs.vertices(), hs.edges(), hs.segs(), num_servers=world, sid=rank)
    out = {}
    for name, (pats, nvars, req) in queries.items():
        ex = capi.plan_exchanges(pats, nvars)
        v2c = {}
        tbl = np.zeros((0, 0), dtype=np.uint32)
        ncols = 0
        for i, (s, p, d, o) in enumerate(pats):
            if ex[i] != -1:
                # all_to_all(v) of rows over gloo; -2 = replicate to every rank
                if ex[i] == -2:
                    parts = [tbl] * world
                else:
                    dest = tbl[:, ex[i]] % world if tbl.size else np.zeros(0, dtype=np.uint32)
                    parts = [tbl[dest == r] for r in range(world)]
                cnt = torch.tensor([x.shape[0] for x in parts], dtype=torch.int64)
                rcnt = torch.zeros(world, dtype=torch.int64)
                dist.all_to_all_single(rcnt, cnt)
                send = torch.from_numpy(np.concatenate([np.ascontiguousarray(x, dtype=np.int64).reshape(-1) for x in parts]))
                recv = torch.zeros(int(rcnt.sum()) * ncols, dtype=torch.int64)
                dist.all_to_all_single(recv, send, output_split_sizes=[int(x) * ncols for x in rcnt],
                                       input_split_sizes=[int(x) * ncols for x in cnt])
                tbl = recv.numpy().astype(np.uint32).reshape(-1, ncols)
            if i == 0 and 1 < s < (1 << 17):          # index_to_unknown on the local index slice
                tbl = st.primitive(O.I2U, None, 0, s, p, d)
                v2c[o] = 0
                ncols = 1
            elif s >= 0:                                 # const_to_unknown on the owner
                tbl = st.primitive(O.C2U, None, 0, s, p, d) if s % world == rank else np.zeros((0, 1), np.uint32)
                v2c[o] = 0
                ncols = 1
            elif o >= 0:
                tbl = st.primitive(O.K2C, tbl, ncols, v2c[s], p, d, a_end=o)
            elif o in v2c:
                tbl = st.primitive(O.K2K, tbl, ncols, v2c[s], p, d, a_end=v2c[o])
            else:
                tbl = st.primitive(O.K2U, tbl, ncols, v2c[s], p, d)
                v2c[o] = ncols
                ncols += 1
            tbl = tbl.reshape(-1, ncols)
        out[name] = tbl[:, [v2c[v] for v in req]] if tbl.shape[0] else np.zeros((0, len(req)), np.uint32)
    return out


def run_gpu(rank, world, dist, hs, queries, p2p=False, inplace=False):
    import torch
    from wukong_b200 import capi
    gst = hs.upload(rank)
    eng = capi.Engine(gst, rbuf_bytes=128 << 20)
    if p2p:      # peer-memory exchange: gather every rank's IPC handles
        mine = eng.p2p_export(world, rank)
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        eng.p2p_import(b"".join(allh))
        if inplace:   # ... and the peers' store arrays: const-start plans are answered in place by the constant's owner
            blobs = [None] * world
            dist.all_gather_object(blobs, eng.p2p_export_store())
            eng.p2p_import_store(blobs)
        dist.barrier()
    else:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(world, rank, uid[0])
    out = {}
    for name, (pats, nvars, req) in queries.items():
        rc, rows, cols, tbl = eng.query_sharded(pats, nvars, req)
        assert rc == 0, (name, rc)
        out[name] = tbl.copy() if rows else np.zeros((0, len(req)), np.uint32)
        rc, rows_b, _, _ = eng.query_sharded(pats, nvars, req, blind=True)
        assert rc == 0 and rows_b == rows
    out["__stats__"] = np.array(list(eng.comm_stats().values()), dtype=np.uint64)
    eng.close()
    gst.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--port", type=int, required=True)
    ap.add_argument("--mode", default="cpu")
    ap.add_argument("--univs", type=int, default=2)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(a.port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=a.rank, world_size=a.world)
    from conftest import PLANS, load_query
    from wukong_b200 import datagen, host
    tr = datagen.lubm_shard(a.univs, a.world, a.rank, seed=a.seed, chunk=1)
    hs = host.HostStore(tr, num_servers=a.world, sid=a.rank, kvstore_bytes=48 << 20)
    queries = {}
    for q in range(1, 8):
        for plan in PLANS:
            queries["q%d_%s" % (q, plan)] = load_query(q, plan)[:3]
    if a.mode == "gpu":
        res = run_gpu(a.rank, a.world, dist, hs, queries)
    elif a.mode == "gpu_p2p":
        res = run_gpu(a.rank, a.world, dist, hs, queries, p2p=True)
    elif a.mode == "gpu_inplace":
        # a const-start plan whose table outgrows shared memory (members of one department x their courses):
        # the owner's verdict sends every shard through the exchange path instead
        import sparql_mini as M
        from oracle import oracle as O
        P = {n: i for i, n in enumerate(M.LUBM_INDEX)}
        d0 = M.lubm_str2id("<http://www.Department0.University0.edu>")
        queries["spill"] = ([(d0, P[M.UB + "memberOf>"], O.IN, -1), (-1, P[M.UB + "takesCourse>"], O.OUT, -2),
                             (-2, P[M.UB + "teacherOf>"], O.IN, -3)], 3, [-1, -2, -3])
        res = run_gpu(a.rank, a.world, dist, hs, queries, p2p=True, inplace=True)
    else:
        res = run_cpu(a.rank, a.world, dist, hs, queries)
    np.savez(os.path.join(a.out, "rank%d.npz" % a.rank), **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
