"""Sharded (vid % n) execution: world_size-2/3 gloo run on CPU for the host-side logic, NCCL run on GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import sparql_mini as M
from conftest import PLANS, ROOT, load_query, rows_equal
from wukong_b200 import capi, datagen


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_world(world, mode, tmp_path, univs=2, seed=7):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "sharded_worker.py"), "--rank", str(r), "--world",
                               str(world), "--port", str(port), "--mode", mode, "--univs", str(univs), "--seed", str(seed),
                               "--out", str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(world)]
    outs = [p.communicate(timeout=900)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]


def _check_against_bruteforce(results, univs, seed):
    tr = datagen.lubm(univs, seed=seed)
    for q in range(1, 8):
        _, _, req, raw = load_query(q, PLANS[0])
        bf = M.bruteforce_bgp(tr, raw, req)
        for plan in PLANS:
            name = "q%d_%s" % (q, plan)
            got = np.concatenate([r[name].reshape(-1, len(req)) for r in results])
            assert rows_equal(got, bf), name


def test_exchange_plan():
    # Q7, optimal10240 plan: SURVEY appendix walk-through: 3 exchanges (by ?X, ?Y, ?Z)
    pats, nvars, req, _ = load_query(7, "optimal10240_plan")
    ex = capi.plan_exchanges(pats, nvars)
    assert [e for e in ex if e != -1] == [1, 0, 2] and ex[0] == -1 and ex[1] == -1 and ex[2] == -1
    # Q2: type-index seed then ?X name ?Y on the same (local) variable: no exchange at all
    pats, nvars, req, _ = load_query(2, "osdi16_plan")
    assert capi.plan_exchanges(pats, nvars) == [-1, -1]
    # a type-index lookup of a known variable must be replicated to every shard (sparql.hpp:1091-1110)
    assert capi.plan_exchanges([(18, 1, 0, -1), (-1, 5, 0, -2), (-2, 1, 0, -3)], 3)[2] == -2


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_cpu_gloo(world, tmp_path):
    """product shard builder + product exchange plan + gloo all_to_all, probed with the oracle primitives"""
    res = _run_world(world, "cpu", tmp_path)
    _check_against_bruteforce(res, 2, 7)


@pytest.mark.gpu
def test_partition_counts_single_gpu(lubm1):
    from wukong_b200 import host
    gst = host.HostStore(lubm1).upload(0)
    eng = capi.Engine(gst, rbuf_bytes=64 << 20)
    rng = np.random.default_rng(0)
    tbl = rng.integers(1 << 17, 1 << 24, (100003, 3), dtype=np.uint32)
    for nparts, col in ((2, 0), (8, 2), (5, 1)):
        eng.upload(tbl)
        got = eng.partition(col, nparts)
        assert np.array_equal(got, np.bincount(tbl[:, col] % nparts, minlength=nparts).astype(np.uint64))
    eng.close()
    gst.close()


@pytest.mark.gpu
def test_sharded_gpu_nccl(tmp_path):
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(capi.device_count(), 4)
    res = _run_world(world, "gpu", tmp_path)
    _check_against_bruteforce(res, 2, 7)
    st = res[0]["__stats__"]
    assert st[0] > 0          # exchanges happened


@pytest.mark.gpu
def test_sharded_gpu_peer_memory(tmp_path):
    """same queries, exchange through CUDA-IPC peer memory (kernel stores over NVLink, no NCCL, no host sync)"""
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(capi.device_count(), 4)
    res = _run_world(world, "gpu_p2p", tmp_path)
    _check_against_bruteforce(res, 2, 7)
    assert res[0]["__stats__"][0] > 0
    tot = sum(np.asarray(r["__stats__"], dtype=np.uint64) for r in res)      # exchanges, rows sent, rows received
    assert tot[1] == tot[2] and tot[1] > 0                                    # every row pushed to a peer arrived


@pytest.mark.gpu
def test_sharded_gpu_in_place_light_queries(tmp_path):
    """const-start plans answered by the constant's owner alone, probing the other shards through peer memory (the
    reference's in-place mode with one-sided reads); a table that outgrows shared memory falls back to the exchange path"""
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(capi.device_count(), 4)
    res = _run_world(world, "gpu_inplace", tmp_path)
    _check_against_bruteforce(res, 2, 7)
    # light queries no longer exchange: fewer exchanges than the all-collective run of the same workload would need,
    # and every row of a light answer sits on ONE rank (the owner of Department0 / University0)
    for q in (4, 5, 6):
        name = "q%d_%s" % (q, PLANS[0])
        holders = [r for r in range(world) if res[r][name].size]
        assert len(holders) <= 1, (name, holders)
    tr = datagen.lubm(2, seed=7)
    d0 = M.lubm_str2id("<http://www.Department0.University0.edu>")
    P = {n: i for i, n in enumerate(M.LUBM_INDEX)}
    t = np.unique(tr, axis=0)
    mem = t[(t[:, 1] == P[M.UB + "memberOf>"]) & (t[:, 2] == d0)][:, 0]
    tc = t[(t[:, 1] == P[M.UB + "takesCourse>"]) & np.isin(t[:, 0], mem)]
    to = t[t[:, 1] == P[M.UB + "teacherOf>"]]
    want = np.array([(x, c, y) for x, _, c in tc.tolist() for y, _, c2 in to[to[:, 2] == c].tolist()], dtype=np.uint32).reshape(-1, 3)
    got = np.concatenate([r["spill"].reshape(-1, 3) for r in res])
    assert want.shape[0] > 1024 and rows_equal(got, want)
