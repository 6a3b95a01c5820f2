"""GPU: the device-side store builder (wk_store_build, csrc/kernels/store_build.cu) against the host builder
(bit-identical to the oracle's restatement of the reference build) and against the oracle's engine."""
import numpy as np
import pytest

from conftest import PLANS, load_query, rows_equal
from oracle import oracle as O
from wukong_b200 import capi, datagen, host

pytestmark = pytest.mark.gpu
NP = datagen.LUBM_NUM_NORMAL_PREDS


def _segs(s):
    return [(x.index, x.pid, x.dir, x.num_keys, x.num_buckets, x.bucket_start, x.num_edges, x.edge_start,
             x.ext_start, x.ext_num) for x in s]


def _entries(v):
    """the (key, ptr) pairs a probe can find: every non-empty slot except the chain-pointer slot of a bucket"""
    v = np.asarray(v).reshape(-1, 2)
    keep = (np.arange(v.shape[0]) % 8 != 7) & (v[:, 0] != 0)
    e = v[keep]
    return e[np.lexsort((e[:, 1], e[:, 0]))]


def _same_store(gs, hs, nservers=1, sid=0):
    assert _segs(gs.segs()) == _segs(hs.segs())
    v, e = gs.download()
    assert np.array_equal(e, hs.edges())                       # sorted runs in the same layout
    assert np.array_equal(_entries(v), _entries(hs.vertices()))
    w = O.Store.wrap(v, e, gs.segs(), num_servers=nservers, sid=sid)
    assert w.check() == 0                                      # gsck: every key reachable, index <-> normal consistent
    return w


@pytest.mark.parametrize("nservers,sid,kv", [(1, 0, 48 << 20), (1, 0, 0), (2, 1, 32 << 20), (3, 2, 0)])
def test_same_store_as_host_builder(lubm2, nservers, sid, kv):
    gs = capi.Store.build(lubm2, NP, num_servers=nservers, sid=sid, kvstore_bytes=kv)
    hs = host.HostStore(lubm2, num_servers=nservers, sid=sid, kvstore_bytes=kv)
    _same_store(gs, hs, nservers, sid)
    st = gs.build_stats
    assert st["num_keys"] == hs.num_keys and st["num_slots"] == hs.num_slots
    gs.close()


def test_duplicates_shuffle_and_chains(lubm1):
    # duplicated and shuffled input gives the same store; a high load factor forces multi-bucket chains
    rng = np.random.default_rng(3)
    t = np.concatenate([lubm1, lubm1[rng.integers(0, lubm1.shape[0], 5000)]])
    t = t[rng.permutation(t.shape[0])]
    for lf in (55, 62, 80):
        try:
            hs = host.HostStore(lubm1, est_load_factor=lf)
        except RuntimeError:
            # a segment outgrew its single 15% ext extent (meta.hpp:38-40): both builders must refuse
            with pytest.raises(RuntimeError):
                capi.Store.build(t, NP, est_load_factor=lf)
            continue
        gs = capi.Store.build(t, NP, est_load_factor=lf)
        _same_store(gs, hs)
        assert gs.build_stats["used_ext"] > 0
        gs.close()


def test_queries_on_device_built_store(lubm2, ostore2):
    gs = capi.Store.build(lubm2, NP)
    eng = capi.Engine(gs, rbuf_bytes=64 << 20)
    for q in range(1, 8):
        for plan in PLANS[:1]:
            pats, nvars, req, _ = load_query(q, plan)
            want = O.run_query([ostore2], pats, nvars, req, mt_factor=1)
            rc, rows, cols, tbl = eng.query(pats, nvars, req)
            assert rc == 0 and rows == want.rows and rows_equal(tbl, want.table), q
    # host-side probe through the device arrays
    s, p, o = (int(x) for x in lubm2[777])
    assert np.array_equal(gs.get_edges(s, p, O.OUT), ostore2.get_edges(s, p, O.OUT))
    assert np.array_equal(np.sort(gs.get_edges(0, p, O.IN)), np.sort(ostore2.get_edges(0, p, O.IN)))
    eng.close()
    gs.close()


def test_build_errors(lubm1):
    bad = lubm1.copy()
    bad[5, 1] = 99
    with pytest.raises(RuntimeError):
        capi.Store.build(bad, NP)                              # predicate id outside str_index
    with pytest.raises(RuntimeError):
        capi.Store.build(lubm1, NP, kvstore_bytes=1 << 20)     # far too small
    with pytest.raises(RuntimeError):
        capi.Store.build(lubm1, NP, num_servers=2, sid=2)
    # an empty shard is a valid (empty) store
    gs = capi.Store.build(np.zeros((0, 3), dtype=np.uint32), NP)
    assert gs.build_stats["num_triples_out"] == 0
    gs.close()
