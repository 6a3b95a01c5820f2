"""Power-law (R-MAT) graphs: skewed degrees, hubs with thousands of edges, duplicates in the input."""
import numpy as np
import pytest

import sparql_mini as M
from conftest import rows_equal
from oracle import oracle as O
from wukong_b200 import datagen, host

P = datagen.RMAT_PRED
TWO_HOP = [(P, O.PREDICATE_ID, O.IN, -1), (-1, P, O.OUT, -2), (-2, P, O.OUT, -3)]


@pytest.fixture(scope="module")
def rmat_small():
    return datagen.rmat(10, 6000, seed=3)


@pytest.fixture(scope="module")
def rmat_mid():
    return datagen.rmat(14, 200000, seed=42)


def test_rmat_oracle_vs_bruteforce(rmat_small):
    st = O.Store.build(rmat_small, kvstore_bytes=16 << 20, num_normal_preds=datagen.RMAT_NUM_NORMAL_PREDS)
    assert st.check() == 0
    hs = host.HostStore(rmat_small, num_normal_preds=datagen.RMAT_NUM_NORMAL_PREDS, kvstore_bytes=16 << 20)
    assert np.array_equal(hs.vertices(), st.vertices()) and np.array_equal(hs.edges(), st.edges())
    raw = [(-1, P, O.OUT, -2), (-2, P, O.OUT, -3)]
    bf = M.bruteforce_bgp(rmat_small, raw, [-1, -2, -3])
    r = O.run_query([st], TWO_HOP, 3, [-1, -2, -3], mt_factor=3)
    assert r.status == 0 and r.rows == bf.shape[0] and rows_equal(r.table, bf)


@pytest.mark.gpu
def test_rmat_two_hop_gpu(rmat_mid):
    from wukong_b200 import capi
    hs = host.HostStore(rmat_mid, num_normal_preds=datagen.RMAT_NUM_NORMAL_PREDS)
    ost = O.Store.wrap(hs.vertices(), hs.edges(), hs.segs())
    gst = hs.upload(0)
    eng = capi.Engine(gst, rbuf_bytes=1 << 30)
    want = O.run_query([ost], TWO_HOP, 3, [-1, -2, -3])
    assert want.rows > 1_000_000                     # heavy fan-out through the hubs
    rc, rows, cols, tbl = eng.query(TWO_HOP, 3, [-1, -2, -3], out=np.empty(want.rows * 3 + 16, dtype=np.uint32))
    assert rc == 0 and rows == want.rows
    assert rows_equal(tbl, want.table)
    # filters over hub lists: keep (a, b) pairs where b also points back to a
    pats = [(P, O.PREDICATE_ID, O.IN, -1), (-1, P, O.OUT, -2), (-2, P, O.OUT, -1)]
    want = O.run_query([ost], pats, 2, [-1, -2])
    rc, rows, cols, tbl = eng.query(pats, 2, [-1, -2])
    assert rc == 0 and rows == want.rows and rows_equal(tbl, want.table)
    # type check of every reached vertex
    pats = [(P, O.PREDICATE_ID, O.OUT, -1), (-1, 1, O.OUT, datagen.RMAT_TYPE), (-1, P, O.IN, -2)]
    want = O.run_query([ost], pats, 2, [-2, -1])
    rc, rows, cols, tbl = eng.query(pats, 2, [-2, -1])
    assert rc == 0 and rows == want.rows and rows_equal(tbl, want.table)
    eng.close()
    gst.close()


@pytest.mark.gpu
def test_rmat_device_built_store(rmat_mid):
    """the device-side store build on a skewed graph (hubs with thousands of edges): same two-hop answer as the oracle"""
    from wukong_b200 import capi
    hs = host.HostStore(rmat_mid, num_normal_preds=datagen.RMAT_NUM_NORMAL_PREDS)
    ost = O.Store.wrap(hs.vertices(), hs.edges(), hs.segs())
    gst = capi.Store.build(rmat_mid, datagen.RMAT_NUM_NORMAL_PREDS)
    eng = capi.Engine(gst, rbuf_bytes=1 << 30)
    want = O.run_query([ost], TWO_HOP, 3, [-1, -2, -3])
    rc, rows, cols, tbl = eng.query(TWO_HOP, 3, [-1, -2, -3], out=np.empty(want.rows * 3 + 16, dtype=np.uint32))
    assert rc == 0 and rows == want.rows and rows_equal(tbl, want.table)
    subj = gst.get_edges(0, P, O.IN)
    assert np.array_equal(np.sort(subj), np.sort(hs.get_edges(0, P, O.IN)))
    eng.close()
    gst.close()


def test_scramble_is_a_relabelling_and_streams_do_not_depend_on_threads():
    """CPU: the Graph500-style scramble permutes vertex labels (same degree multiset as the raw R-MAT graph, no longer
    readable from the low id bits), and the edge list is a function of (scale, edges, seed) only -- ranks of a sharded run
    generate it with different thread counts and must agree"""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    scale, ne = 12, 200000
    raw = datagen.rmat(scale, ne, seed=5, typed=False, scramble=False)
    scr = datagen.rmat(scale, ne, seed=5, typed=False, scramble=True)
    assert raw.shape == scr.shape == (ne, 3)
    base = 1 << 17
    assert scr[:, [0, 2]].min() >= base and scr[:, [0, 2]].max() < base + (1 << scale)
    for c in (0, 2):
        d_raw = np.sort(np.unique(raw[:, c], return_counts=True)[1])
        d_scr = np.sort(np.unique(scr[:, c], return_counts=True)[1])
        assert np.array_equal(d_raw, d_scr)
    # the same edge keeps its place in the list, so the relabelling is one map for subjects and objects alike
    m = {}
    for a, b in zip(raw[:, 0].tolist() + raw[:, 2].tolist(), scr[:, 0].tolist() + scr[:, 2].tolist()):
        assert m.setdefault(a, b) == b
    assert len(set(m.values())) == len(m)
    # raw ids: the out-degree mass of ids = 0 mod 8 is several times 1/8; scrambled: close to 1/8
    share_raw = float((raw[:, 0] % 8 == 0).mean())
    share_scr = float((scr[:, 0] % 8 == 0).mean())
    assert share_raw > 0.3 and abs(share_scr - 0.125) < 0.05, (share_raw, share_scr)
    code = ("import sys, hashlib; sys.path.insert(0, %r); from wukong_b200 import datagen; "
            "print(hashlib.sha256(datagen.rmat(%d, %d, seed=5, typed=False).tobytes()).hexdigest())" % (ROOT, scale, ne))
    outs = set()
    for nt in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=nt)
        outs.add(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300).stdout.strip())
    assert len(outs) == 1 and len(next(iter(outs))) == 64, outs
