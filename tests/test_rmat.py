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
