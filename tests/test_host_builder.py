"""CPU: the product store builder (csrc/store/host_builder.cpp) against the oracle's restatement of the
reference build, and the data generator / dataset directory format."""
import os

import numpy as np
import pytest

import sparql_mini as M
from conftest import load_query, rows_equal
from oracle import oracle as O
from wukong_b200 import datagen, host


def _segs(s):
    return [(x.index, x.pid, x.dir, x.num_keys, x.num_buckets, x.bucket_start, x.num_edges, x.edge_start,
             x.ext_start, x.ext_num) for x in s]


@pytest.mark.parametrize("nservers,sid", [(1, 0), (2, 0), (2, 1), (3, 2)])
def test_bit_identical_to_oracle_build(lubm2, nservers, sid):
    kv = 48 << 20
    hs = host.HostStore(lubm2, num_servers=nservers, sid=sid, kvstore_bytes=kv)
    os_ = O.Store.build(lubm2, num_servers=nservers, sid=sid, kvstore_bytes=kv, num_engines=1)
    assert _segs(hs.segs()) == _segs(os_.segs())
    assert np.array_equal(hs.edges(), os_.edges())
    assert np.array_equal(hs.vertices(), os_.vertices())


def test_invariants_and_queries_on_product_store(lubm1):
    # auto-sized store (est_load_factor), CPU ext extents, and tiny header forcing long chains
    for kw in (dict(), dict(kvstore_bytes=64 << 20, gpu_ext_extents=False), dict(est_load_factor=90, gpu_ext_extents=False)):
        hs = host.HostStore(lubm1, **kw)
        w = O.Store.wrap(hs.vertices(), hs.edges(), hs.segs())
        assert w.check() == 0                      # gsck invariants
        for q in (1, 4, 7):
            pats, nvars, req, raw = load_query(q, "osdi16_plan")
            r = O.run_query([w], pats, nvars, req)
            assert r.status == 0 and rows_equal(r.table, M.bruteforce_bgp(lubm1, raw, req))
    t = np.unique(lubm1, axis=0)
    hs = host.HostStore(lubm1)
    s, p, o = (int(x) for x in t[12345])
    assert np.array_equal(hs.get_edges(s, p, O.OUT), np.sort(t[(t[:, 0] == s) & (t[:, 1] == p)][:, 2]))
    assert hs.get_edges(4242, 5, O.OUT).size == 0


def test_builder_errors(lubm1):
    with pytest.raises(RuntimeError):
        host.HostStore(lubm1, kvstore_bytes=1 << 20)       # far too small
    bad = lubm1.copy()
    bad[0, 1] = 99                                        # predicate id outside str_index
    with pytest.raises(RuntimeError):
        host.HostStore(bad)


def test_generator_is_deterministic_and_partitionable():
    a = datagen.lubm(3, seed=5)
    b = datagen.lubm(3, seed=5)
    assert np.array_equal(a, b)
    parts = np.concatenate([datagen.lubm(3, seed=5, u_begin=u, u_end=u + 1) for u in range(3)])
    assert np.array_equal(a, parts)
    assert not np.array_equal(a, datagen.lubm(3, seed=6))
    # LUBM Q3 is empty by construction: undergraduates carry no undergraduateDegreeFrom
    ug = set(a[(a[:, 1] == 1) & (a[:, 2] == M.LUBM_INDEX.index(M.UB + "UndergraduateStudent>"))][:, 0].tolist())
    deg = set(a[a[:, 1] == M.LUBM_INDEX.index(M.UB + "undergraduateDegreeFrom>")][:, 0].tolist())
    assert not (ug & deg)
    assert a[:, 0].min() >= (1 << 17) and a[:, 1].max() < 32


def test_dataset_directory_format(tmp_path, lubm1):
    # reference ID-triple directory: id_*.nt "s p o" lines, str_index, str_normal (datagen/README.md)
    d = str(tmp_path / "id_lubm_1")
    n = datagen.lubm_write_dir(d, 1, seed=1)
    assert n == lubm1.shape[0]
    files = sorted(os.listdir(d))
    assert "str_index" in files and "str_normal" in files and "id_uni0.nt" in files
    t = np.loadtxt(os.path.join(d, "id_uni0.nt"), dtype=np.uint32)
    assert np.array_equal(t, lubm1)
    idx = [l.rstrip("\n").split("\t") for l in open(os.path.join(d, "str_index"))]
    assert len(idx) - 1 == datagen.LUBM_NUM_NORMAL_PREDS            # base_loader.hpp:409-424
    assert [s for s, _ in idx] == M.LUBM_INDEX
    normal = dict(l.rstrip("\n").split("\t") for l in open(os.path.join(d, "str_normal")))
    assert int(normal["<http://www.Department0.University0.edu>"]) == M.lubm_str2id("<http://www.Department0.University0.edu>")
    assert int(normal["<http://www.University0.edu>"]) == M.lubm_str2id("<http://www.University0.edu>")
