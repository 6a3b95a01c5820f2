"""CPU: the oracle against the brute-force joiner, and the host builder against the oracle's build, on seeded random
graphs (skewed degrees, hubs, self loops, duplicates, multi-typed vertices) with random planned patterns."""
import numpy as np
import pytest

import random_bgp as R
import sparql_mini as M
from conftest import rows_equal
from oracle import oracle as O
from wukong_b200 import host


@pytest.mark.parametrize("gseed", range(12))
def test_oracle_vs_bruteforce_on_random_graphs(gseed):
    tr, meta = R.graph(gseed, nv=200, ntriples=1200)
    npreds = meta["num_normal_preds"]
    stores = {n: [O.Store.build(tr, num_servers=n, sid=s, kvstore_bytes=4 << 20, num_engines=2, num_normal_preds=npreds)
                  for s in range(n)] for n in (1, 3)}
    hs = host.HostStore(tr, num_normal_preds=npreds, kvstore_bytes=4 << 20)
    o1 = O.Store.build(tr, kvstore_bytes=4 << 20, num_engines=1, num_normal_preds=npreds)
    assert np.array_equal(hs.vertices(), o1.vertices()) and np.array_equal(hs.edges(), o1.edges())
    assert stores[1][0].check() == 0
    nonempty = 0
    for qseed in range(30):
        planned, semantic, nvars, req = R.query(1000 * gseed + qseed, tr, meta)
        if O.run_query(stores[1], planned, nvars, req, blind=True).rows > 20000:
            continue                      # hub x hub blow-ups: fine for the engines, too slow for the Python joiner
        want = M.bruteforce_bgp(tr, semantic, req)
        for n in (1, 3):
            for mt in (1, 2):
                got = O.run_query(stores[n], planned, nvars, req, mt_factor=mt)
                assert got.status == 0, (gseed, qseed, planned, got.status)
                assert got.rows == want.shape[0], (gseed, qseed, n, mt, planned)
                if got.rows:
                    assert rows_equal(got.table, want), (gseed, qseed, n, mt, planned)
        nonempty += want.shape[0] > 0
        # DISTINCT as final_process defines it (sparql.hpp:1428-1472): rows ordered by ALL columns, then neighbours that
        # agree on the required columns collapse -- so duplicates separated by a differing non-required column survive
        bound = [-(i + 1) for i in range(nvars)]
        full = O.run_query(stores[1], planned, nvars, bound)                # column order = binding order
        d = O.run_query(stores[1], planned, nvars, req, distinct=True, offset=1, limit=50)
        model = M.py_final_process(full.table, [bound.index(v) for v in req], True, 1, 50) if full.rows else np.zeros((0, len(req)))
        assert d.rows == model.shape[0], (gseed, qseed, planned, req)
        if d.rows:
            assert np.array_equal(d.table, model)
            assert d.rows + 1 >= min(51, np.unique(want, axis=0).shape[0])   # never fewer than the true distinct count
    assert nonempty >= 10
