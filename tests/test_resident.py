"""The resident light-query server (wk_server.cuh): const-start plans answered through a doorbell in mapped memory by a
kernel that stays on the device, against the launch-per-query path and the oracle; lifetime rules (idle exit, parking
around grid-filling kernels, destroy while resident); a table that outgrows shared memory continues on the multi-CTA path."""
import threading
import time

import numpy as np
import pytest

import sparql_mini as M
from conftest import PLANS, load_query, rows_equal
from oracle import oracle as O
from wukong_b200 import capi

pytestmark = pytest.mark.gpu

P = {n: i for i, n in enumerate(M.LUBM_INDEX)}
def pid(name): return P[M.UB + name + ">"]


def _spill_query():
    d0 = M.lubm_str2id("<http://www.Department0.University0.edu>")
    return ([(d0, pid("memberOf"), O.IN, -1), (-1, pid("takesCourse"), O.OUT, -2), (-2, pid("teacherOf"), O.IN, -3)], 3, [-1, -2, -3])


def test_resident_matches_launch_path_and_oracle(gstore2, ostore2):
    eng = capi.Engine(gstore2, rbuf_bytes=64 << 20)
    assert eng.get_option(capi.WK_OPT_RESIDENT_LIGHT) == 1
    for q in (4, 5, 6):
        for plan in PLANS:
            pats, nvars, req, _ = load_query(q, plan)
            want = O.run_query([ostore2], pats, nvars, req)
            eng.set_resident(True)
            l0, r0 = eng.launch_count(), eng.get_option(capi.WK_INFO_RESIDENT_REQUESTS)
            rc, rows, cols, tbl = eng.query(pats, nvars, req)
            assert rc == 0 and rows == want.rows and rows_equal(tbl, want.table), (q, plan)
            assert eng.get_option(capi.WK_INFO_LAST_RESIDENT) == 1
            assert eng.get_option(capi.WK_INFO_RESIDENT_REQUESTS) == r0 + 1
            assert eng.get_option(capi.WK_INFO_LAST_RESIDENT_NS) > 0
            rc, rows_b, _, _ = eng.query(pats, nvars, req, blind=True)
            assert rc == 0 and rows_b == want.rows
            # no kernel launch per query once the server is up (at most one launch: the server itself)
            assert eng.launch_count() - l0 <= 1
            eng.set_resident(False)
            rc, rows, cols, tbl2 = eng.query(pats, nvars, req)
            assert rc == 0 and rows == want.rows and rows_equal(tbl2, want.table)
            assert eng.get_option(capi.WK_INFO_LAST_RESIDENT) == 0
    eng.close()


def test_resident_lifetime(gstore2, ostore2):
    eng = capi.Engine(gstore2, rbuf_bytes=64 << 20)
    eng.set_option(capi.WK_OPT_RESIDENT_IDLE_US, 300)
    pats, nvars, req, _ = load_query(5, PLANS[0])
    want = O.run_query([ostore2], pats, nvars, req)
    for _ in range(3):
        rc, rows, _, tbl = eng.query(pats, nvars, req)
        assert rc == 0 and rows_equal(tbl, want.table)
    n0 = eng.get_option(capi.WK_INFO_RESIDENT_LAUNCHES)
    assert n0 >= 1
    time.sleep(0.05)          # far beyond the idle limit: the server has left on its own ...
    assert eng.get_option(capi.WK_INFO_RESIDENT_RUNNING) == 0
    eng.sync()
    rc, rows, _, tbl = eng.query(pats, nvars, req)   # ... and comes back on demand
    assert rc == 0 and rows_equal(tbl, want.table)
    assert eng.get_option(capi.WK_INFO_RESIDENT_LAUNCHES) == n0 + 1
    # heavy queries park the server and bring it back; light queries in between stay correct
    eng.set_option(capi.WK_OPT_RESIDENT_IDLE_US, 10000)
    for q in (5, 2, 4, 1, 6, 7, 5, 3, 4):
        pq, nv, rq, _ = load_query(q, PLANS[0])
        w = O.run_query([ostore2], pq, nv, rq)
        rc, rows, _, tbl = eng.query(pq, nv, rq)
        assert rc == 0 and rows == w.rows and rows_equal(tbl, w.table), q
        if q in (4, 5, 6):
            assert eng.get_option(capi.WK_INFO_LAST_RESIDENT) == 1
        else:
            assert eng.get_option(capi.WK_INFO_LAST_RESIDENT) == 0
            assert eng.get_option(capi.WK_INFO_RESIDENT_RUNNING) == 1   # relaunched as soon as the heavy query was over
    # primitives park it as well
    eng.reset()
    n = eng.index_to_unknown(pid("Course"), O.IN)
    assert n > 0
    eng.close()               # destroy while the server may be resident: must not hang


def test_resident_spill_continues_on_the_multi_cta_path(gstore2, ostore2):
    eng = capi.Engine(gstore2, rbuf_bytes=64 << 20)
    pats, nvars, req = _spill_query()
    want = O.run_query([ostore2], pats, nvars, req)
    assert want.rows > 1024
    for resident in (True, False, True):
        eng.set_resident(resident)
        rc, rows, cols, tbl = eng.query(pats, nvars, req)
        assert rc == 0 and rows == want.rows and rows_equal(tbl, want.table), resident
        rc, rows_b, _, _ = eng.query(pats, nvars, req, blind=True)
        assert rc == 0 and rows_b == want.rows
        # a light query right after the spilled one
        p5, n5, r5, _ = load_query(5, PLANS[0])
        w5 = O.run_query([ostore2], p5, n5, r5)
        rc, rows, _, t5 = eng.query(p5, n5, r5)
        assert rc == 0 and rows_equal(t5, w5.table)
    eng.close()


def test_resident_step_stats(gstore2):
    eng = capi.Engine(gstore2, rbuf_bytes=64 << 20)
    pats, nvars, req, _ = load_query(6, PLANS[0])
    eng.set_profiling(2)
    out = {}
    for resident in (True, False):
        eng.set_resident(resident)
        rc, rows, _, _ = eng.query(pats, nvars, req, blind=True)
        assert rc == 0
        out[resident] = [(s["kind"], s["in_rows"], s["out_rows"], s["buckets_visited"], s["edges_touched"], s["algo_bytes"])
                         for s in eng.step_stats()]
    assert out[True] == out[False] and len(out[True]) == len(pats)
    eng.close()


def test_resident_servers_of_concurrent_engines(gstore2, ostore2):
    """several engines, each with its own resident server, answer light queries from their own host threads"""
    engs = [capi.Engine(gstore2, rbuf_bytes=32 << 20) for _ in range(3)]
    plans = [load_query(q, PLANS[0])[:3] for q in (4, 5, 6)]
    wants = [O.run_query([ostore2], *p) for p in plans]
    errs = []

    def work(i):
        try:
            for it in range(60):
                k = (it + i) % 3
                rc, rows, _, tbl = engs[i].query(*plans[k])
                assert rc == 0 and rows == wants[k].rows and rows_equal(tbl, wants[k].table), (i, it)
        except Exception as ex:   # noqa: BLE001
            errs.append(repr(ex))

    th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for e in engs:
        assert e.get_option(capi.WK_INFO_RESIDENT_REQUESTS) == 60
        e.close()
