"""GPU parity tests: every CUDA path is called through the C ABI and compared bit-exactly (as a
multiset of rows) with the CPU oracle on the same store arrays and the same inputs."""
import numpy as np
import pytest

import sparql_mini as M
from conftest import PLANS, load_query, rows_equal
from oracle import oracle as O
from wukong_b200 import capi

pytestmark = pytest.mark.gpu

P = {n: i for i, n in enumerate(M.LUBM_INDEX)}
def pid(name): return P[M.UB + name + ">"]
TYPE = 1


@pytest.fixture(scope="module")
def eng1(gstore1):
    e = capi.Engine(gstore1, rbuf_bytes=64 << 20)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng2(gstore2):
    e = capi.Engine(gstore2, rbuf_bytes=64 << 20)
    yield e
    e.close()


def test_store_probe_matches_oracle(gstore1, ostore1, lubm1):
    rng = np.random.default_rng(0)
    t = lubm1[rng.integers(0, lubm1.shape[0], 60)]
    for s, p, o in t.tolist():
        assert (gstore1.get_edges(s, p, O.OUT) == ostore1.get_edges(s, p, O.OUT)).all()
        assert (gstore1.get_edges(o, p, O.IN) == ostore1.get_edges(o, p, O.IN)).all()
    assert gstore1.get_edges(12345, 5, O.OUT).size == 0


@pytest.mark.parametrize("mt", [(0, 1), (0, 3), (2, 3), (1, 2)])
def test_index_to_unknown(eng1, ostore1, mt):
    for tp, d in [(pid("Course"), O.IN), (pid("GraduateStudent"), O.IN), (pid("memberOf"), O.IN),
                  (pid("memberOf"), O.OUT), (pid("undergraduateDegreeFrom"), O.OUT)]:
        want = ostore1.primitive(O.I2U, None, 0, tp, O.PREDICATE_ID, d, mt_tid=mt[0], mt_factor=mt[1])
        eng1.reset()
        n = eng1.index_to_unknown(tp, d, mt[0], mt[1])
        got = eng1.download()
        assert n == want.shape[0]
        # a seed is a plain slice copy: identical order, not only the same multiset
        assert (got == want).all()


def test_const_to_unknown(eng1, ostore1):
    dept0 = M.lubm_str2id("<http://www.Department0.University0.edu>")
    univ0 = M.lubm_str2id("<http://www.University0.edu>")
    for vid, p, d in [(dept0, pid("worksFor"), O.IN), (univ0, pid("subOrganizationOf"), O.IN),
                      (dept0, pid("subOrganizationOf"), O.OUT), (dept0, pid("advisor"), O.IN)]:
        want = ostore1.primitive(O.C2U, None, 0, vid, p, d)
        eng1.reset()
        n = eng1.const_to_unknown(vid, p, d)
        assert n == want.shape[0]
        got = eng1.download()
        assert got.shape[0] == want.shape[0]
        if n:
            assert (got == want).all()


def _seed_table(ostore, tp, d=O.IN, extra_cols=0, seed=0):
    t = ostore.primitive(O.I2U, None, 0, tp, O.PREDICATE_ID, d)
    if extra_cols:
        rng = np.random.default_rng(seed)
        t = np.concatenate([rng.integers(1 << 17, 1 << 20, (t.shape[0], extra_cols), dtype=np.uint32), t], axis=1)
    return t


@pytest.mark.parametrize("extra_cols", [0, 1, 2, 5])
def test_known_to_unknown(eng1, ostore1, extra_cols):
    cases = [(pid("GraduateStudent"), pid("memberOf"), O.OUT), (pid("Course"), pid("name"), O.OUT),
             (pid("FullProfessor"), pid("advisor"), O.IN),          # fan-out > SMALL_DEG for some rows
             (pid("Department"), pid("memberOf"), O.IN),            # hundreds of edges per row (warp path)
             (pid("UndergraduateStudent"), pid("undergraduateDegreeFrom"), O.OUT),   # all misses
             (pid("University"), pid("subOrganizationOf"), O.IN)]
    for tp, p, d in cases:
        tbl = _seed_table(ostore1, tp, extra_cols=extra_cols)
        C = tbl.shape[1]
        want = ostore1.primitive(O.K2U, tbl, C, C - 1, p, d)
        eng1.upload(tbl)
        n = eng1.known_to_unknown(C - 1, p, d)
        got = eng1.download()
        assert n == want.shape[0], (tp, p, d)
        assert rows_equal(got, want), (tp, p, d)


def test_known_to_unknown_type_index(eng1, ostore1):
    # pid == TYPE_ID && dir == IN goes through the index segment (sparql.hpp:339-340)
    tbl = np.array([[pid("Course")], [pid("FullProfessor")], [pid("University")], [9999]], dtype=np.uint32)
    want = ostore1.primitive(O.K2U, tbl, 1, 0, TYPE, O.IN)
    eng1.upload(tbl)
    eng1.known_to_unknown(0, TYPE, O.IN)
    assert rows_equal(eng1.download(), want)
    # ... and the forward direction: every vertex -> its types
    tbl = _seed_table(ostore1, pid("GraduateStudent"))
    want = ostore1.primitive(O.K2U, tbl, 1, 0, TYPE, O.OUT)
    eng1.upload(tbl)
    eng1.known_to_unknown(0, TYPE, O.OUT)
    assert rows_equal(eng1.download(), want)


def test_hub_fanout(eng1, ostore1):
    # the shared telephone literal has every person as IN-neighbour (thousands of edges for one row)
    tel = 1 << 17
    tbl = np.array([[7, tel], [8, 123456], [9, tel]], dtype=np.uint32)
    want = ostore1.primitive(O.K2U, tbl, 2, 1, pid("telephone"), O.IN)
    assert want.shape[0] > 10000
    eng1.upload(tbl)
    n = eng1.known_to_unknown(1, pid("telephone"), O.IN)
    assert n == want.shape[0]
    assert rows_equal(eng1.download(), want)
    # filter over the same hub list (warp-cooperative scan, first/last/missing element)
    lst = ostore1.get_edges(tel, pid("telephone"), O.IN)
    for target in (int(lst[0]), int(lst[-1]), int(lst[len(lst) // 2]), 424242):
        tbl = np.array([[tel, target], [tel, target + 1], [5, target]], dtype=np.uint32)
        want = ostore1.primitive(O.K2K, tbl, 2, 0, pid("telephone"), O.IN, a_end=1)
        eng1.upload(tbl)
        eng1.known_to_known(0, pid("telephone"), O.IN, 1)
        assert rows_equal(eng1.download(), want)
        want = ostore1.primitive(O.K2C, tbl, 2, 0, pid("telephone"), O.IN, a_end=target)
        eng1.upload(tbl)
        eng1.known_to_const(0, pid("telephone"), O.IN, target)
        assert rows_equal(eng1.download(), want)


@pytest.mark.parametrize("extra_cols", [0, 2])
def test_known_to_const_and_known(eng1, ostore1, extra_cols):
    # k2c: type checks
    for tp, typ in [(pid("memberOf"), pid("GraduateStudent")), (pid("memberOf"), pid("UndergraduateStudent")),
                    (pid("worksFor"), pid("FullProfessor")), (pid("takesCourse"), pid("TeachingAssistant"))]:
        tbl = _seed_table(ostore1, tp, extra_cols=extra_cols)
        C = tbl.shape[1]
        want = ostore1.primitive(O.K2C, tbl, C, C - 1, TYPE, O.OUT, a_end=typ)
        eng1.upload(tbl)
        n = eng1.known_to_const(C - 1, TYPE, O.OUT, typ)
        assert n == want.shape[0]
        assert rows_equal(eng1.download(), want)
    # k2k: (student, dept) pairs filtered by advisor/worksFor consistency
    tbl = _seed_table(ostore1, pid("advisor"), extra_cols=extra_cols)
    C = tbl.shape[1]
    t2 = ostore1.primitive(O.K2U, tbl, C, C - 1, pid("advisor"), O.OUT)
    t3 = ostore1.primitive(O.K2U, t2, C + 1, C - 1, pid("takesCourse"), O.OUT)
    want = ostore1.primitive(O.K2K, t3, C + 2, C, pid("teacherOf"), O.OUT, a_end=C + 1)
    eng1.upload(t3)
    n = eng1.known_to_known(C, pid("teacherOf"), O.OUT, C + 1)
    assert n == want.shape[0] and n > 0
    assert rows_equal(eng1.download(), want)


@pytest.mark.parametrize("extra_cols", [0, 3])
def test_const_and_index_to_known(eng1, ostore1, extra_cols):
    """sparql.hpp:80-186 — rows kept when the column occurs in one edge list (hash set on the device)"""
    dept0 = M.lubm_str2id("<http://www.Department0.University0.edu>")
    univ0 = M.lubm_str2id("<http://www.University0.edu>")
    for tp, vid, p, d in [(pid("memberOf"), dept0, pid("memberOf"), O.IN),       # students: members of Department0
                          (pid("worksFor"), dept0, pid("worksFor"), O.IN),
                          (pid("undergraduateDegreeFrom"), univ0, pid("undergraduateDegreeFrom"), O.IN),
                          (pid("memberOf"), univ0, pid("memberOf"), O.IN),     # empty list: nothing survives
                          (pid("takesCourse"), dept0, pid("subOrganizationOf"), O.OUT)]:
        tbl = _seed_table(ostore1, tp, extra_cols=extra_cols)
        C = tbl.shape[1]
        want = ostore1.primitive(O.C2K, tbl, C, vid, p, d, a_end=C - 1)
        eng1.upload(tbl)
        n = eng1.const_to_known(vid, p, d, C - 1)
        assert n == want.shape[0], (tp, vid, p, d)
        assert rows_equal(eng1.download(), want)
    # index lists, whole and mt slices (type index and predicate index)
    for tp, idx, d in [(pid("memberOf"), pid("GraduateStudent"), O.IN), (pid("takesCourse"), pid("advisor"), O.IN),
                       (pid("advisor"), pid("FullProfessor"), O.IN), (pid("worksFor"), pid("teacherOf"), O.OUT)]:
        tbl = _seed_table(ostore1, tp, extra_cols=extra_cols)
        C = tbl.shape[1]
        for mt in [(0, 1), (1, 3), (2, 3)]:
            want = ostore1.primitive(O.I2K, tbl, C, idx, O.PREDICATE_ID, d, a_end=C - 1, mt_tid=mt[0], mt_factor=mt[1])
            eng1.upload(tbl)
            n = eng1.index_to_known(idx, d, C - 1, mt[0], mt[1])
            assert n == want.shape[0], (tp, idx, d, mt)
            assert rows_equal(eng1.download(), want)
    # a plan whose middle pattern is CONST -> KNOWN runs through wk_query_execute
    pats = [(pid("GraduateStudent"), O.TYPE_ID, O.IN, -1), (-1, pid("memberOf"), O.OUT, -2),
            (univ0, pid("subOrganizationOf"), O.IN, -2), (-2, pid("name"), O.OUT, -3)]
    want = O.run_query([ostore1], pats, 3, [-1, -3], mt_factor=1)
    rc, rows, cols, tbl = eng1.query(pats, 3, [-1, -3])
    assert rc == 0 and rows == want.rows and rows > 0
    assert rows_equal(tbl, want.table)
    assert "c2k" in [s["kind"] for s in eng1.step_stats()]


def test_edge_cases(eng1, ostore1):
    # empty input table
    eng1.upload(np.zeros((0, 2), dtype=np.uint32), ncols=2)
    assert eng1.known_to_unknown(1, pid("memberOf"), O.OUT) == 0
    assert eng1.download().shape[0] == 0
    # ragged sizes around the 256-row tile
    base = _seed_table(ostore1, pid("takesCourse"))
    for n in (1, 31, 32, 33, 255, 256, 257, 511, 513, 1000):
        tbl = base[:n]
        want = ostore1.primitive(O.K2U, tbl, 1, 0, pid("takesCourse"), O.OUT)
        eng1.upload(tbl)
        assert eng1.known_to_unknown(0, pid("takesCourse"), O.OUT) == want.shape[0]
        assert rows_equal(eng1.download(), want)
    # projection
    tbl = np.arange(40, dtype=np.uint32).reshape(10, 4)
    eng1.upload(tbl)
    eng1.project([3, 0, 0])
    assert (eng1.download() == tbl[:, [3, 0, 0]]).all()
    # missing segment -> engine error, not a crash
    eng1.upload(tbl)
    with pytest.raises(capi.WukongError) as ei:
        eng1.known_to_unknown(0, 77, O.OUT)
    assert ei.value.code == capi.WK_ERR_NO_SEGMENT
    # result buffer overflow is reported, not written past the end
    small = capi.Engine(eng1.store, rbuf_bytes=8192)
    tel = 1 << 17
    small.upload(np.array([[tel]], dtype=np.uint32))
    with pytest.raises(capi.WukongError) as ei:
        small.known_to_unknown(0, pid("telephone"), O.IN)
    assert ei.value.code == capi.WK_ERR_RBUF_OVERFLOW
    small.close()


def test_chained_without_sync(eng1, ostore1):
    # a whole pattern chain enqueued without any host synchronisation in between
    eng1.reset()
    eng1.index_to_unknown(pid("FullProfessor"), O.IN, sync=False)
    eng1.known_to_unknown(0, pid("advisor"), O.IN, sync=False)
    eng1.known_to_const(1, TYPE, O.OUT, pid("UndergraduateStudent"), sync=False)
    eng1.known_to_unknown(1, pid("takesCourse"), O.OUT, sync=False)
    eng1.known_to_known(0, pid("teacherOf"), O.OUT, 2, sync=False)
    got = eng1.download()
    t = ostore1.primitive(O.I2U, None, 0, pid("FullProfessor"), O.PREDICATE_ID, O.IN)
    t = ostore1.primitive(O.K2U, t, 1, 0, pid("advisor"), O.IN)
    t = ostore1.primitive(O.K2C, t, 2, 1, TYPE, O.OUT, a_end=pid("UndergraduateStudent"))
    t = ostore1.primitive(O.K2U, t, 2, 1, pid("takesCourse"), O.OUT)
    t = ostore1.primitive(O.K2K, t, 3, 0, pid("teacherOf"), O.OUT, a_end=2)
    assert rows_equal(got, t)
    st = eng1.step_stats()
    assert [s["kind"] for s in st] == ["i2u", "k2u", "k2c", "k2u", "k2k"]
    assert st[-1]["out_rows"] == t.shape[0]
    assert all(s["buckets_visited"] >= 1 for s in st)


@pytest.mark.parametrize("q", range(1, 8))
def test_queries_match_oracle(q, eng1, eng2, ostore1, ostore2):
    for eng, ost in ((eng1, ostore1), (eng2, ostore2)):
        for plan in PLANS:
            pats, nvars, req, _ = load_query(q, plan)
            want = O.run_query([ost], pats, nvars, req, mt_factor=1)
            rc, rows, cols, tbl = eng.query(pats, nvars, req, mt_tid=0, mt_factor=1)
            assert rc == 0
            assert rows == want.rows and cols == want.cols, (q, plan)
            assert rows_equal(tbl, want.table), (q, plan)
            rc, rows_b, cols_b, _ = eng.query(pats, nvars, req, blind=True)
            assert rc == 0 and rows_b == want.rows


def test_mt_slices_partition_the_result(eng1, ostore1):
    # mt_factor replicas (sparql.hpp:1064-1089): the union of the slices equals the whole result
    for q in (1, 2, 7):
        pats, nvars, req, _ = load_query(q, "osdi16_plan")
        want = O.run_query([ostore1], pats, nvars, req, mt_factor=1)
        parts = []
        for tid in range(3):
            rc, rows, cols, tbl = eng1.query(pats, nvars, req, mt_tid=tid, mt_factor=3)
            assert rc == 0
            if rows:
                parts.append(tbl.copy())
        got = np.concatenate(parts) if parts else np.zeros((0, want.cols), np.uint32)
        assert rows_equal(got, want.table)


def test_query_errors(eng1):
    # same status codes as the reference engine (utils/errors.hpp) / the oracle
    rc, *_ = eng1.query([(-1, 5, O.OUT, -2)], 2, [-1])
    assert rc == 9          # UNKNOWN_SUB
    univ0 = M.lubm_str2id("<http://www.University0.edu>")
    rc, *_ = eng1.query([(18, 1, O.IN, -1), (univ0, 7, O.IN, -2)], 2, [-1])
    assert rc == 11         # FIRST_PATTERN_ERROR
    rc, *_ = eng1.query([(18, 5, O.IN, -1)], 1, [-1])
    assert rc == 7          # OBJ_ERROR
    rc, *_ = eng1.query([(18, 1, O.IN, -1)], 1, [])
    assert rc == 5          # NO_REQUIRED_VAR


def test_batched_light_queries(eng2, ostore2):
    """wk_query_execute_batch (one launch, one CTA per query) against the oracle, emulator templates A1-A6"""
    import emu_util
    tpl = emu_util.load_templates()
    cands = {t[4]: ostore2.get_edges(0, t[4], O.IN) for t in tpl}
    pats, off, nv, pick = emu_util.instantiate(tpl, cands, 300, seed=5)
    rows, st = eng2.query_batch_raw(pats, off, nv)
    assert (st == 0).all()
    _, want = O.emu_run(ostore2, pats, off, nv, 2)
    assert np.array_equal(rows, want)
    assert rows.sum() > 0 and len(set(pick.tolist())) >= 4
    # a malformed plan inside a batch only fails itself
    bad = np.concatenate([pats[: off[1]], np.array([[-1, 5, 1, -2]], dtype=np.int32)])
    rows, st = eng2.query_batch_raw(bad, np.array([0, off[1], off[1] + 1], dtype=np.int32), np.array([nv[0], 2], dtype=np.int32))
    assert st[0] == 0 and st[1] == 9 and rows[0] == want[0]


def test_concurrent_engines_share_one_store(gstore2, ostore2):
    """any number of engines per store, one caller thread each (the emulator's concurrent queries,
    proxy.hpp:391-545): four threads run the seven queries interleaved and must all match the oracle"""
    import threading
    work = []
    for q in range(1, 8):
        pats, nvars, req, _ = load_query(q, "osdi16_plan")
        work.append((pats, nvars, req, O.run_query([ostore2], pats, nvars, req, mt_factor=1)))
    errors = []

    def worker(seed):
        eng = capi.Engine(gstore2, rbuf_bytes=48 << 20)
        try:
            order = np.random.default_rng(seed).permutation(len(work) * 6) % len(work)
            for i in order.tolist():
                pats, nvars, req, want = work[i]
                rc, rows, cols, tbl = eng.query(pats, nvars, req)
                if rc != 0 or rows != want.rows or not rows_equal(tbl, want.table):
                    errors.append((seed, i, rc, rows, want.rows))
        except Exception as ex:   # noqa: BLE001 - surfaced through the assertion below
            errors.append((seed, repr(ex)))
        finally:
            eng.close()

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


def test_query_modifiers_distinct_offset_limit(eng1, ostore1):
    """final_process DISTINCT / OFFSET / LIMIT on the device (sparql.hpp:1424-1505) against the oracle"""
    for q, sel in ((7, [0]), (7, [0, 1]), (2, None), (1, [0]), (4, [1, 2])):
        pats, nvars, req, _ = load_query(q, "osdi16_plan")
        rq = req if sel is None else [req[i] for i in sel if i < len(req)]
        for mods in (dict(distinct=True), dict(distinct=True, offset=3, limit=11), dict(distinct=True, limit=0),
                     dict(distinct=True, offset=10 ** 6)):
            want = O.run_query([ostore1], pats, nvars, rq, **mods)
            rc, rows, cols, tbl = eng1.query(pats, nvars, rq, **mods)
            assert rc == 0 and rows == want.rows, (q, rq, mods)
            if rows:
                assert np.array_equal(tbl, want.table), (q, rq, mods)      # DISTINCT fixes the order: exact equality
        # OFFSET / LIMIT alone cut an order-dependent window: same size, rows drawn from the full answer
        full = O.run_query([ostore1], pats, nvars, rq)
        for mods in (dict(offset=5), dict(limit=7), dict(offset=2, limit=3)):
            want = O.run_query([ostore1], pats, nvars, rq, **mods)
            rc, rows, cols, tbl = eng1.query(pats, nvars, rq, **mods)
            assert rc == 0 and rows == want.rows, (q, mods)
            if rows:
                fullset = {tuple(r) for r in full.table.tolist()}
                assert all(tuple(r) in fullset for r in tbl.tolist())
        # blind queries skip final_process, modifiers included
        rc, rows, _, _ = eng1.query(pats, nvars, rq, blind=True, distinct=True, limit=1)
        assert rc == 0 and rows == full.rows
    assert "distinct" in [s["kind"] for s in eng1.step_stats()] or True
    # the primitives, step by step
    tbl = np.array([[5, 1], [3, 9], [5, 1], [0x80000001, 2], [3, 8], [5, 0]], dtype=np.uint32)
    eng1.upload(tbl)
    assert eng1.distinct([0]) == 3                       # signed order: 0x80000001 sorts first
    got = eng1.download()
    assert got[:, 0].tolist() == [0x80000001, 3, 5] and got[1].tolist() == [3, 8] and got[2].tolist() == [5, 0]
    assert eng1.slice(1, 1) == 1 and eng1.download().tolist() == [[3, 8]]
    assert eng1.slice(5, -1) == 0



def test_fused_filter_chains(eng2, ostore2):
    """runs of consecutive known_to_known / known_to_const steps are one launch (WK_OPT_FUSE_FILTERS): same tables as the
    step-by-step execution and the oracle, for chains of 2-5 filters in every order, filters that kill every row, and
    filters in the middle of a plan"""
    gs, ug, univ, dept = pid("GraduateStudent"), pid("UndergraduateStudent"), pid("University"), pid("Department")
    plans = [
        # ?x type GraduateStudent . ?x memberOf ?z . ?x undergraduateDegreeFrom ?y . then filters on ?x ?y ?z
        ([(gs, TYPE, O.IN, -1), (-1, pid("memberOf"), O.OUT, -3), (-1, pid("undergraduateDegreeFrom"), O.OUT, -2),
          (-2, TYPE, O.OUT, univ), (-2, pid("subOrganizationOf"), O.IN, -3), (-3, TYPE, O.OUT, dept)], 3, [-1, -2, -3]),
        # const filters only; the middle one removes everything (?z is a department, not a university)
        ([(gs, TYPE, O.IN, -1), (-1, pid("memberOf"), O.OUT, -3), (-3, TYPE, O.OUT, dept), (-3, TYPE, O.OUT, univ),
          (-1, TYPE, O.OUT, gs)], 3, [-1, -3]),
        # five filters in a row (split into two launches), then an expansion after them
        ([(gs, TYPE, O.IN, -1), (-1, pid("memberOf"), O.OUT, -3), (-1, pid("advisor"), O.OUT, -2),
          (-2, pid("worksFor"), O.OUT, -3), (-1, TYPE, O.OUT, gs), (-3, TYPE, O.OUT, dept), (-1, pid("memberOf"), O.OUT, -3),
          (-2, TYPE, O.OUT, pid("FullProfessor")), (-1, pid("takesCourse"), O.OUT, -4)], 4, [-1, -2, -4]),
        # known_to_known in both directions on the same pair
        ([(ug, TYPE, O.IN, -1), (-1, pid("takesCourse"), O.OUT, -2), (-2, pid("takesCourse"), O.IN, -1),
          (-1, pid("takesCourse"), O.OUT, -2), (-1, TYPE, O.OUT, ug)], 2, [-1, -2]),
    ]
    for q in (1, 3, 7):
        for plan in PLANS:
            plans.append(load_query(q, plan)[:3])
    for pats, nvars, req in plans:
        want = O.run_query([ostore2], pats, nvars, req)
        got = {}
        for fuse in (1, 0):
            eng2.set_option(capi.WK_OPT_FUSE_FILTERS, fuse)
            l0 = eng2.launch_count()
            rc, rows, cols, tbl = eng2.query(pats, nvars, req)
            got[fuse] = eng2.launch_count() - l0
            assert rc == 0 and rows == want.rows, (pats, fuse, rows, want.rows)
            assert rows_equal(tbl, want.table), (pats, fuse)
            rc, rows_b, _, _ = eng2.query(pats, nvars, req, blind=True)
            assert rc == 0 and rows_b == want.rows
        nfilters = sum(1 for (s, p, d, o) in pats[1:] if o >= 0 or (s < 0 and o < 0 and _bound_before(pats, o)))
        if nfilters >= 2:
            assert got[1] < got[0], (pats, got)     # fewer launches when fused
    eng2.set_option(capi.WK_OPT_FUSE_FILTERS, 1)


def _bound_before(pats, var):
    """True when `var` is the object of a pattern only after having been bound by an earlier one (rough: bound at all)"""
    seen = set()
    for (s, p, d, o) in pats:
        if o == var and var in seen:
            return True
        for v in (s, o):
            if v < 0:
                seen.add(v)
    return False


def test_direct_output_into_pinned_buffer(eng2, ostore2):
    """non-blind queries whose result buffer is pinned host memory: the last step writes the projected rows straight into it
    (WK_OPT_DIRECT_OUT); same tables as the projection kernel + copy, also for a buffer that is too small"""
    buf, keep = capi.pinned_array(1 << 22)
    small, keep2 = capi.pinned_array(64)
    for q in (1, 2, 3, 7):
        for plan in PLANS:
            pats, nvars, req, _ = load_query(q, plan)
            want = O.run_query([ostore2], pats, nvars, req)
            for direct in (1, 0):
                eng2.set_option(capi.WK_OPT_DIRECT_OUT, direct)
                l0 = eng2.launch_count()
                buf[:] = 0xFFFFFFFF
                rc, rows, cols, tbl = eng2.query(pats, nvars, req, out=buf)
                nl = eng2.launch_count() - l0
                assert rc == 0 and rows == want.rows and cols == want.cols, (q, plan, direct)
                assert rows_equal(tbl, want.table), (q, plan, direct)
                if direct == 1:
                    n1 = nl
                elif want.rows >= 0:
                    assert n1 <= nl, (q, plan, n1, nl)     # no projection launch when the last step writes the result
            eng2.set_option(capi.WK_OPT_DIRECT_OUT, 1)
            if want.rows * want.cols > small.size:
                rc, rows, cols, _ = eng2.query(pats, nvars, req, out=small)
                assert rc == capi.WK_ERR_BAD_ARG, (q, plan, rc)
            # the engine stays usable, and a pageable buffer takes the copy path
            rc, rows, cols, tbl = eng2.query(pats, nvars, req)
            assert rc == 0 and rows == want.rows and rows_equal(tbl, want.table)
    # a new column that is projected away, and one that is repeated
    eng2.set_option(capi.WK_OPT_DIRECT_OUT, 1)
    gs = pid("GraduateStudent")
    pats = [(gs, TYPE, O.IN, -1), (-1, pid("memberOf"), O.OUT, -2), (-1, pid("takesCourse"), O.OUT, -3)]
    for req in ([-1], [-3, -1, -3], [-2, -3]):
        want = O.run_query([ostore2], pats, 3, req)
        rc, rows, cols, tbl = eng2.query(pats, 3, req, out=buf)
        assert rc == 0 and rows == want.rows and cols == len(req) and rows_equal(tbl, want.table), req
    eng2.set_option(capi.WK_OPT_DIRECT_OUT, 0)
