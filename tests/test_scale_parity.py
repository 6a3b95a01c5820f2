"""Content parity at scale (BASELINE config 2 and the multi-tile steady state of the step kernels).

* LUBM-40 (~5.5 M triples), Q1-Q7 x 3 plan sets: the GPU engine's tables against the committed answers of the reference's
  OWN engine (tests/golden/ref_engine_lubm40.json, made by make_ref_engine_lubm40.py from oracle/_ref), and against that
  engine live over the very store arrays the GPU holds whenever oracle/_ref travelled with the snapshot.
* Tables of 3.2 M rows x {1, 2, 3, 5} columns through known_to_unknown / known_to_known / known_to_const: the persistent
  grid is 592 CTAs x 256-row tiles, so every CTA runs ~21 iterations of the triple-buffered pipeline (the largest table of
  the other parity tests gives a CTA at most two tiles), the last tile is ragged, and a tenth of the rows miss.
Everything goes through the C ABI; the checker is the oracle (CPU) on the same arrays.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import sparql_mini as M
from conftest import PLANS, ROOT, load_query
from oracle import oracle as O
from oracle import ref as REF
from wukong_b200 import capi, datagen

pytestmark = pytest.mark.gpu

P = {n: i for i, n in enumerate(M.LUBM_INDEX)}
def pid(name): return P[M.UB + name + ">"]
TYPE = 1
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_engine_lubm40.json")


class _Arrays:
    def __init__(self, gst):
        self.v, self.e = gst.download()
        self.s = gst.segs()


@pytest.fixture(scope="module")
def lubm40():
    tr = datagen.lubm(40, seed=1)
    gst = capi.Store.build(tr, datagen.LUBM_NUM_NORMAL_PREDS)      # wk_store_build: the store the bench uses
    arr = _Arrays(gst)
    yield tr, gst, arr
    gst.close()


@pytest.fixture(scope="module")
def eng40(lubm40):
    e = capi.Engine(lubm40[1], rbuf_bytes=768 << 20)
    yield e
    e.close()


@pytest.fixture(scope="module")
def ost40(lubm40):
    a = lubm40[2]
    return O.Store.wrap(a.v, a.e, a.s)


def test_lubm40_matches_reference_engine(lubm40, eng40):
    gold = json.load(open(GOLDEN))
    assert gold["triples"] == int(lubm40[0].shape[0])
    live = REF.RefStore.adopt(lubm40[2].v, lubm40[2].e, lubm40[2].s) if REF.available() else None
    for q in range(1, 8):
        for plan in PLANS:
            pats, nvars, req, _ = load_query(q, plan)
            g = gold["queries"]["q%d_%s" % (q, plan)]
            for resident in (True, False):
                eng40.set_resident(resident)
                rc, rows, cols, tbl = eng40.query(pats, nvars, req)
                assert rc == 0 and rows == g["rows"], (q, plan, resident, rc, rows)
                if rows:
                    assert cols == g["cols"]
                    assert REF.table_digest(tbl) == g["digest"], (q, plan, resident)
                    assert hashlib.sha256(M.sort_rows(tbl).tobytes()).hexdigest() == g["sha256"], (q, plan, resident)
                rc, rows_b, _, _ = eng40.query(pats, nvars, req, blind=True)
                assert rc == 0 and rows_b == g["rows"]
            if live is not None:   # the reference's engine over the arrays the device builder produced
                rc, us, rrows, dg = live.time_query(pats, nvars, req, reps=1, mt_factor=8, threaded=True, digest=True)
                assert rc == 0 and rrows == g["rows"] and dg == g["digest"], (q, plan)
    eng40.set_resident(True)


def _big_table(ost, tp, n, extra_cols, key_col, seed):
    """n rows: column key_col = instances of type tp (resampled), a tenth replaced by ids that own no key; the other
    columns random"""
    rng = np.random.default_rng(seed)
    inst = ost.primitive(O.I2U, None, 0, tp, O.PREDICATE_ID, O.IN).reshape(-1)
    assert inst.size > 1000
    keys = inst[rng.integers(0, inst.size, n)]
    miss = rng.random(n) < 0.1
    keys[miss] = rng.integers(1 << 28, 1 << 29, int(miss.sum()), dtype=np.uint32)
    t = rng.integers(1 << 17, 1 << 24, (n, extra_cols + 1), dtype=np.uint32)
    t[:, key_col] = keys
    return t


N_BIG = 3_200_017   # 592 CTAs x 256 rows x ~21 tiles, ragged tail


@pytest.mark.parametrize("extra_cols,key_col", [(0, 0), (1, 1), (2, 0), (4, 2)])
def test_multi_tile_known_to_unknown(eng40, ost40, extra_cols, key_col):
    tbl = _big_table(ost40, pid("GraduateStudent"), N_BIG, extra_cols, key_col, seed=extra_cols)
    C = extra_cols + 1
    for p, d in ((pid("takesCourse"), O.OUT), (pid("memberOf"), O.OUT)):
        want = ost40.primitive(O.K2U, tbl, C, key_col, p, d)
        eng40.upload(tbl)
        n = eng40.known_to_unknown(key_col, p, d)
        got = eng40.download()
        assert n == want.shape[0] and n > N_BIG // 2
        assert got.shape == want.shape
        assert REF.table_digest(got) == REF.table_digest(want), (extra_cols, key_col, p)


@pytest.mark.parametrize("extra_cols", [0, 1, 2, 4])
def test_multi_tile_known_to_const_and_known(eng40, ost40, extra_cols):
    C = extra_cols + 1
    # known_to_const: is the row's vertex a GraduateStudent? (the table mixes graduate and undergraduate students)
    rng = np.random.default_rng(100 + extra_cols)
    a = _big_table(ost40, pid("GraduateStudent"), N_BIG // 2, extra_cols, extra_cols, seed=10 + extra_cols)
    b = _big_table(ost40, pid("UndergraduateStudent"), N_BIG - N_BIG // 2, extra_cols, extra_cols, seed=20 + extra_cols)
    tbl = np.concatenate([a, b])
    tbl = tbl[rng.permutation(tbl.shape[0])]
    want = ost40.primitive(O.K2C, tbl, C, extra_cols, TYPE, O.OUT, a_end=pid("GraduateStudent"))
    eng40.upload(tbl)
    n = eng40.known_to_const(extra_cols, TYPE, O.OUT, pid("GraduateStudent"))
    got = eng40.download()
    assert n == want.shape[0] and 0 < n < tbl.shape[0]
    assert REF.table_digest(got) == REF.table_digest(want)
    # known_to_known: (student, course) rows, half of the courses replaced by another row's course
    s = _big_table(ost40, pid("GraduateStudent"), N_BIG // 3, 0, 0, seed=30 + extra_cols)
    sc = ost40.primitive(O.K2U, s, 1, 0, pid("takesCourse"), O.OUT)
    sc = sc[: N_BIG] if sc.shape[0] > N_BIG else sc
    swap = rng.random(sc.shape[0]) < 0.5
    sc[swap, 1] = sc[rng.permutation(sc.shape[0])[: int(swap.sum())], 1]
    if extra_cols:
        sc = np.concatenate([sc, rng.integers(1 << 17, 1 << 24, (sc.shape[0], extra_cols), dtype=np.uint32)], axis=1)
    Ck = sc.shape[1]
    assert sc.shape[0] > 592 * 256 * 3
    for (cs, ce, p, d) in ((0, 1, pid("takesCourse"), O.OUT), (1, 0, pid("takesCourse"), O.IN)):
        want = ost40.primitive(O.K2K, sc, Ck, cs, p, d, a_end=ce)
        eng40.upload(sc)
        n = eng40.known_to_known(cs, p, d, ce)
        got = eng40.download()
        assert n == want.shape[0] and 0 < n < sc.shape[0]
        assert REF.table_digest(got) == REF.table_digest(want), (extra_cols, cs, ce)
