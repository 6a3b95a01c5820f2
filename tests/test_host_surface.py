"""The C++ host layer that mirrors the reference's surface (csrc/host/: Global config, StringServer, Parser,
Planner::set_plan, DGraph loader, GPUEngine, Proxy::run_single_query)."""
import os

import numpy as np
import pytest

import sparql_mini as M
from conftest import PLANS, WORKLOADS, load_query, rows_equal
from oracle import oracle as O
from wukong_b200 import datagen, host

CONFIG = """
# same key names as the reference's config file (scripts/config)
global_num_proxies              1
global_num_engines              4
global_input_folder             %s
global_memstore_size_gb         20
global_est_load_factor          55
global_mt_threshold             8
global_silent                   0
global_enable_planner           0
global_gpu_rbuf_size_mb         64
global_use_rdma                 0
global_some_future_key          7
"""


@pytest.fixture(scope="module")
def dataset_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("id_lubm_1"))
    datagen.lubm_write_dir(d, 1, seed=1)
    return d


def _text(q, plan):
    return (open(os.path.join(WORKLOADS, "lubm_q%d" % q)).read(), open(os.path.join(WORKLOADS, plan, "lubm_q%d.fmt" % q)).read())


def test_config_loader_parser_planner_cpu(dataset_dir, lubm1):
    env = host.Env(CONFIG % dataset_dir, device=-1)          # host-only: no GPU needed
    assert env.num_triples == lubm1.shape[0]                  # id_*.nt reader
    assert env.num_normal_preds == 31                         # str_index lines - 1 (base_loader.hpp:409-424)
    assert env.config_int("global_num_engines") == 4 and env.config_int("global_mt_threshold") == 8
    assert env.config_int("global_silent") == 0 and env.config_int("global_num_threads") == 5
    for q in range(1, 8):
        for plan in PLANS:
            query, fmt = _text(q, plan)
            rc, pats, nvars, req = env.parse_plan(query, fmt)
            want_pats, want_nvars, want_req, _ = load_query(q, plan)          # independent Python reader
            assert rc == 0 and pats == want_pats and nvars == want_nvars and req == want_req
            assert pats == O.set_plan(load_query(q, plan)[3], fmt)           # and the oracle's set_plan
    query, fmt = _text(2, "osdi16_plan")
    assert env.parse_plan(query.replace("ub:Course", "ub:NoSuchClass"), fmt)[0] == 2      # SYNTAX_ERROR: unknown IRI
    assert env.parse_plan(query.replace("SELECT", "SELEKT"), fmt)[0] == 2
    assert env.parse_plan(query.replace("SELECT", "SELECT DISTINCT") + " LIMIT 10 OFFSET 2", fmt)[0] == 0   # solution modifiers
    assert env.parse_plan(query + " ORDER BY ?X", fmt)[0] == 2                             # string ordering: not on this path
    assert env.parse_plan(query + " LIMIT -3", fmt)[0] == 2
    assert env.parse_plan(query, "1 <\n")[0] == 2                                          # plan shorter than the query
    assert env.parse_plan(query, "1 <\n9 >\n")[0] == 2                                     # pattern number out of range
    env.close()
    with pytest.raises(RuntimeError):
        host.Env("global_num_engines 0\nglobal_input_folder %s\n" % dataset_dir, device=-1)   # ASSERT(num_engines > 0)
    with pytest.raises(RuntimeError):
        host.Env("global_num_engines 2\n", device=-1)                                          # no input folder


@pytest.mark.gpu
def test_run_single_query_gpu(dataset_dir, ostore1):
    env = host.Env(CONFIG % dataset_dir, device=0)
    for q in range(1, 8):
        query, fmt = _text(q, "osdi16_plan")
        pats, nvars, req, _ = load_query(q, "osdi16_plan")
        want = O.run_query([ostore1], pats, nvars, req)
        for per_pattern in (False, True):        # one wk_query_execute call vs the reference-style agent loop
            rc, rows, cols, tbl, lat = env.run_single_query(query, fmt, cnt=3, per_pattern=per_pattern)
            assert rc == 0 and rows == want.rows, (q, per_pattern)
            if rows:
                assert cols == want.cols and rows_equal(tbl, want.table), (q, per_pattern)
            assert lat > 0
    # SELECT DISTINCT ... LIMIT / OFFSET: final_process modifiers through both host paths (sparql.hpp:1428-1499)
    query, fmt = _text(7, "osdi16_plan")
    pats, nvars, req, _ = load_query(7, "osdi16_plan")
    assert "SELECT ?X ?Y ?Z" in query
    q1 = query.replace("SELECT ?X ?Y ?Z", "SELECT DISTINCT ?X")
    q1 = q1 + " LIMIT 20 OFFSET 4"
    want = O.run_query([ostore1], pats, nvars, req[:1], distinct=True, offset=4, limit=20)
    for per_pattern in (False, True):
        rc, rows, cols, tbl, lat = env.run_single_query(q1, fmt, per_pattern=per_pattern)
        assert rc == 0 and rows == want.rows == 20 and cols == 1, per_pattern
        assert np.array_equal(tbl, want.table), per_pattern
    # silent mode: only the row count comes back (Global::silent, proxy.hpp:360-369)
    env2 = host.Env((CONFIG % dataset_dir).replace("global_silent                   0", "global_silent 1"), device=0)
    query, fmt = _text(2, "osdi16_plan")
    rc, rows, cols, tbl, lat = env2.run_single_query(query, fmt, cnt=2)
    assert rc == 0 and rows == 942 and (tbl is None or tbl.size == 0)
    # error codes surface through the reply like in the reference (sparql.hpp:1663-1667)
    bad = "SELECT ?X WHERE { ?X <http://swat.cse.lehigh.edu/onto/univ-bench.owl#memberOf> ?Y . }"
    rc, *_ = env2.run_single_query(bad, "1 >\n")
    assert rc == 9      # UNKNOWN_SUB: starts from an unbound variable
    env.close()
    env2.close()
