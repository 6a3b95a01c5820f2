"""Regenerates tests/golden/lubm_rows.json.

The reference's own tests hold no golden vectors for this path and the reference cannot be run here
(SURVEY.md §8c), so the fixture is produced from the oracle AFTER cross-checking every entry against
the independent brute-force joiner (tests/sparql_mini.py).  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import sparql_mini as M  # noqa: E402
from conftest import load_query, rows_equal  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_oracle import _checksum  # noqa: E402
from wukong_b200 import build, datagen  # noqa: E402

build.build_all()
out = {}
for name, U, seed, kv, ne in (("lubm1_seed1", 1, 1, 32 << 20, 4), ("lubm2_seed7", 2, 7, 48 << 20, 3)):
    tr = datagen.lubm(U, seed=seed)
    st = O.Store.build(tr, kvstore_bytes=kv, num_engines=ne)
    out[name] = {"triples": int(tr.shape[0])}
    for q in range(1, 8):
        pats, nvars, req, raw = load_query(q, "osdi16_plan")
        r = O.run_query([st], pats, nvars, req)
        bf = M.bruteforce_bgp(tr, raw, req)
        assert r.status == 0 and rows_equal(r.table, bf), (name, q)
        out[name]["q%d" % q] = {"rows": int(r.rows), "cols": int(r.cols), "checksum": _checksum(r.table)}
json.dump(out, open(os.path.join(HERE, "lubm_rows.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
