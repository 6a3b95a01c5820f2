"""Light-query emulator workload (reference scripts/sparql_query/lubm/emulator: templates A1-A6 with a %type
placeholder, mix_config weights): instantiate templates with random instances of the placeholder's type
(Proxy::fill_template, proxy.hpp:69-129; SPARQLQuery_Template::instantiate, query.hpp:835-855)."""
import os
import re

import numpy as np

import sparql_mini as M
from conftest import ROOT

EMU = os.path.join(ROOT, "workloads", "lubm", "emulator")


def load_templates():
    """-> list of (name, weight, planned patterns with placeholder marker, nvars, type id of the placeholder, marker)"""
    cfg = [l.split() for l in open(os.path.join(EMU, "mix_config")) if l.strip() and not l.startswith("#")]
    out = []
    for name, w in cfg[1:]:
        text = open(os.path.join(EMU, name)).read()
        m = re.search(r"%(\w+:\w+)", text)
        tname = m.group(1)
        PLACEHOLDER = 0x7FFFFFF0
        pfx, local = tname.split(":")
        type_id = M.lubm_str2id(M.UB + local + ">")
        pats, nvars, req = M.parse_query(text.replace("%" + tname, "<__PH__>"),
                                         str2id=lambda s: PLACEHOLDER if s == "<__PH__>" else M.lubm_str2id(s))
        planned = M.apply_plan(pats, open(os.path.join(EMU, "osdi16_plan", name + ".fmt")).read())
        out.append((name, int(w), planned, nvars, type_id, PLACEHOLDER))
    return out


def instantiate(templates, candidates, n, seed=0):
    """candidates: {type_id: array of instance ids}.  -> (pats (m,4) int32, off (n+1) int32, nvars (n) int32, template idx (n))"""
    rng = np.random.default_rng(seed)
    w = np.array([t[1] for t in templates], dtype=np.float64)
    pick = rng.choice(len(templates), size=n, p=w / w.sum())
    pats, off, nv = [], [0], []
    for ti in pick:
        name, _, planned, nvars, type_id, ph = templates[ti]
        c = candidates[type_id]
        const = int(c[rng.integers(0, len(c))])
        for (s, p, d, o) in planned:
            pats.append((const if s == ph else s, p, d, const if o == ph else o))
        off.append(len(pats))
        nv.append(nvars)
    return np.array(pats, dtype=np.int32), np.array(off, dtype=np.int32), np.array(nv, dtype=np.int32), pick
