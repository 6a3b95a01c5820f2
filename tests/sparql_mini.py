"""Independent (Python) mini reader for the basic LUBM query files + brute-force BGP evaluator.

Used only by tests: it gives a second opinion that shares no code with either the oracle
(oracle/wukong_oracle.cpp) or the product's C++ parser (wukong_b200/csrc/host/parser.hpp).

Variable numbering follows the reference parser (SPARQLParser.hpp:226-235, 1108-1135;
parser.hpp:186-197): named variables get -1, -2, ... in order of first appearance, and the
SELECT projection is parsed before WHERE.
"""
import re
from collections import defaultdict

import numpy as np

RDF_TYPE = "<http://www.w3.org/1999/02/22-rdf-syntax-ns#type>"
UB = "<http://swat.cse.lehigh.edu/onto/univ-bench.owl#"

IN, OUT = 0, 1
PREDICATE_ID, TYPE_ID = 0, 1

LUBM_INDEX = ["__PREDICATE__", RDF_TYPE] + [UB + n + ">" for n in (
    "name emailAddress telephone memberOf worksFor subOrganizationOf undergraduateDegreeFrom "
    "mastersDegreeFrom doctoralDegreeFrom advisor takesCourse teacherOf teachingAssistantOf "
    "researchInterest headOf publicationAuthor University Department FullProfessor AssociateProfessor "
    "AssistantProfessor Lecturer UndergraduateStudent GraduateStudent Course GraduateCourse "
    "ResearchGroup Publication TeachingAssistant ResearchAssistant").split()]

VID_BASE = 1 << 17
UNIV_BASE = VID_BASE + (1 << 17)
UNIV_BLOCK = (1 << 17) - 1


def lubm_str2id(s):
    """Closed-form string -> id of the synthetic generator (wukong_b200/csrc/datagen/lubm_vocab.h)."""
    if s in LUBM_INDEX:
        return LUBM_INDEX.index(s)
    m = re.fullmatch(r"<http://www\.University(\d+)\.edu>", s)
    if m:
        return UNIV_BASE + int(m.group(1)) * UNIV_BLOCK
    m = re.fullmatch(r"<http://www\.Department(\d+)\.University(\d+)\.edu>", s)
    if m:
        return UNIV_BASE + int(m.group(2)) * UNIV_BLOCK + 2 + int(m.group(1))
    raise KeyError(s)


def parse_query(text, str2id=lubm_str2id):
    """Returns (patterns [(s,p,OUT,o)], nvars, required_vars)."""
    prefixes = {}
    for m in re.finditer(r"PREFIX\s+(\w*):\s*<([^>]*)>", text):
        prefixes[m.group(1)] = m.group(2)
    m = re.search(r"SELECT\s+(.*?)\s+WHERE\s*\{(.*)\}", text, re.S)
    proj, body = m.group(1), m.group(2)
    var_ids = {}

    def var(name):
        if name not in var_ids:
            var_ids[name] = -(len(var_ids) + 1)
        return var_ids[name]

    required = [var(v) for v in re.findall(r"\?(\w+)", proj)]

    def term(tok):
        if tok.startswith("?"):
            return var(tok[1:])
        if tok == "__PREDICATE__":
            return PREDICATE_ID
        if tok.startswith("<"):
            return str2id(tok)
        pfx, local = tok.split(":", 1)
        return str2id("<" + prefixes[pfx] + local + ">")

    patterns = []
    for stmt in body.split(" ."):
        toks = stmt.replace("\t", " ").split()
        if toks and toks[-1] == ".":
            toks = toks[:-1]
        if not toks:
            continue
        assert len(toks) == 3, toks
        s, p, o = (term(t) for t in toks)
        patterns.append((s, p, OUT, o))
    return patterns, len(var_ids), required


def apply_plan(patterns, fmt_text):
    """.fmt plan application (reference planner.hpp:1647-1754), independent restatement."""
    out = []
    for line in fmt_text.splitlines():
        line = line.strip()
        if not line or line.startswith("#") or line == "{":
            continue
        if line == "}":          # end of the group: whatever follows belongs to nobody (planner.hpp:1728-1729)
            break
        parts = line.split()
        order, d = int(parts[0]), (parts[1] if len(parts) > 1 else ">")
        s, p, _, o = patterns[order - 1]
        if d == "<":
            out.append((o, p, IN, s))
        elif d == ">":
            out.append((s, p, OUT, o))
        elif d == "<<":
            out.append((p, PREDICATE_ID, IN, s))
        elif d == ">>":
            out.append((p, PREDICATE_ID, OUT, o))
        else:
            raise ValueError(d)
    assert len(out) >= len(patterns)
    return out


def bruteforce_bgp(triples, patterns, required):
    """Evaluate the basic graph pattern by plain joins over the deduplicated triple list.

    Returns the projected bindings as a lexicographically sorted (rows, len(required)) array.
    """
    t = np.unique(np.asarray(triples, dtype=np.uint32).reshape(-1, 3), axis=0)
    by_p = defaultdict(list)
    for s, p, o in t.tolist():
        by_p[p].append((s, o))
    rows = [dict()]
    for (s, p, _d, o) in patterns:
        assert p >= 0, "variable predicates are out of scope"
        pairs = by_p.get(p, [])
        idx_s, idx_o = defaultdict(list), defaultdict(list)
        for a, b in pairs:
            idx_s[a].append(b)
            idx_o[b].append(a)
        pair_set = set(pairs)
        new_rows = []
        for r in rows:
            sv = s if s >= 0 else r.get(s)
            ov = o if o >= 0 else r.get(o)
            if sv is not None and ov is not None:
                if (sv, ov) in pair_set:
                    new_rows.append(r)
            elif sv is not None:
                for b in idx_s.get(sv, ()):
                    nr = dict(r)
                    nr[o] = b
                    new_rows.append(nr)
            elif ov is not None:
                for a in idx_o.get(ov, ()):
                    nr = dict(r)
                    nr[s] = a
                    new_rows.append(nr)
            else:
                for a, b in pairs:
                    if s == o and a != b:
                        continue
                    nr = dict(r)
                    nr[s] = a
                    nr[o] = b
                    new_rows.append(nr)
        rows = new_rows
    out = np.array([[r[v] for v in required] for r in rows], dtype=np.uint32).reshape(-1, len(required))
    return sort_rows(out)


def sort_rows(a):
    a = np.asarray(a, dtype=np.uint32)
    if a.size == 0:
        return a.reshape(0, a.shape[1] if a.ndim == 2 else 0)
    idx = np.lexsort(a.T[::-1])
    return a[idx]


def py_final_process(full, req_cols, distinct, offset, limit):
    """independent statement of final_process' modifiers (sparql.hpp:1424-1550) on a full binding table"""
    t = np.asarray(full, dtype=np.uint32)
    if t.shape[0] == 0:
        return t[:, req_cols]
    if distinct:
        s = t.view(np.int32)
        order = np.lexsort(tuple(s[:, c] for c in range(t.shape[1] - 1, -1, -1)))   # all columns, signed, col 0 first
        t = t[order]
        key = t[:, req_cols]
        keep = np.ones(t.shape[0], dtype=bool)
        keep[1:] = np.any(key[1:] != key[:-1], axis=1)
        t = t[keep]
    if offset > 0:
        t = t[offset:]
    if limit >= 0:
        t = t[:limit]
    return t[:, req_cols]
