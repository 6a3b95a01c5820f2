"""Seeded random RDF graphs and random planned basic graph patterns (shared by the CPU and GPU randomized tests).

Id space like the reference's str_index (base_loader.hpp:409-424): 0 __PREDICATE__, 1 rdf:type, predicates 2..P+1,
types P+2..P+T+1 (all < 2^17), vertices from 2^17.  A query is returned twice: `planned` in the engine's form
(start, predicate, direction, end) including index seeds, and `semantic` (s, p, _, o) triples for the brute-force joiner."""
import numpy as np

IN, OUT = 0, 1
PREDICATE_ID, TYPE_ID = 0, 1
V0 = 1 << 17


def graph(seed, nv=300, npred=4, ntype=3, ntriples=2500):
    rng = np.random.default_rng(seed)
    preds = np.arange(2, 2 + npred)
    types = np.arange(2 + npred, 2 + npred + ntype)
    verts = V0 + rng.permutation(4 * nv)[:nv].astype(np.uint32)          # sparse, unordered ids
    s = verts[np.minimum(rng.zipf(1.6, ntriples) - 1, nv - 1)]           # skewed subjects: a few hubs
    o = verts[rng.integers(0, nv, ntriples)]
    p = preds[rng.integers(0, npred, ntriples)]
    tr = np.stack([s, p, o], axis=1)
    # a hub on the object side, self loops, duplicates, and vertices with several types / none
    hub = verts[0]
    tr = np.concatenate([tr, np.stack([verts[rng.integers(0, nv, 200)], np.full(200, preds[0]), np.full(200, hub)], axis=1)])
    tr = np.concatenate([tr, np.stack([verts[:10], np.full(10, preds[1]), verts[:10]], axis=1)])
    typed = verts[rng.integers(0, nv, int(1.3 * nv))]
    tt = np.stack([typed, np.full(typed.size, TYPE_ID), types[rng.integers(0, ntype, typed.size)]], axis=1)
    tr = np.concatenate([tr, tt, tr[rng.integers(0, tr.shape[0], 150)]]).astype(np.uint32)
    tr = tr[rng.permutation(tr.shape[0])]
    return tr, dict(preds=preds, types=types, verts=verts, num_normal_preds=int(1 + npred + ntype))


def query(seed, tr, meta, max_steps=4):
    """-> (planned, semantic, nvars, required)"""
    rng = np.random.default_rng(seed)
    preds, types, verts = meta["preds"], meta["types"], meta["verts"]
    planned, semantic = [], []
    bound = []                      # variables in binding order
    type_vars = set()               # variables that hold type ids

    def new_var():
        v = -(len(bound) + 1)
        bound.append(v)
        return v

    kind = rng.integers(0, 4)
    if kind == 0:                                             # type-index seed
        t = int(types[rng.integers(0, len(types))])
        a = new_var()
        planned.append((t, TYPE_ID, IN, a))
        semantic.append((a, TYPE_ID, 0, t))
    elif kind == 1:                                           # predicate-index seed, subject side ("<<")
        p = int(preds[rng.integers(0, len(preds))])
        a = new_var()
        b = new_var()
        planned += [(p, PREDICATE_ID, IN, a), (a, p, OUT, b)]
        semantic.append((a, p, 0, b))
    elif kind == 2:                                           # predicate-index seed, object side (">>")
        p = int(preds[rng.integers(0, len(preds))])
        a = new_var()
        b = new_var()
        planned += [(p, PREDICATE_ID, OUT, a), (a, p, IN, b)]
        semantic.append((b, p, 0, a))
    else:                                                     # constant start
        row = tr[rng.integers(0, tr.shape[0])]
        a = new_var()
        if rng.integers(0, 2):
            planned.append((int(row[0]), int(row[1]), OUT, a))
            semantic.append((int(row[0]), int(row[1]), 0, a))
            if row[1] == TYPE_ID:
                type_vars.add(a)
        else:
            if row[1] == TYPE_ID:                             # [t|type|IN] is not a normal key: use a real edge instead
                row = tr[tr[:, 1] != TYPE_ID][0]
            planned.append((int(row[2]), int(row[1]), IN, a))
            semantic.append((a, int(row[1]), 0, int(row[2])))
    for _ in range(int(rng.integers(1, max_steps + 1))):
        cands = [v for v in bound if v not in type_vars]
        if not cands:
            break
        x = cands[rng.integers(0, len(cands))]
        r = rng.random()
        if r < 0.15:                                          # type check / types of x
            if rng.integers(0, 2):
                t = int(types[rng.integers(0, len(types))])
                planned.append((x, TYPE_ID, OUT, t))
                semantic.append((x, TYPE_ID, 0, t))
            else:
                tv = new_var()
                type_vars.add(tv)
                planned.append((x, TYPE_ID, OUT, tv))
                semantic.append((x, TYPE_ID, 0, tv))
            continue
        p = int(preds[rng.integers(0, len(preds))])
        d = int(rng.integers(0, 2))
        others = [v for v in bound if v != x and v not in type_vars]
        if r < 0.30 and others:                               # known_to_known
            y = others[rng.integers(0, len(others))]
            end = y
        elif r < 0.45:                                        # known_to_const (mostly an existing neighbour: some hits)
            end = int(verts[rng.integers(0, 20)])
        else:
            end = new_var()
        planned.append((x, p, d, end))
        semantic.append((x, p, 0, end) if d == OUT else (end, p, 0, x))
    k = int(rng.integers(1, len(bound) + 1))
    required = [bound[i] for i in sorted(rng.permutation(len(bound))[:k].tolist())]
    return planned, semantic, len(bound), required
