"""CPU: the wire codec of a SPARQL request / reply (csrc/host/bundle.hpp = the reference's Bundle + Boost binary archive of a
SPARQLQuery, core/query.hpp:917-1232) against an independent Python statement of the archive layout, byte for byte, and
decode(encode(x)) == x.  Not pinned against a real Boost build (none in this image; said so in bundle.hpp and DESIGN.md)."""
import ctypes as C
import struct

import numpy as np

from conftest import PLANS, load_query
from wukong_b200 import host

OCCUPIED, EMPTY = b"\x00", b"\x01"


def _lib():
    L = host.lib()
    L.wkh_bundle_encode.restype = C.c_int64
    L.wkh_bundle_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    L.wkh_bundle_decode.restype = C.c_int64
    L.wkh_bundle_decode.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
    L.wkh_bundle_roundtrip_query.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    return L


# ---- the query as nested Python data --------------------------------------------------------------------------------------
def group(pats, newvars=(), unions=(), optionals=()):
    return {"pats": [tuple(p) for p in pats], "newvars": sorted(set(newvars)), "unions": list(unions), "optionals": list(optionals)}


def flat_group(g):
    out = [len(g["pats"])]
    for p in g["pats"]:
        out += list(p)
    out += [len(g["newvars"])] + list(g["newvars"])
    out.append(len(g["unions"]))
    for u in g["unions"]:
        out += flat_group(u)
    out.append(len(g["optionals"]))
    for o in g["optionals"]:
        out += flat_group(o)
    return out


def flat_query(q):
    out = list(q["scalars"]) + flat_group(q["group"]) + [len(q["orders"])]
    for o in q["orders"]:
        out += list(o)
    out += list(q["result"])
    for k in ("required", "v2c", "matched", "table"):
        out += [len(q[k])] + [int(x) for x in q[k]]
    out += list(q["gpu"])
    return out


# ---- the archive layout, restated here without looking at bundle.hpp's code --------------------------------------------------
def ar_header():
    sig = b"serialization::archive"
    return struct.pack("<Q", len(sig)) + sig + struct.pack("<H", 16) + bytes([4, 8, 4, 8]) + struct.pack("<i", 1)


def ar_group(g):
    b = struct.pack("<QI", len(g["pats"]), 0)
    for (s, p, d, o, t) in g["pats"]:
        b += struct.pack("<iiiib", s, p, o, d, t)                     # save order: subject, predicate, OBJECT, direction, pred_type
    b += struct.pack("<QI", len(g["newvars"]), 0) + b"".join(struct.pack("<i", v) for v in g["newvars"])
    b += EMPTY                                                        # filters
    if g["optionals"]:
        b += OCCUPIED + struct.pack("<QI", len(g["optionals"]), 0) + b"".join(ar_group(x) for x in g["optionals"])
    else:
        b += EMPTY
    if g["unions"]:
        b += OCCUPIED + struct.pack("<QI", len(g["unions"]), 0) + b"".join(ar_group(x) for x in g["unions"])
    else:
        b += EMPTY
    return b


def ar_query(q, gpu_build=False):
    s = q["scalars"]
    b = ar_header()
    b += struct.pack("<11i", *s[:11]) + struct.pack("<?ii?i", bool(s[11]), s[12], s[13], bool(s[14]), s[15])
    b += struct.pack("<iI?", s[16], s[17], bool(s[18]))
    b += ar_group(q["group"])
    if q["orders"]:
        b += OCCUPIED + struct.pack("<QI", len(q["orders"]), 0) + b"".join(struct.pack("<i?", i, bool(d)) for i, d in q["orders"])
    else:
        b += EMPTY
    r = q["result"]
    b += struct.pack("<iiii?i", r[0], r[1], r[2], r[3], bool(r[4]), r[5])
    b += struct.pack("<Q", len(q["required"])) + np.array(q["required"], dtype="<i4").tobytes()
    b += struct.pack("<Q", len(q["v2c"])) + np.array(q["v2c"], dtype="<i4").tobytes()
    b += struct.pack("<Q", len(q["matched"])) + bytes(1 if x else 0 for x in q["matched"])
    if r[1] > 0:                                                      # row_num > 0: the tables are in the archive
        b += OCCUPIED + struct.pack("<Q", len(q["table"])) + np.array(q["table"], dtype="<u4").tobytes() + struct.pack("<QI", 0, 0)
    else:
        b += EMPTY
    if gpu_build:
        b += struct.pack("<Qi", q["gpu"][0], q["gpu"][1])
    return struct.pack("<i", 0) + b                                   # Bundle::to_str: req_type SPARQL_QUERY first


def sample_queries():
    rng = np.random.default_rng(3)
    out = []
    # a request as the proxy sends it: planned Q7, nothing bound yet
    pats, nvars, req, _ = load_query(7, PLANS[0])
    out.append({"scalars": [-1, 12, 0, 0, 1, 0, 0, 4, 1, 0, 0, 0, 0, 0, 0, 0, -1, 0, 0],
                "group": group([(s, p, d, o, 0) for (s, p, d, o) in pats]), "orders": [],
                "result": [0, 0, 0, 0, 1, nvars], "required": req, "v2c": [0xFFFF] * nvars, "matched": [], "table": [], "gpu": [0, 0]})
    # a reply with a table, modifiers and an ORDER BY
    tbl = rng.integers(1 << 17, 1 << 30, 3 * 5).tolist()
    out.append({"scalars": [1027, 3, 0, 5, 0, 0, 2, 1, 0, 6, -2, 1, 3, 4, 1, 2, 10, 7, 1],
                "group": group([]), "orders": [(-1, 1), (-3, 0)],
                "result": [3, 5, 0, 0, 0, 3], "required": [-1, -2, -3], "v2c": [0, 1, 2], "matched": [1, 0, 1, 1, 0], "table": tbl, "gpu": [15, 3]})
    # unions and optionals, nested, with optional_new_vars
    g = group([(-1, 9, 0, 131077, 0)],
              unions=[group([(-1, 1, 1, 22, 0), (-1, 8, 1, -2, 0)], optionals=[group([(-2, 8, 1, -5, 0)], newvars=[-5])]),
                      group([(23, 1, 0, -1, 0), (-1, 8, 1, -2, 0)])],
              optionals=[group([(-1, 12, 1, -3, 0), (-3, 8, 1, 140000, 0)], newvars=[-3]), group([(-1, 13, 1, -4, 0)], newvars=[-4])])
    out.append({"scalars": [5, 5, 2, 3, 1, 1, 0, 1, 0, 1, -1, 0, 0, 0, 0, 1, -1, 0, 0],
                "group": g, "orders": [], "result": [2, 0, 0, 6, 0, 5], "required": [-1, -2], "v2c": [0, 1, 0xFFFF, 0xFFFF, 0xFFFF],
                "matched": [], "table": [], "gpu": [0, 0]})
    return out


def test_archive_bytes_match_the_restated_layout():
    L = _lib()
    for q in sample_queries():
        for gpu_build in (False, True):
            flat = np.array(flat_query(q), dtype=np.int64)
            buf = np.zeros(1 << 16, dtype=np.uint8)
            n = L.wkh_bundle_encode(flat.ctypes.data_as(C.c_void_p), flat.size, 1 if gpu_build else 0, buf.ctypes.data_as(C.c_void_p), buf.size)
            assert n > 0
            got = bytes(buf[:n])
            want = ar_query(q, gpu_build)
            assert got == want, (len(got), len(want), next(i for i in range(min(len(got), len(want))) if got[i] != want[i]))
            # and back
            out = np.zeros(1 << 14, dtype=np.int64)
            m = L.wkh_bundle_decode(buf.ctypes.data_as(C.c_void_p), n, 1 if gpu_build else 0, out.ctypes.data_as(C.c_void_p), out.size)
            exp = flat_query(q)
            if not gpu_build:
                exp = exp[:-2] + [0, 0]
            if q["result"][1] <= 0:
                pass
            assert m == len(exp) and out[:m].tolist() == exp


def test_header_is_what_boost_writes():
    # 8-byte length 22, the signature, archive version 16 (Boost 1.66 / 1.67), sizeof(int, long, float, double), int 1
    h = ar_header()
    assert h[:8] == (22).to_bytes(8, "little") and h[8:30] == b"serialization::archive" and h[30:32] == b"\x10\x00"
    assert h[32:36] == bytes([4, 8, 4, 8]) and h[36:40] == b"\x01\x00\x00\x00" and len(h) == 40


def test_malformed_archives_are_refused():
    L = _lib()
    q = sample_queries()[1]
    good = ar_query(q)
    out = np.zeros(1 << 14, dtype=np.int64)

    def dec(b):
        a = np.frombuffer(b, dtype=np.uint8).copy()
        return L.wkh_bundle_decode(a.ctypes.data_as(C.c_void_p), a.size, 0, out.ctypes.data_as(C.c_void_p), out.size)

    assert dec(good) > 0
    assert dec(good[:-3]) == -1                                    # truncated
    assert dec(good + b"\x00") == -1                               # trailing bytes
    assert dec(struct.pack("<i", 1) + good[4:]) == -1              # DYNAMIC_LOAD bundle, not a query
    bad_sig = bytearray(good); bad_sig[12 + 3] ^= 0x20
    assert dec(bytes(bad_sig)) == -1
    # a FILTER in the pattern group is refused, not misread: flip the group's "filters: empty" marker
    g0 = len(struct.pack("<i", 0)) + len(ar_header()) + struct.calcsize("<11i") + struct.calcsize("<?ii?i") + struct.calcsize("<iI?")
    marker = g0 + 12 + 12          # empty pattern vector (count + item version), empty new-vars set, then the marker byte
    b = bytearray(good)
    assert b[marker] == 1
    b[marker] = 0
    assert dec(bytes(b)) == -1


def test_host_query_survives_the_wire():
    L = _lib()
    for q in (2, 4, 7):
        pats, nvars, req, _ = load_query(q, PLANS[1])
        p = np.array(pats, dtype=np.int32)
        r = np.array(req, dtype=np.int32)
        tbl = np.arange(4 * len(req), dtype=np.uint32) + 131072
        assert L.wkh_bundle_roundtrip_query(p.ctypes.data_as(C.c_void_p), len(pats), nvars, r.ctypes.data_as(C.c_void_p), len(req), 0,
                                            tbl.ctypes.data_as(C.c_void_p), 4, len(req)) == 0
        assert L.wkh_bundle_roundtrip_query(p.ctypes.data_as(C.c_void_p), len(pats), nvars, r.ctypes.data_as(C.c_void_p), len(req), 1,
                                            None, 4, len(req)) == 0
