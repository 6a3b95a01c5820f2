"""ctypes binding of the CPU oracle (oracle/wukong_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (wukong_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwukong_oracle.so")

IN, OUT = 0, 1
PREDICATE_ID, TYPE_ID = 0, 1
I2U, C2U, K2U, K2K, K2C = 0, 1, 2, 3, 4
C2K, I2K = 6, 7


class SegMeta(C.Structure):
    _fields_ = [("index", C.c_int32), ("dir", C.c_int32), ("pid", C.c_uint32), ("_pad", C.c_uint32),
                ("num_keys", C.c_uint64), ("num_buckets", C.c_uint64), ("bucket_start", C.c_uint64),
                ("num_edges", C.c_uint64), ("edge_start", C.c_uint64), ("ext_start", C.c_uint64),
                ("ext_num", C.c_uint64)]


def build(force=False):
    src = os.path.join(_HERE, "wukong_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    u64, u32, i32, vp = C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p
    L.wko_store_build.restype = vp
    L.wko_store_build.argtypes = [vp, u64, C.c_int, C.c_int, C.c_int, u64, C.c_int, C.c_int]
    L.wko_store_wrap.restype = vp
    L.wko_store_wrap.argtypes = [C.c_int, C.c_int, vp, u64, vp, u64, vp, C.c_int]
    L.wko_store_free.argtypes = [vp]
    L.wko_store_ok.argtypes = [vp]
    L.wko_store_error.restype = C.c_char_p
    L.wko_store_error.argtypes = [vp]
    for f in ("wko_store_vertices", "wko_store_edges"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    for f in ("wko_store_num_slots", "wko_store_num_buckets", "wko_store_num_entries",
              "wko_store_used_entries", "wko_store_used_ext", "wko_store_check"):
        getattr(L, f).restype = u64
        getattr(L, f).argtypes = [vp]
    L.wko_store_num_segs.argtypes = [vp]
    L.wko_store_segs.argtypes = [vp, vp]
    L.wko_get_edges.restype = u64
    L.wko_get_edges.argtypes = [vp, u32, u32, C.c_int, C.POINTER(vp)]
    L.wko_hash_u64.restype = u64
    L.wko_hash_u64.argtypes = [u64]
    L.wko_hash_prime_u64.restype = u64
    L.wko_hash_prime_u64.argtypes = [u64]
    L.wko_make_ptr.restype = u64
    L.wko_make_ptr.argtypes = [u64, u64]
    L.wko_is_tpid.argtypes = [C.c_int64]
    L.wko_less_pso.argtypes = [vp, vp]
    L.wko_less_pos.argtypes = [vp, vp]
    L.wko_make_key.restype = u64
    L.wko_make_key.argtypes = [u64, u64, u64]
    L.wko_set_plan.argtypes = [vp, C.c_int, C.c_char_p, vp, C.c_int]
    L.wko_query_run.restype = vp
    L.wko_query_run.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.wko_query_run_ex.restype = vp
    L.wko_query_run_ex.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int64, C.c_int64]
    L.wko_result_free.argtypes = [vp]
    L.wko_result_status.argtypes = [vp]
    L.wko_result_rows.restype = u64
    L.wko_result_rows.argtypes = [vp]
    L.wko_result_cols.argtypes = [vp]
    L.wko_result_table.restype = vp
    L.wko_result_table.argtypes = [vp]
    L.wko_result_table_len.restype = u64
    L.wko_result_table_len.argtypes = [vp]
    L.wko_result_usec.restype = C.c_double
    L.wko_result_usec.argtypes = [vp]
    L.wko_run_primitive.restype = C.c_int64
    L.wko_run_primitive.argtypes = [vp, C.c_int, vp, u64, C.c_int, i32, i32, C.c_int, i32, C.c_int, C.c_int,
                                    vp, u64, C.POINTER(C.c_int)]
    L.wko_emu_run.restype = C.c_double
    L.wko_emu_run.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Store:
    """One server's graph store (reference GStore) built by the oracle, or a wrapped foreign one."""

    def __init__(self, handle, keep=()):
        self.h = handle
        self._keep = keep

    @classmethod
    def build(cls, triples, num_servers=1, sid=0, num_engines=1, kvstore_bytes=64 << 20,
              num_normal_preds=31, gpu_ext_mode=True):
        t = np.ascontiguousarray(triples, dtype=np.uint32).reshape(-1, 3)
        h = lib().wko_store_build(_ptr(t), t.shape[0], num_servers, sid, num_engines, kvstore_bytes,
                                  num_normal_preds, 1 if gpu_ext_mode else 0)
        s = cls(h)
        if not lib().wko_store_ok(h):
            raise RuntimeError("oracle store build failed: " + lib().wko_store_error(h).decode())
        return s

    @classmethod
    def wrap(cls, vertices, edges, segs, num_servers=1, sid=0):
        v = np.ascontiguousarray(vertices, dtype=np.uint64).reshape(-1, 2)
        e = np.ascontiguousarray(edges, dtype=np.uint32)
        sa = (SegMeta * len(segs))()
        for i, sg in enumerate(segs):   # accepts any struct with the same field names
            for f, _t in SegMeta._fields_:
                setattr(sa[i], f, getattr(sg, f))
        h = lib().wko_store_wrap(num_servers, sid, _ptr(v), v.shape[0], _ptr(e), e.shape[0],
                                 C.cast(sa, C.c_void_p), len(segs))
        return cls(h, keep=(v, e, sa))

    def __del__(self):
        try:
            if self.h:
                lib().wko_store_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def num_slots(self):
        return lib().wko_store_num_slots(self.h)

    @property
    def num_buckets(self):
        return lib().wko_store_num_buckets(self.h)

    @property
    def used_entries(self):
        return lib().wko_store_used_entries(self.h)

    @property
    def used_ext(self):
        return lib().wko_store_used_ext(self.h)

    def vertices(self):
        """(num_slots, 2) uint64 view: [:,0] = raw ikey_t bits, [:,1] = raw iptr_t bits."""
        n = self.num_slots
        buf = (C.c_uint64 * (2 * n)).from_address(lib().wko_store_vertices(self.h))
        return np.frombuffer(buf, dtype=np.uint64).reshape(n, 2)

    def edges(self, used_only=True):
        n = self.used_entries if used_only else lib().wko_store_num_entries(self.h)
        n = max(int(n), 1)
        buf = (C.c_uint32 * n).from_address(lib().wko_store_edges(self.h))
        return np.frombuffer(buf, dtype=np.uint32)

    def segs(self):
        n = lib().wko_store_num_segs(self.h)
        arr = (SegMeta * n)()
        lib().wko_store_segs(self.h, C.cast(arr, C.c_void_p))
        return list(arr)

    def check(self):
        return lib().wko_store_check(self.h)

    def get_edges(self, vid, pid, d):
        p = C.c_void_p()
        n = lib().wko_get_edges(self.h, vid, pid, d, C.byref(p))
        if n == 0 or not p.value:
            return np.zeros(0, dtype=np.uint32)
        return np.frombuffer((C.c_uint32 * n).from_address(p.value), dtype=np.uint32).copy()

    def primitive(self, kind, table, ncols, a_start, pid, d, a_end=0, mt_tid=0, mt_factor=1, cap_words=None):
        """Run exactly one reference pattern function.  Returns (rows x cols) uint32 array."""
        if table is None:
            tbl, nrows = None, 0
        else:
            tbl = np.ascontiguousarray(table, dtype=np.uint32).reshape(-1, max(ncols, 1))
            nrows = tbl.shape[0] if ncols else 0
        cap = cap_words or (1 << 24)
        while True:
            out = np.empty(cap, dtype=np.uint32)
            oc = C.c_int(0)
            r = lib().wko_run_primitive(self.h, kind, _ptr(tbl), nrows, ncols, a_start, pid, d, a_end,
                                        mt_tid, mt_factor, _ptr(out), cap, C.byref(oc))
            if r < 0:
                raise RuntimeError("oracle primitive failed: %d" % r)
            if r * oc.value <= cap:
                return out[: r * oc.value].reshape(r, oc.value).copy()
            cap = r * oc.value


def set_plan(patterns, fmt_text):
    """patterns: list of (s, p, dir, o).  Applies planner.hpp:1647-1754 semantics."""
    a = np.array(patterns, dtype=np.int32).reshape(-1, 4)
    out = np.zeros((64, 4), dtype=np.int32)
    n = lib().wko_set_plan(_ptr(a), a.shape[0], fmt_text.encode(), _ptr(out), 64)
    if n < 0:
        raise ValueError("bad plan")
    return [tuple(int(x) for x in row) for row in out[:n]]


class QueryResult:
    def __init__(self, status, rows, cols, table, usec):
        self.status, self.rows, self.cols, self.table, self.usec = status, rows, cols, table, usec


def run_query(stores, patterns, nvars, required_vars, mt_factor=1, blind=False, threaded=False,
              distinct=False, offset=0, limit=-1):
    arr = (C.c_void_p * len(stores))(*[s.h for s in stores])
    p = np.array(patterns, dtype=np.int32).reshape(-1, 4)
    rv = np.array(required_vars, dtype=np.int32)
    h = lib().wko_query_run_ex(C.cast(arr, C.c_void_p), len(stores), _ptr(p), p.shape[0], nvars, _ptr(rv), len(rv),
                               mt_factor, 1 if blind else 0, 1 if threaded else 0, 1 if distinct else 0, offset, limit)
    try:
        status = lib().wko_result_status(h)
        rows = lib().wko_result_rows(h)
        cols = lib().wko_result_cols(h)
        n = lib().wko_result_table_len(h)
        if n:
            tbl = np.frombuffer((C.c_uint32 * n).from_address(lib().wko_result_table(h)), dtype=np.uint32).copy()
            tbl = tbl.reshape(-1, cols)
        else:
            tbl = np.zeros((0, max(cols, 1)), dtype=np.uint32)
        return QueryResult(status, rows, cols, tbl, lib().wko_result_usec(h))
    finally:
        lib().wko_result_free(h)


def emu_run(store, pats, off, nv, nthreads):
    """closed-loop emulator in native code: -> (seconds, rows[])"""
    rows = np.zeros(len(nv), dtype=np.uint64)
    sec = lib().wko_emu_run(store.h, _ptr(pats), _ptr(off), _ptr(nv), len(nv), nthreads, _ptr(rows))
    return sec, rows
