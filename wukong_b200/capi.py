"""ctypes binding of the C ABI declared in include/wukong_b200.h (libwukong_b200.so)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_CUDA = os.path.join(_HERE, "libwukong_b200.so")

IN, OUT = 0, 1
PREDICATE_ID, TYPE_ID = 0, 1
KIND_NAMES = ["i2u", "c2u", "k2u", "k2k", "k2c", "project", "c2k", "i2k", "distinct", "slice", "exchange", "filter"]

WK_SUCCESS = 0
WK_ERR_CUDA, WK_ERR_BAD_ARG, WK_ERR_RBUF_OVERFLOW, WK_ERR_NO_SEGMENT, WK_ERR_NO_DEVICE, WK_ERR_COMM = 100, 101, 102, 103, 104, 105
WK_ERR_STORE_FULL = 106
# wk_engine_set_option / wk_engine_get_option
WK_OPT_RESIDENT_LIGHT, WK_OPT_RESIDENT_IDLE_US, WK_OPT_FUSE_FILTERS, WK_OPT_DIRECT_OUT, WK_OPT_RESIDENT_VARIANT = 1, 2, 3, 4, 5
WK_INFO_RESIDENT_LAUNCHES, WK_INFO_RESIDENT_REQUESTS, WK_INFO_LAST_RESIDENT, WK_INFO_LAST_RESIDENT_NS, WK_INFO_RESIDENT_RUNNING = 100, 101, 102, 103, 104
WK_INFO_COMM_BYTES_PUSHED = 110


class WukongError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        msg = lib().wk_strerror(code).decode() if _lib is not None else "?"
        super().__init__("wukong_b200 status %d (%s) %s" % (code, msg, what))


class SegMeta(C.Structure):
    _fields_ = [("index", C.c_int32), ("dir", C.c_int32), ("pid", C.c_uint32), ("_pad", C.c_uint32),
                ("num_keys", C.c_uint64), ("num_buckets", C.c_uint64), ("bucket_start", C.c_uint64),
                ("num_edges", C.c_uint64), ("edge_start", C.c_uint64), ("ext_start", C.c_uint64),
                ("ext_num", C.c_uint64)]


class BuildOpts(C.Structure):
    _fields_ = [("num_servers", C.c_int32), ("sid", C.c_int32), ("num_normal_preds", C.c_int32),
                ("est_load_factor", C.c_int32), ("kvstore_bytes", C.c_uint64), ("triples_on_device", C.c_int32),
                ("_pad", C.c_int32)]


class BuildStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("num_keys", "num_triples_out", "num_triples_in", "num_buckets",
                                          "num_buckets_ext", "used_ext", "num_slots", "num_edges")] + \
               [(n, C.c_float) for n in ("ms_upload", "ms_sort", "ms_insert", "ms_total")]


class QueryOpts(C.Structure):
    _fields_ = [("mt_tid", C.c_int32), ("mt_factor", C.c_int32), ("blind", C.c_int32), ("distinct", C.c_int32),
                ("offset", C.c_int64), ("limit", C.c_int64)]


class StepStats(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_cols", C.c_int32), ("in_rows", C.c_uint64), ("out_rows", C.c_uint64),
                ("buckets_visited", C.c_uint64), ("edges_touched", C.c_uint64), ("algo_bytes", C.c_uint64),
                ("device_us", C.c_float), ("launches", C.c_int32)]


# every symbol include/wukong_b200.h declares (checked by tests/test_capi_symbols.py)
DECLARED_SYMBOLS = [
    "wk_strerror", "wk_version", "wk_device_count", "wk_store_create", "wk_store_adopt", "wk_store_build", "wk_store_info", "wk_store_segs",
    "wk_store_download", "wk_store_destroy",
    "wk_store_get_edges", "wk_engine_create", "wk_engine_destroy", "wk_engine_set_option", "wk_engine_get_option", "wk_engine_set_profiling", "wk_engine_light_trace", "wk_engine_sync",
    "wk_engine_reset", "wk_table_upload", "wk_table_download", "wk_table_info", "wk_index_to_unknown",
    "wk_const_to_unknown", "wk_known_to_unknown", "wk_known_to_known", "wk_known_to_const", "wk_const_to_known", "wk_index_to_known", "wk_table_distinct", "wk_table_slice", "wk_project",
    "wk_query_execute_ex",
    "wk_query_execute", "wk_query_execute_batch", "wk_engine_num_steps", "wk_engine_step_stats", "wk_engine_launch_count", "wk_engine_last_query_device_us", "wk_engine_flush_l2", "wk_host_alloc", "wk_host_free", "wk_partition",
    "wk_partition_ptr", "wk_comm_unique_id", "wk_comm_init", "wk_exchange", "wk_query_execute_sharded",
    "wk_comm_stats", "wk_plan_exchanges", "wk_comm_p2p_export", "wk_comm_p2p_import", "wk_exchange_p2p",
    "wk_comm_p2p_export_store", "wk_comm_p2p_import_store", "wk_comm_local_group",
]

_lib = None


def lib():
    """Load libwukong_b200.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_CUDA):
        raise RuntimeError("libwukong_b200.so is missing: run `python -m wukong_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_CUDA, mode=C.RTLD_GLOBAL)
    u64, u32, i32, vp, ci = C.c_uint64, C.c_uint32, C.c_int32, C.c_void_p, C.c_int
    pu64 = C.POINTER(u64)
    L.wk_strerror.restype = C.c_char_p
    L.wk_strerror.argtypes = [ci]
    L.wk_device_count.argtypes = [C.POINTER(ci)]
    L.wk_store_create.argtypes = [ci, vp, u64, vp, u64, vp, ci, C.POINTER(vp)]
    L.wk_store_adopt.argtypes = [ci, vp, u64, vp, u64, vp, ci, ci, C.POINTER(vp)]
    L.wk_store_destroy.argtypes = [vp]
    L.wk_store_build.argtypes = [ci, vp, u64, vp, C.POINTER(vp), vp]
    L.wk_store_info.argtypes = [vp, pu64, pu64, C.POINTER(ci)]
    L.wk_store_segs.argtypes = [vp, vp, ci]
    L.wk_store_download.argtypes = [vp, vp, u64, vp, u64]
    L.wk_store_get_edges.argtypes = [vp, u32, u32, ci, vp, u64, pu64]
    L.wk_engine_create.argtypes = [vp, u64, C.POINTER(vp)]
    L.wk_engine_destroy.argtypes = [vp]
    L.wk_engine_set_profiling.argtypes = [vp, ci]
    L.wk_engine_set_option.argtypes = [vp, ci, C.c_int64]
    L.wk_engine_get_option.argtypes = [vp, ci, C.POINTER(C.c_int64)]
    L.wk_engine_sync.argtypes = [vp]
    L.wk_engine_reset.argtypes = [vp]
    L.wk_table_upload.argtypes = [vp, vp, u64, ci]
    L.wk_table_download.argtypes = [vp, vp, u64, pu64, C.POINTER(ci)]
    L.wk_table_info.argtypes = [vp, pu64, C.POINTER(ci)]
    L.wk_index_to_unknown.argtypes = [vp, u32, ci, ci, ci, pu64]
    L.wk_const_to_unknown.argtypes = [vp, u32, u32, ci, pu64]
    L.wk_known_to_unknown.argtypes = [vp, ci, u32, ci, pu64]
    L.wk_known_to_known.argtypes = [vp, ci, u32, ci, ci, pu64]
    L.wk_known_to_const.argtypes = [vp, ci, u32, ci, u32, pu64]
    L.wk_project.argtypes = [vp, vp, ci, pu64]
    L.wk_table_distinct.argtypes = [vp, vp, ci, pu64]
    L.wk_table_slice.argtypes = [vp, u64, C.c_int64, pu64]
    L.wk_query_execute_ex.argtypes = [vp, vp, ci, ci, vp, ci, vp, vp, u64, pu64, C.POINTER(ci)]
    L.wk_const_to_known.argtypes = [vp, u32, u32, ci, ci, pu64]
    L.wk_index_to_known.argtypes = [vp, u32, ci, ci, ci, ci, pu64]
    L.wk_query_execute.argtypes = [vp, vp, ci, ci, vp, ci, ci, ci, ci, vp, u64, pu64, C.POINTER(ci)]
    L.wk_query_execute_batch.argtypes = [vp, vp, vp, vp, ci, vp, vp]
    L.wk_engine_num_steps.argtypes = [vp]
    L.wk_engine_step_stats.argtypes = [vp, ci, C.POINTER(StepStats)]
    L.wk_engine_launch_count.restype = u64
    L.wk_engine_launch_count.argtypes = [vp]
    L.wk_engine_last_query_device_us.argtypes = [vp, C.POINTER(C.c_float)]
    L.wk_engine_flush_l2.argtypes = [vp]
    L.wk_partition.argtypes = [vp, ci, ci, vp]
    L.wk_partition_ptr.argtypes = [vp, ci, C.POINTER(vp), pu64]
    L.wk_comm_unique_id.argtypes = [vp]
    L.wk_comm_init.argtypes = [vp, ci, ci, vp]
    L.wk_exchange.argtypes = [vp, ci, pu64]
    L.wk_comm_stats.argtypes = [vp, pu64, pu64, pu64]
    L.wk_comm_p2p_export.argtypes = [vp, ci, ci, vp]
    L.wk_comm_p2p_import.argtypes = [vp, vp]
    L.wk_exchange_p2p.argtypes = [vp, ci, pu64]
    L.wk_comm_p2p_export_store.argtypes = [vp, vp, u64, pu64]
    L.wk_comm_p2p_import_store.argtypes = [vp, vp, vp, ci]
    L.wk_plan_exchanges.argtypes = [vp, ci, ci, vp]
    L.wk_comm_local_group.argtypes = [vp, ci]
    L.wk_query_execute_sharded.argtypes = [vp, vp, ci, ci, vp, ci, ci, ci, ci, vp, u64, pu64, C.POINTER(ci)]
    L.wk_host_alloc.argtypes = [u64, C.POINTER(vp)]
    L.wk_host_free.argtypes = [vp]
    L.wk_selftest_hash.restype = u64
    L.wk_selftest_hash.argtypes = [u64]
    L.wk_selftest_fastmod.restype = u64
    L.wk_selftest_fastmod.argtypes = [u64, u64]
    L.wk_engine_light_trace.argtypes = [vp, vp, ci]
    L.wk_selftest_ptr_size.restype = u64
    L.wk_selftest_ptr_size.argtypes = [u64]
    L.wk_selftest_ptr_off.restype = u64
    L.wk_selftest_ptr_off.argtypes = [u64]
    L.wk_selftest_make_key.restype = u64
    L.wk_selftest_make_key.argtypes = [u64, u32, u32]
    L.wk_selftest_owner.restype = u64
    L.wk_selftest_owner.argtypes = [u64, u64]
    _lib = L
    return L


def _check(rc, what=""):
    if rc != 0:
        raise WukongError(rc, what)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    n = C.c_int(0)
    lib().wk_device_count(C.byref(n))
    return n.value


class Store:
    """GPU-resident graph store (wk_store_t)."""

    def __init__(self, vertices, edges, segs, device=0):
        v = np.ascontiguousarray(vertices, dtype=np.uint64).reshape(-1, 2)
        e = np.ascontiguousarray(edges, dtype=np.uint32)
        sa = (SegMeta * len(segs))()
        for i, s in enumerate(segs):
            for f, _t in SegMeta._fields_:
                setattr(sa[i], f, getattr(s, f))
        h = C.c_void_p()
        _check(lib().wk_store_create(device, _ptr(v), v.shape[0], _ptr(e), e.shape[0], C.cast(sa, C.c_void_p),
                                     len(segs), C.byref(h)), "wk_store_create")
        self.h = h
        self.device = device

    @classmethod
    def adopt(cls, d_vertices, num_slots, d_edges, num_edges, segs, device=0, take_ownership=True):
        self = cls.__new__(cls)
        sa = (SegMeta * len(segs))(*segs)
        h = C.c_void_p()
        _check(lib().wk_store_adopt(device, d_vertices, num_slots, d_edges, num_edges, C.cast(sa, C.c_void_p),
                                    len(segs), 1 if take_ownership else 0, C.byref(h)), "wk_store_adopt")
        self.h = h
        self.device = device
        return self

    @classmethod
    def build(cls, triples, num_normal_preds, num_servers=1, sid=0, est_load_factor=55, kvstore_bytes=0, device=0):
        """wk_store_build: sort / dedup / partition + hash-table construction on the device.
        Returns the store; .build_stats holds the counters and timings of the build."""
        t = np.ascontiguousarray(triples, dtype=np.uint32).reshape(-1, 3)
        o = BuildOpts(num_servers, sid, num_normal_preds, est_load_factor, kvstore_bytes, 0, 0)
        st = BuildStats()
        h = C.c_void_p()
        _check(lib().wk_store_build(device, _ptr(t), t.shape[0], C.byref(o), C.byref(h), C.byref(st)), "wk_store_build")
        self = cls.__new__(cls)
        self.h = h
        self.device = device
        self.build_stats = {n: getattr(st, n) for n, _ in BuildStats._fields_}
        return self

    def info(self):
        ns, ne, k = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        _check(lib().wk_store_info(self.h, C.byref(ns), C.byref(ne), C.byref(k)))
        return ns.value, ne.value, k.value

    def segs(self):
        k = self.info()[2]
        sa = (SegMeta * k)()
        _check(lib().wk_store_segs(self.h, C.cast(sa, C.c_void_p), k))
        return list(sa)

    def download(self):
        """-> (vertices (num_slots, 2) uint64, edges uint32) copied back from the device"""
        ns, ne, _ = self.info()
        v = np.empty((ns, 2), dtype=np.uint64)
        e = np.empty(ne, dtype=np.uint32)
        _check(lib().wk_store_download(self.h, _ptr(v), ns, _ptr(e), ne))
        return v, e

    def get_edges(self, vid, pid, d, cap=1 << 20):
        out = np.empty(cap, dtype=np.uint32)
        n = C.c_uint64(0)
        _check(lib().wk_store_get_edges(self.h, vid, pid, d, _ptr(out), cap, C.byref(n)))
        return out[: n.value].copy()

    def close(self):
        if getattr(self, "h", None):
            lib().wk_store_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One query engine (wk_engine_t): device-resident binding table + pattern primitives."""

    def __init__(self, store, rbuf_bytes=256 << 20):
        self.store = store
        h = C.c_void_p()
        _check(lib().wk_engine_create(store.h, rbuf_bytes, C.byref(h)), "wk_engine_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().wk_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_profiling(self, level):
        _check(lib().wk_engine_set_profiling(self.h, int(level)))

    def set_option(self, option, value):
        _check(lib().wk_engine_set_option(self.h, int(option), int(value)), "wk_engine_set_option")

    def get_option(self, option):
        v = C.c_int64(0)
        _check(lib().wk_engine_get_option(self.h, int(option), C.byref(v)), "wk_engine_get_option")
        return v.value

    def set_resident(self, on):
        """light queries through the resident server kernel (default) or one launch per query"""
        self.set_option(WK_OPT_RESIDENT_LIGHT, 1 if on else 0)

    def light_trace(self):
        """profiling level 3: SM clocks at the light interpreter's phase boundaries (diagnostics; layout in wk_light.cuh)"""
        a = np.zeros(128, dtype=np.int64)
        _check(lib().wk_engine_light_trace(self.h, _ptr(a), 128))
        return a

    def last_query_device_us(self):
        us = C.c_float(0)
        _check(lib().wk_engine_last_query_device_us(self.h, C.byref(us)))
        return us.value

    def reset(self):
        _check(lib().wk_engine_reset(self.h))

    def sync(self):
        _check(lib().wk_engine_sync(self.h))

    def upload(self, table, ncols=None):
        t = np.ascontiguousarray(table, dtype=np.uint32)
        if ncols is None:
            ncols = t.shape[1]
        t = t.reshape(-1, ncols) if ncols else t.reshape(0, 0)
        _check(lib().wk_table_upload(self.h, _ptr(t) if t.size else None, t.shape[0], ncols))

    def info(self):
        n, c = C.c_uint64(0), C.c_int(0)
        _check(lib().wk_table_info(self.h, C.byref(n), C.byref(c)))
        return n.value, c.value

    def download(self):
        n, c = self.info()
        out = np.empty((max(n, 1), max(c, 1)), dtype=np.uint32)
        n2, c2 = C.c_uint64(0), C.c_int(0)
        _check(lib().wk_table_download(self.h, _ptr(out), out.size, C.byref(n2), C.byref(c2)))
        return out[: n2.value, : c2.value].copy() if c2.value else np.zeros((0, 0), np.uint32)

    def _rows(self, sync):
        return C.byref(C.c_uint64(0)) if sync else None

    def index_to_unknown(self, tpid, d, mt_tid=0, mt_factor=1, sync=True):
        n = C.c_uint64(0)
        _check(lib().wk_index_to_unknown(self.h, tpid, d, mt_tid, mt_factor, C.byref(n) if sync else None))
        return n.value

    def const_to_unknown(self, vid, pid, d, sync=True):
        n = C.c_uint64(0)
        _check(lib().wk_const_to_unknown(self.h, vid, pid, d, C.byref(n) if sync else None))
        return n.value

    def known_to_unknown(self, col_start, pid, d, sync=True):
        n = C.c_uint64(0)
        _check(lib().wk_known_to_unknown(self.h, col_start, pid, d, C.byref(n) if sync else None))
        return n.value

    def known_to_known(self, col_start, pid, d, col_end, sync=True):
        n = C.c_uint64(0)
        _check(lib().wk_known_to_known(self.h, col_start, pid, d, col_end, C.byref(n) if sync else None))
        return n.value

    def known_to_const(self, col_start, pid, d, end_const, sync=True):
        n = C.c_uint64(0)
        _check(lib().wk_known_to_const(self.h, col_start, pid, d, end_const, C.byref(n) if sync else None))
        return n.value

    def const_to_known(self, vid, pid, d, col_end, sync=True):
        n = C.c_uint64(0)
        _check(lib().wk_const_to_known(self.h, vid, pid, d, col_end, C.byref(n) if sync else None))
        return n.value

    def index_to_known(self, tpid, d, col_end, mt_tid=0, mt_factor=1, sync=True):
        n = C.c_uint64(0)
        _check(lib().wk_index_to_known(self.h, tpid, d, col_end, mt_tid, mt_factor, C.byref(n) if sync else None))
        return n.value

    def distinct(self, cols):
        a = np.array(cols, dtype=np.int32)
        n = C.c_uint64(0)
        _check(lib().wk_table_distinct(self.h, _ptr(a), len(a), C.byref(n)))
        return n.value

    def slice(self, offset, limit):
        n = C.c_uint64(0)
        _check(lib().wk_table_slice(self.h, offset, limit, C.byref(n)))
        return n.value

    def project(self, cols, sync=True):
        a = np.array(cols, dtype=np.int32)
        n = C.c_uint64(0)
        _check(lib().wk_project(self.h, _ptr(a), len(cols), C.byref(n) if sync else None))
        return n.value

    def query(self, patterns, nvars, required_vars, mt_tid=0, mt_factor=1, blind=False, out=None,
              distinct=False, offset=0, limit=-1):
        """wk_query_execute (wk_query_execute_ex when a modifier is given).  Returns (status, rows, cols, table-or-None)."""
        p = np.array(patterns, dtype=np.int32).reshape(-1, 4)
        rv = np.array(required_vars, dtype=np.int32)
        n, c = C.c_uint64(0), C.c_int(0)
        if distinct or offset > 0 or limit >= 0:
            o = QueryOpts(mt_tid, mt_factor, 1 if blind else 0, 1 if distinct else 0, offset, limit)
            if out is None:
                out = self._out_buf
            rc = lib().wk_query_execute_ex(self.h, _ptr(p), p.shape[0], nvars, _ptr(rv), len(rv), C.byref(o),
                                           None if blind else _ptr(out), 0 if blind else out.size, C.byref(n), C.byref(c))
            tbl = None
            if rc == 0 and not blind:
                tbl = out.reshape(-1)[: n.value * c.value].reshape(n.value, c.value) if c.value else np.zeros((0, 0), np.uint32)
            return rc, n.value, c.value, tbl
        if blind:
            rc = lib().wk_query_execute(self.h, _ptr(p), p.shape[0], nvars, _ptr(rv), len(rv), mt_tid, mt_factor, 1,
                                        None, 0, C.byref(n), C.byref(c))
            return rc, n.value, c.value, None
        if out is None:
            out = self._out_buf
        rc = lib().wk_query_execute(self.h, _ptr(p), p.shape[0], nvars, _ptr(rv), len(rv), mt_tid, mt_factor, 0,
                                    _ptr(out), out.size, C.byref(n), C.byref(c))
        tbl = None
        if rc == 0:
            tbl = out.reshape(-1)[: n.value * c.value].reshape(n.value, c.value) if c.value else np.zeros((0, 0), np.uint32)
        return rc, n.value, c.value, tbl

    def query_batch(self, plans):
        """wk_query_execute_batch.  plans: list of (patterns, nvars).  -> (rows[], status[])"""
        pats = np.concatenate([np.array(p, dtype=np.int32).reshape(-1, 4) for p, _ in plans])
        off = np.zeros(len(plans) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(p) for p, _ in plans])
        nv = np.array([n for _, n in plans], dtype=np.int32)
        return self.query_batch_raw(pats, off, nv)

    def query_batch_raw(self, pats, off, nv):
        rows = np.zeros(len(nv), dtype=np.uint64)
        st = np.zeros(len(nv), dtype=np.int32)
        _check(lib().wk_query_execute_batch(self.h, _ptr(pats), _ptr(off), _ptr(nv), len(nv), _ptr(rows), _ptr(st)),
               "wk_query_execute_batch")
        return rows, st

    _out_cache = None

    @property
    def _out_buf(self):
        if self._out_cache is None:
            self._out_cache = np.empty(64 << 20, dtype=np.uint32)
        return self._out_cache

    # ---- sharded execution ----
    def comm_init(self, nranks, rank, unique_id_bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id_bytes)
        _check(lib().wk_comm_init(self.h, nranks, rank, C.cast(buf, C.c_void_p)), "wk_comm_init")

    def p2p_export(self, nranks, rank):
        buf = (C.c_ubyte * 192)()
        _check(lib().wk_comm_p2p_export(self.h, nranks, rank, C.cast(buf, C.c_void_p)), "wk_comm_p2p_export")
        return bytes(buf)

    def p2p_import(self, all_handles_bytes):
        buf = (C.c_ubyte * len(all_handles_bytes)).from_buffer_copy(all_handles_bytes)
        _check(lib().wk_comm_p2p_import(self.h, C.cast(buf, C.c_void_p)), "wk_comm_p2p_import")

    def p2p_export_store(self):
        """handles of this rank's store arrays + its segment table (in-place light queries on a sharded store)"""
        n = C.c_uint64(0)
        lib().wk_comm_p2p_export_store(self.h, None, 0, C.byref(n))
        buf = (C.c_ubyte * n.value)()
        _check(lib().wk_comm_p2p_export_store(self.h, C.cast(buf, C.c_void_p), n.value, C.byref(n)), "wk_comm_p2p_export_store")
        return bytes(buf)

    def p2p_import_store(self, blobs):
        off = np.zeros(len(blobs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(b) for b in blobs])
        allb = b"".join(blobs)
        buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
        _check(lib().wk_comm_p2p_import_store(self.h, C.cast(buf, C.c_void_p), _ptr(off), len(blobs)), "wk_comm_p2p_import_store")

    def partition(self, col, nparts):
        out = np.zeros(nparts, dtype=np.uint64)
        _check(lib().wk_partition(self.h, col, nparts, _ptr(out)), "wk_partition")
        return out

    def exchange(self, col):
        n = C.c_uint64(0)
        _check(lib().wk_exchange(self.h, col, C.byref(n)), "wk_exchange")
        return n.value

    def exchange_p2p(self, col):
        n = C.c_uint64(0)
        _check(lib().wk_exchange_p2p(self.h, col, C.byref(n)), "wk_exchange_p2p")
        return n.value

    def comm_stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().wk_comm_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(exchanges=a.value, rows_sent=b.value, rows_recv=c.value)

    def query_sharded(self, patterns, nvars, required_vars, mt_tid=0, mt_factor=1, blind=False, out=None):
        p = np.array(patterns, dtype=np.int32).reshape(-1, 4)
        rv = np.array(required_vars, dtype=np.int32)
        n, c = C.c_uint64(0), C.c_int(0)
        if blind:
            rc = lib().wk_query_execute_sharded(self.h, _ptr(p), p.shape[0], nvars, _ptr(rv), len(rv), mt_tid, mt_factor, 1,
                                                None, 0, C.byref(n), C.byref(c))
            return rc, n.value, c.value, None
        if out is None:
            out = self._out_buf
        rc = lib().wk_query_execute_sharded(self.h, _ptr(p), p.shape[0], nvars, _ptr(rv), len(rv), mt_tid, mt_factor, 0,
                                            _ptr(out), out.size, C.byref(n), C.byref(c))
        tbl = None
        if rc == 0:
            tbl = out.reshape(-1)[: n.value * c.value].reshape(n.value, c.value) if c.value else np.zeros((0, 0), np.uint32)
        return rc, n.value, c.value, tbl

    def flush_l2(self):
        _check(lib().wk_engine_flush_l2(self.h))

    def step_stats(self):
        n = lib().wk_engine_num_steps(self.h)
        out = []
        for i in range(n):
            s = StepStats()
            _check(lib().wk_engine_step_stats(self.h, i, C.byref(s)))
            out.append(dict(kind=KIND_NAMES[s.kind], in_cols=s.in_cols, in_rows=s.in_rows, out_rows=s.out_rows,
                            buckets_visited=s.buckets_visited, edges_touched=s.edges_touched,
                            algo_bytes=s.algo_bytes, device_us=s.device_us, launches=s.launches))
        return out

    def launch_count(self):
        return lib().wk_engine_launch_count(self.h)


def pinned_array(nwords):
    """uint32 numpy array backed by page-locked host memory (cudaHostAlloc)."""
    p = C.c_void_p()
    _check(lib().wk_host_alloc(nwords * 4, C.byref(p)), "wk_host_alloc")
    arr = np.frombuffer((C.c_uint32 * nwords).from_address(p.value), dtype=np.uint32)
    return arr, p


def local_group(engines):
    """wk_comm_local_group: engines of this process become ranks 0..n-1 of one peer-memory group"""
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    _check(lib().wk_comm_local_group(C.cast(arr, C.c_void_p), len(engines)), "wk_comm_local_group")


def comm_unique_id():
    buf = (C.c_ubyte * 128)()
    _check(lib().wk_comm_unique_id(C.cast(buf, C.c_void_p)), "wk_comm_unique_id")
    return bytes(buf)


def plan_exchanges(patterns, nvars):
    """Host-only: per step -1 (no exchange), -2 (replicate to all shards) or the column to re-shard by."""
    p = np.array(patterns, dtype=np.int32).reshape(-1, 4)
    out = np.zeros(p.shape[0], dtype=np.int32)
    _check(lib().wk_plan_exchanges(_ptr(p), p.shape[0], nvars, _ptr(out)), "wk_plan_exchanges")
    return [int(x) for x in out]
