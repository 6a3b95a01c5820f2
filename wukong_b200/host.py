"""ctypes binding of the host-side C++ layer (libwukong_host.so): store builder + Wukong-surface mirror."""
import ctypes as C

import numpy as np

from . import capi, datagen


def lib():
    L = datagen.lib()
    if getattr(L, "_wkh_ready", False):
        return L
    u64, u32, vp, ci = C.c_uint64, C.c_uint32, C.c_void_p, C.c_int
    L.wkh_store_build.restype = vp
    L.wkh_store_build.argtypes = [vp, u64, ci, ci, ci, u64, ci, ci]
    L.wkh_store_free.argtypes = [vp]
    L.wkh_store_ok.argtypes = [vp]
    L.wkh_store_error.restype = C.c_char_p
    L.wkh_store_error.argtypes = [vp]
    for f in ("wkh_store_vertices", "wkh_store_edges"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    for f in ("wkh_store_num_slots", "wkh_store_num_edges", "wkh_store_num_keys", "wkh_store_num_buckets",
              "wkh_store_used_ext"):
        getattr(L, f).restype = u64
        getattr(L, f).argtypes = [vp]
    L.wkh_store_num_segs.argtypes = [vp]
    L.wkh_store_segs.argtypes = [vp, vp]
    L.wkh_store_get_edges.restype = u64
    L.wkh_store_get_edges.argtypes = [vp, u32, u32, ci, C.POINTER(vp)]
    L.wkh_store_upload.argtypes = [vp, ci, C.POINTER(vp)]
    L.wkh_time_query.argtypes = [vp, vp, ci, ci, vp, ci, ci, ci, ci, vp, u64, ci, ci, vp, vp, C.POINTER(u64),
                                 C.POINTER(ci)]
    L.wkh_time_query_sharded.argtypes = [vp, vp, ci, ci, vp, ci, ci, vp, u64, ci, vp, ci, ci, ci, C.c_int64, C.POINTER(C.c_double),
                                         C.POINTER(C.c_float), C.POINTER(u64), C.POINTER(ci), C.POINTER(ci), C.POINTER(u64)]
    L.wkh_env_create.restype = vp
    L.wkh_env_create.argtypes = [C.c_char_p, ci]
    L.wkh_env_destroy.argtypes = [vp]
    L.wkh_env_error.restype = C.c_char_p
    L.wkh_env_error.argtypes = [vp]
    L.wkh_env_num_triples.restype = u64
    L.wkh_env_num_triples.argtypes = [vp]
    L.wkh_env_num_normal_preds.argtypes = [vp]
    L.wkh_env_num_keys.restype = u64
    L.wkh_env_num_keys.argtypes = [vp]
    L.wkh_env_config_int.argtypes = [vp, C.c_char_p]
    L.wkh_parse_plan.argtypes = [vp, C.c_char_p, C.c_char_p, vp, ci, C.POINTER(ci), C.POINTER(ci), vp, ci, C.POINTER(ci)]
    L.wkh_run_single_query.argtypes = [vp, C.c_char_p, C.c_char_p, ci, ci, ci, vp, u64, C.POINTER(u64), C.POINTER(ci),
                                       C.POINTER(C.c_double), C.POINTER(u64)]
    L._wkh_ready = True
    return L


class HostStore:
    """Cluster-hash graph store built on the host by the product builder (csrc/store/host_builder.cpp)."""

    def __init__(self, triples, num_servers=1, sid=0, num_normal_preds=datagen.LUBM_NUM_NORMAL_PREDS,
                 kvstore_bytes=0, est_load_factor=55, gpu_ext_extents=True):
        t = np.ascontiguousarray(triples, dtype=np.uint32).reshape(-1, 3)
        self.h = lib().wkh_store_build(t.ctypes.data_as(C.c_void_p), t.shape[0], num_servers, sid, num_normal_preds,
                                       kvstore_bytes, est_load_factor, 1 if gpu_ext_extents else 0)
        if not lib().wkh_store_ok(self.h):
            msg = lib().wkh_store_error(self.h).decode()
            lib().wkh_store_free(self.h)
            self.h = None
            raise RuntimeError("store build failed: " + msg)

    def __del__(self):
        try:
            if self.h:
                lib().wkh_store_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def num_slots(self):
        return lib().wkh_store_num_slots(self.h)

    @property
    def num_edges(self):
        return lib().wkh_store_num_edges(self.h)

    @property
    def num_keys(self):
        return lib().wkh_store_num_keys(self.h)

    def vertices(self):
        n = self.num_slots
        return np.frombuffer((C.c_uint64 * (2 * n)).from_address(lib().wkh_store_vertices(self.h)), dtype=np.uint64).reshape(n, 2)

    def edges(self):
        n = self.num_edges
        return np.frombuffer((C.c_uint32 * n).from_address(lib().wkh_store_edges(self.h)), dtype=np.uint32)

    def segs(self):
        n = lib().wkh_store_num_segs(self.h)
        arr = (capi.SegMeta * n)()
        lib().wkh_store_segs(self.h, C.cast(arr, C.c_void_p))
        return list(arr)

    def get_edges(self, vid, pid, d):
        p = C.c_void_p()
        n = lib().wkh_store_get_edges(self.h, vid, pid, d, C.byref(p))
        if n == 0 or not p.value:
            return np.zeros(0, dtype=np.uint32)
        return np.frombuffer((C.c_uint32 * n).from_address(p.value), dtype=np.uint32).copy()

    def upload(self, device=0):
        """-> capi.Store living on `device` (wk_store_create through the C ABI)."""
        h = C.c_void_p()
        capi._check(lib().wkh_store_upload(self.h, device, C.byref(h)), "wkh_store_upload")
        st = capi.Store.__new__(capi.Store)
        st.h = h
        st.device = device
        return st


def time_query(engine, patterns, nvars, required_vars, reps, blind=True, table=None, flush=True, mt_tid=0,
               mt_factor=1, device_times=False):
    """Timed loop in native code around wk_query_execute.  Returns (wall_us[], dev_us[] or None, rows, cols)."""
    p = np.array(patterns, dtype=np.int32).reshape(-1, 4)
    rv = np.array(required_vars, dtype=np.int32)
    wall = np.zeros(reps, dtype=np.float64)
    dev = np.zeros(reps, dtype=np.float32) if device_times else None
    rows, cols = C.c_uint64(0), C.c_int(0)
    rc = lib().wkh_time_query(engine.h, p.ctypes.data_as(C.c_void_p), p.shape[0], nvars,
                              rv.ctypes.data_as(C.c_void_p), len(rv), mt_tid, mt_factor, 1 if blind else 0,
                              table.ctypes.data_as(C.c_void_p) if table is not None else None,
                              table.size if table is not None else 0, reps, 1 if flush else 0,
                              wall.ctypes.data_as(C.c_void_p),
                              dev.ctypes.data_as(C.c_void_p) if dev is not None else None,
                              C.byref(rows), C.byref(cols))
    capi._check(rc, "wkh_time_query")
    return wall, dev, rows.value, cols.value


class ShardedTimer:
    """Timed collective queries of a sharded group in native code (wkh_time_query_sharded): L2 flush, stream sync, a spin barrier
    over `slots` (a shared int64 array, rank r's generation counter at slots[8 * r]), then the clock around
    wk_query_execute_sharded.  Every rank calls time() for the same query at the same time."""

    def __init__(self, engine, slots, rank, world):
        self.engine, self.slots, self.rank, self.world, self.gen = engine, slots, rank, world, 0
        if slots is not None:
            self.gen = int(slots[8 * rank])

    def time(self, patterns, nvars, required_vars, blind=True, table=None, flush=True):
        """-> (wall_us, dev_us, rows, cols, resident, server_ns)"""
        p = np.array(patterns, dtype=np.int32).reshape(-1, 4)
        rv = np.array(required_vars, dtype=np.int32)
        wall, dev = C.c_double(0), C.c_float(0)
        rows, cols, res, ns = C.c_uint64(0), C.c_int(0), C.c_int(0), C.c_uint64(0)
        self.gen += 1
        rc = lib().wkh_time_query_sharded(self.engine.h, p.ctypes.data_as(C.c_void_p), p.shape[0], nvars, rv.ctypes.data_as(C.c_void_p),
                                          len(rv), 1 if blind else 0, table.ctypes.data_as(C.c_void_p) if table is not None else None,
                                          table.size if table is not None else 0, 1 if flush else 0,
                                          self.slots.ctypes.data_as(C.c_void_p) if self.slots is not None else None, 8, self.rank,
                                          self.world, self.gen, C.byref(wall), C.byref(dev), C.byref(rows), C.byref(cols),
                                          C.byref(res), C.byref(ns))
        capi._check(rc, "wkh_time_query_sharded")
        return wall.value, dev.value, rows.value, cols.value, bool(res.value), ns.value


CONFIG_ITEMS = ["num_servers", "num_threads", "num_proxies", "num_engines", "data_port_base", "ctrl_port_base", "rdma_buf_size_mb",
                "rdma_rbf_size_mb", "use_rdma", "rdma_threshold", "mt_threshold", "enable_caching", "enable_workstealing",
                "stealing_pattern", "silent", "enable_planner", "generate_statistics", "enable_vattr", "memstore_size_gb",
                "est_load_factor", "num_gpus", "gpu_kvcache_size_gb", "gpu_rbuf_size_mb", "gpu_rdma_buf_size_mb",
                "gpu_key_blk_size_mb", "gpu_value_blk_size_mb", "gpu_enable_pipeline"]


def set_plan_tree(tree_ints, fmt_text):
    """Planner::set_plan of the C++ host mirror on a pattern-group tree ([npat, (s,p,d,o)*, nunions, tree*, noptional, tree*])
    -> planned tree as a list of ints, or None when the plan is refused"""
    L = lib()
    L.wkh_set_plan_tree.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
    a = np.array(tree_ints, dtype=np.int32)
    out = np.zeros(4096, dtype=np.int32)
    n = L.wkh_set_plan_tree(a.ctypes.data_as(C.c_void_p), a.size, fmt_text.encode(), out.ctypes.data_as(C.c_void_p), out.size)
    return None if n < 0 else out[:n].tolist()


def load_config(fname, nsrvs, reload="", gpu_build=False):
    """Global::load_config(fname, nsrvs) [+ reload_config(reload)] of the C++ host mirror, starting from the reference's
    defaults -> dict of the items (None when the file cannot be read or a value is refused)"""
    L = lib()
    L.wkh_config_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    out = np.zeros(len(CONFIG_ITEMS), dtype=np.int32)
    folder = C.create_string_buffer(4096)
    n = L.wkh_config_load(fname.encode(), nsrvs, 1 if gpu_build else 0, reload.encode(), out.ctypes.data_as(C.c_void_p), len(out), folder, 4096)
    if n < 0:
        return None
    d = {k: int(v) for k, v in zip(CONFIG_ITEMS, out)}
    d["input_folder"] = folder.value.decode()
    return d


class Env:
    """One Wukong-surface server: Global config + StringServer + DGraph + GPUEngine + Proxy (C++, csrc/host/)."""

    def __init__(self, config_text, device=0):
        self.h = lib().wkh_env_create(config_text.encode(), device)
        err = lib().wkh_env_error(self.h).decode()
        if err:
            lib().wkh_env_destroy(self.h)
            self.h = None
            raise RuntimeError(err)

    def close(self):
        if self.h:
            lib().wkh_env_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_triples(self):
        return lib().wkh_env_num_triples(self.h)

    @property
    def num_normal_preds(self):
        return lib().wkh_env_num_normal_preds(self.h)

    def config_int(self, key):
        return lib().wkh_env_config_int(self.h, key.encode())

    def parse_plan(self, query_text, fmt_text):
        """-> (status, patterns [(s,p,d,o)], nvars, required_vars) from the C++ Parser + Planner::set_plan"""
        pats = np.zeros((64, 4), dtype=np.int32)
        req = np.zeros(64, dtype=np.int32)
        n, nv, nr = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = lib().wkh_parse_plan(self.h, query_text.encode(), fmt_text.encode(), pats.ctypes.data_as(C.c_void_p), 64,
                                  C.byref(n), C.byref(nv), req.ctypes.data_as(C.c_void_p), 64, C.byref(nr))
        return rc, [tuple(int(x) for x in r) for r in pats[: n.value]], nv.value, [int(x) for x in req[: nr.value]]

    def run_single_query(self, query_text, fmt_text, mt_factor=1, cnt=1, per_pattern=False, cap_words=1 << 24):
        """Proxy::run_single_query -> (status, rows, cols, table or None, latency_us)"""
        out = np.empty(cap_words, dtype=np.uint32)
        rows, cols, lat, nw = C.c_uint64(0), C.c_int(0), C.c_double(0), C.c_uint64(0)
        rc = lib().wkh_run_single_query(self.h, query_text.encode(), fmt_text.encode(), mt_factor, cnt, 1 if per_pattern else 0,
                                        out.ctypes.data_as(C.c_void_p), cap_words, C.byref(rows), C.byref(cols), C.byref(lat),
                                        C.byref(nw))
        tbl = None
        if rc == 0 and cols.value and nw.value and nw.value <= cap_words:
            tbl = out[: nw.value].reshape(-1, cols.value).copy()
        return rc, rows.value, cols.value, tbl, lat.value
