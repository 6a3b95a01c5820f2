"""ctypes binding of the synthetic data generators in libwukong_host.so."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_HOST = os.path.join(_HERE, "libwukong_host.so")
_lib = None

LUBM_NUM_INDEX_IDS = 32          # str_index lines (incl. __PREDICATE__)
LUBM_NUM_NORMAL_PREDS = 31       # what the reference loader derives: lines - 1 (base_loader.hpp:409-424)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_HOST):
            raise RuntimeError("libwukong_host.so is missing: run `python -m wukong_b200.build`")
        from . import capi
        capi.lib()  # libwukong_host.so depends on libwukong_b200.so
        L = C.CDLL(LIB_HOST, mode=C.RTLD_GLOBAL)
        L.wkgen_lubm.restype = C.c_uint64
        L.wkgen_lubm.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
        L.wkgen_lubm_write_dir.restype = C.c_uint64
        L.wkgen_lubm_write_dir.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64]
        L.wkgen_lubm_univ_id.restype = C.c_uint32
        L.wkgen_lubm_univ_id.argtypes = [C.c_uint32]
        L.wkgen_lubm_dept_id.restype = C.c_uint32
        L.wkgen_lubm_dept_id.argtypes = [C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def lubm(num_univs, seed=1, u_begin=0, u_end=None):
    """ID triples (n, 3) uint32 of universities [u_begin, u_end) of a `num_univs` dataset."""
    if u_end is None:
        u_end = num_univs
    L = lib()
    n = L.wkgen_lubm(u_begin, u_end, num_univs, seed, None, 0)
    out = np.empty((n, 3), dtype=np.uint32)
    m = L.wkgen_lubm(u_begin, u_end, num_univs, seed, out.ctypes.data_as(C.c_void_p), n)
    assert m == n
    return out


def lubm_write_dir(path, num_univs, seed=1):
    os.makedirs(path, exist_ok=True)
    n = lib().wkgen_lubm_write_dir(path.encode(), num_univs, seed)
    if n == 0:
        raise RuntimeError("failed to write dataset to " + path)
    return n


def lubm_shard(num_univs, nranks, rank, seed=1, chunk=128):
    """Triples of a `num_univs` dataset that shard `rank` of `nranks` stores (subject or object owned by it:
    vid % nranks == rank, reference base_loader.hpp:169-181), generated chunk-wise to bound host memory."""
    parts = []
    for u0 in range(0, num_univs, chunk):
        t = lubm(num_univs, seed=seed, u_begin=u0, u_end=min(num_univs, u0 + chunk))
        keep = ((t[:, 0] % nranks) == rank) | ((t[:, 2] % nranks) == rank)
        parts.append(t[keep])
    return np.concatenate(parts) if parts else np.zeros((0, 3), dtype=np.uint32)


RMAT_PRED, RMAT_TYPE, RMAT_NUM_NORMAL_PREDS = 2, 3, 3   # str_index: __PREDICATE__, rdf:type, <edge>, <Vertex>


def rmat(scale, nedges, seed=42, a=0.57, b=0.19, c=0.19, typed=True, scramble=True):
    """R-MAT power-law graph as ID triples: (s, 2, o) edges, plus (v, 1, 3) for every vertex that occurs.  scramble: relabel the
    vertices through a bijection (as Graph500 does), so that the degree is not readable from the low bits of the id."""
    L = lib()
    L.wkgen_rmat_edges.restype = C.c_uint64
    L.wkgen_rmat_edges.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
    out = np.empty((nedges, 3), dtype=np.uint32)
    L.wkgen_rmat_edges(scale, nedges, seed, a, b, c, 1 if scramble else 0, out.ctypes.data_as(C.c_void_p))
    if not typed:
        return out
    verts = np.unique(np.concatenate([out[:, 0], out[:, 2]]))
    tt = np.empty((verts.shape[0], 3), dtype=np.uint32)
    tt[:, 0] = verts
    tt[:, 1] = 1
    tt[:, 2] = RMAT_TYPE
    return np.concatenate([out, tt])
