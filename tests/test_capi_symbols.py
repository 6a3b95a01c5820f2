"""CPU: the C-ABI library loads without a GPU and exports every symbol include/wukong_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O
from wukong_b200 import capi


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "wukong_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wk_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(capi.DECLARED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    for name in _header_symbols():
        assert hasattr(L, name), name
    assert L.wk_version() >= 100
    assert b"overflow" in L.wk_strerror(capi.WK_ERR_RBUF_OVERFLOW)


def test_device_arithmetic_matches_oracle():
    # the hash / key packing / multiply-shift modulo compiled into the CUDA library (host instances
    # of the same __host__ __device__ functions the kernels use) against the oracle's
    L, Lo = capi.lib(), O.lib()
    rng = np.random.default_rng(3)
    for k in [0, 1, (1 << 64) - 1] + [int(x) for x in rng.integers(0, 1 << 63, 2000)]:
        assert L.wk_selftest_hash(k) == Lo.wko_hash_u64(k)
    for vid, pid, d in [(0, 1, 0), (131072, 5, 1), ((1 << 32) - 1, (1 << 17) - 1, 1)]:
        assert L.wk_selftest_make_key(vid, pid, d) == Lo.wko_make_key(vid, pid, d)
    ds = [1, 2, 3, 5, 7, 8, 255, 256, 98317, 196613, 12582917, 201326611, 1610612741, (1 << 31) - 1, (1 << 32) - 1]
    ds += [int(x) for x in rng.integers(1, 1 << 32, 200)]
    for d in ds:
        ns = [0, 1, d - 1, d, d + 1, (1 << 64) - 1, (1 << 63), (1 << 63) - 1] + [int(x) for x in rng.integers(0, 1 << 63, 300)]
        ns += [(int(x) << 1) | 1 for x in rng.integers(1 << 62, 1 << 63, 100)]
        for n in ns:
            assert L.wk_selftest_fastmod(n, d) == n % d, (n, d)
    # the exchange's owner function (row[col] % nranks by multiply-shift with magic = ceil(2^32 / n)), every n a box can hold
    xs = [0, 1, 2, 15, 16, 17, (1 << 17), (1 << 31) - 1, 1 << 31, (1 << 32) - 2, (1 << 32) - 1] + [int(x) for x in rng.integers(0, 1 << 32, 3000)]
    for n in range(1, 17):
        for x in xs:
            assert L.wk_selftest_owner(x, n) == x % n, (x, n)


def test_compute_fails_loudly_without_gpu():
    if capi.device_count() > 0:
        return
    v = np.zeros((8, 2), dtype=np.uint64)
    e = np.zeros(1, dtype=np.uint32)
    seg = capi.SegMeta()
    seg.num_buckets = 1
    try:
        capi.Store(v, e, [seg])
    except capi.WukongError as ex:
        assert ex.code in (capi.WK_ERR_NO_DEVICE, capi.WK_ERR_CUDA)
    else:
        raise AssertionError("store creation must fail without a device (no CPU fallback)")


def test_header_is_plain_c():
    """the boundary is a C ABI: the header must compile as C99 (and as C++) with no torch / C++ types in it"""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "wukong_b200.h")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not cc:
        pytest.skip("no C compiler")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call([cc, "-std=c++11", "-fsyntax-only", "-x", "c++", hdr])
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)      # declarations only, comments stripped
    assert "torch" not in code and "std::" not in code and "at::" not in code and "pytest" not in code


def test_argument_checks_need_no_gpu():
    """every entry point rejects null handles / malformed arguments before touching CUDA; without a device the store
    constructors say so instead of falling back to anything"""
    L = capi.lib()
    null = C.c_void_p(None)
    n = C.c_uint64(0)
    assert L.wk_engine_sync(null) != 0
    assert L.wk_table_distinct(null, null, 0, C.byref(n)) == capi.WK_ERR_BAD_ARG
    assert L.wk_table_slice(null, 0, -1, C.byref(n)) == capi.WK_ERR_BAD_ARG
    assert L.wk_const_to_known(null, 1, 2, 0, 0, C.byref(n)) == capi.WK_ERR_BAD_ARG
    assert L.wk_index_to_known(null, 2, 0, 0, 0, 1, C.byref(n)) == capi.WK_ERR_BAD_ARG
    assert L.wk_store_info(null, None, None, None) == capi.WK_ERR_BAD_ARG
    assert L.wk_query_execute_ex(null, null, 0, 0, null, 0, null, null, 0, C.byref(n), None) == capi.WK_ERR_BAD_ARG
    h = C.c_void_p()
    assert L.wk_store_build(0, None, 5, None, C.byref(h), None) == capi.WK_ERR_BAD_ARG          # no options
    o = capi.BuildOpts(2, 2, 31, 55, 0, 0, 0)                                                   # sid out of range
    assert L.wk_store_build(0, None, 0, C.byref(o), C.byref(h), None) == capi.WK_ERR_BAD_ARG
    if capi.device_count() == 0:
        o = capi.BuildOpts(1, 0, 31, 55, 0, 0, 0)
        t = np.zeros((1, 3), dtype=np.uint32)
        assert L.wk_store_build(0, t.ctypes.data_as(C.c_void_p), 1, C.byref(o), C.byref(h), None) == capi.WK_ERR_NO_DEVICE
        v = np.zeros((8, 2), dtype=np.uint64)
        assert L.wk_store_create(0, v.ctypes.data_as(C.c_void_p), 8, None, 0, None, 0, C.byref(h)) != 0
    assert b"store build" in L.wk_strerror(capi.WK_ERR_STORE_FULL)


def test_c_example_builds_against_the_abi(tmp_path):
    """examples/query_c_abi.c: a plain C99 program that uses only include/wukong_b200.h compiles and links against the library
    (it needs a GPU to do anything useful: without one it reports `no CUDA device` and exits 1)"""
    import shutil
    import subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not cc:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "query_c_abi")
    libdir = os.path.join(ROOT, "wukong_b200")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "query_c_abi.c"), "-L", libdir, "-l:libwukong_b200.so",
                           "-Wl,-rpath," + libdir, "-o", exe])
    if capi.device_count() == 0:
        f = tmp_path / "t.bin"
        np.zeros((4, 3), dtype=np.uint32).tofile(str(f))
        r = subprocess.run([exe, str(f), "31"], capture_output=True)
        assert r.returncode == 1 and b"no CUDA device" in r.stderr
