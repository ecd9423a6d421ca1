"""ctypes access to the CHECKER (oracle/liboracle.so and, when built, oracle/_ref/libref_cl.so).

Test infrastructure only: imported from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_cl.so")
RESULT_DTYPE = np.dtype([("score", "<i4"), ("end_i", "<i4"), ("end_j", "<i4")])


class _Res(ctypes.Structure):
    _fields_ = [("score", ctypes.c_int32), ("end_i", ctypes.c_int32), ("end_j", ctypes.c_int32)]


_o = None
_r = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def oracle():
    global _o
    if _o is None:
        if not os.path.exists(ORACLE_SO):
            build()
        o = ctypes.CDLL(ORACLE_SO)
        o.sw_linear.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(_Res)]
        o.sw_last_row_max.restype = ctypes.c_int32
        o.sw_last_row_max.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64]
        o.ref_compat_align.restype = ctypes.c_int32
        o.ref_compat_align.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32]
        for f in (o.sw_linear_batch, o.sw_simd_batch):
            f.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
        o.sw_simd_isa.restype = ctypes.c_char_p
        o.sw_simd_force_isa.argtypes = [ctypes.c_int]
        _o = o
    return _o


def ref_cl():
    """The reference's own kernels compiled from /root/reference (None when not built)."""
    global _r
    if _r is None and os.path.exists(REF_SO):
        r = ctypes.CDLL(REF_SO)
        r.refcl_run_detailed.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32,
                                         ctypes.c_uint32, ctypes.POINTER(ctypes.c_int32)]
        r.refcl_gpu_align.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                      ctypes.c_uint32, ctypes.POINTER(ctypes.c_int32)]
        r.refcl_run_align.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.POINTER(ctypes.c_int32)]
        _r = r
    return _r


def _buf(x):
    if isinstance(x, str):
        x = x.encode()
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8)
    return np.frombuffer(bytes(x), dtype=np.uint8)


def sw_linear(a, b):
    a, b = _buf(a), _buf(b)
    res = _Res()
    rc = oracle().sw_linear(a.ctypes.data, a.size, b.ctypes.data, b.size, ctypes.byref(res))
    assert rc == 0
    return int(res.score), int(res.end_i), int(res.end_j)


def traceback(a, b, end_i, end_j):
    """(start_i, start_j, [(length, op), ...]) behind an end cell; op is one of '=', 'X', 'I', 'D'."""
    a, b = _buf(a), _buf(b)
    cap = a.size + b.size + 2
    ops = np.zeros(cap, dtype=np.uint32)
    si, sj = ctypes.c_int32(), ctypes.c_int32()
    o = oracle()
    o.sw_traceback.restype = ctypes.c_int
    o.sw_traceback.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32,
                               ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p, ctypes.c_uint32]
    n = o.sw_traceback(a.ctypes.data, a.size, b.ctypes.data, b.size, int(end_i), int(end_j), ctypes.byref(si), ctypes.byref(sj),
                       ops.ctypes.data, cap)
    assert n >= 0
    return int(si.value), int(sj.value), [(int(v >> 4), "MIDNSHP=X"[int(v & 15)]) for v in ops[:n]]


def last_row_max(a, b):
    a, b = _buf(a), _buf(b)
    return int(oracle().sw_last_row_max(a.ctypes.data, a.size, b.ctypes.data, b.size))


def ref_compat_align(a, b, dev_max_wg=1024):
    a, b = _buf(a), _buf(b)
    return int(oracle().ref_compat_align(a.ctypes.data, a.size, b.ctypes.data, b.size, dev_max_wg))


def batch(q, qo, r, ro, threads=1, simd=False):
    q = np.ascontiguousarray(q, dtype=np.uint8); r = np.ascontiguousarray(r, dtype=np.uint8)
    qo = np.ascontiguousarray(qo, dtype=np.uint64); ro = np.ascontiguousarray(ro, dtype=np.uint64)
    n = qo.size - 1
    out = np.zeros(n, dtype=RESULT_DTYPE)
    fn = oracle().sw_simd_batch if simd else oracle().sw_linear_batch
    rc = fn(q.ctypes.data, qo.ctypes.data, r.ctypes.data, ro.ctypes.data, n, out.ctypes.data, int(threads))
    assert rc == 0
    return out


def simd_isa():
    return oracle().sw_simd_isa().decode()


def ref_detailed(a, b, local_size=256):
    a, b = _buf(a), _buf(b)
    out = ctypes.c_int32()
    rc = ref_cl().refcl_run_detailed(a.ctypes.data, a.size, b.ctypes.data, b.size, local_size, ctypes.byref(out))
    assert rc == 0, "refcl_run_detailed refused the input (len2 <= 256 <= local_size required)"
    return int(out.value)


def ref_gpu_align(a, b, dev_max_wg=256):
    a, b = _buf(a), _buf(b)
    out = ctypes.c_int32()
    rc = ref_cl().refcl_gpu_align(a.ctypes.data, a.size, b.ctypes.data, b.size, dev_max_wg, ctypes.byref(out))
    assert rc == 0
    return int(out.value)
