"""ctypes binding of include/swb200.h (the C ABI a Rust ``extern "C"`` block would bind)."""
import ctypes
import os

import numpy as np

LIB_PATH = os.environ.get("SWB_LIB_OVERRIDE") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libswb200.so")   # override: timing experiments only


class SwbResult(ctypes.Structure):
    _fields_ = [("score", ctypes.c_int32), ("end_i", ctypes.c_int32), ("end_j", ctypes.c_int32)]


RESULT_DTYPE = np.dtype([("score", "<i4"), ("end_i", "<i4"), ("end_j", "<i4")])
ALIGNMENT_DTYPE = np.dtype([("start_i", "<i4"), ("start_j", "<i4"), ("cigar_len", "<u4"), ("status", "<u4"), ("cigar_off", "<u8")])   # swb_alignment

_u8p = ctypes.c_void_p
_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32
_vp = ctypes.c_void_p
_int = ctypes.c_int

# every symbol include/swb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "swb_device_count": (_int, []),
    "swb_device_info": (_int, [_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_int)]),
    "swb_memory_info": (_int, [_int, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "swb_create": (_int, [ctypes.POINTER(_vp), _int, _vp]),
    "swb_destroy": (None, [_vp]),
    "swb_score_pair": (_int, [_vp, _u8p, _u64, _u8p, _u64, ctypes.POINTER(SwbResult)]),
    "swb_score_batch": (_int, [_vp, _u8p, _vp, _u8p, _vp, _u64, _vp]),
    "swb_score_batch_device": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, _u64, _u32, _u32, _vp]),
    "swb_sync": (_int, [_vp]),
    "swb_set_reference": (_int, [_vp, _u8p, _u64]),
    "swb_score_batch_vs_reference": (_int, [_vp, _u8p, _vp, _u64, _vp, _vp, _vp]),
    "swb_score_batch_ranges": (_int, [_vp, _u8p, _vp, _u64, _u8p, _u64, _vp, _vp, _vp]),
    "swb_last_ranges_info": (_int, [_vp, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "swb_bind_thread": (_int, [_vp]),
    "swb_debug_guard_check": (_int, [_vp, ctypes.c_char_p, _u64, ctypes.POINTER(_u64)]),
    "swb_create_multi": (_int, [ctypes.POINTER(_vp), _vp, _int, _vp]),
    "swb_destroy_multi": (None, [_vp]),
    "swb_multi_device_count": (_int, [_vp]),
    "swb_multi_ctx": (_vp, [_vp, _int]),
    "swb_multi_score_batch": (_int, [_vp, _u8p, _vp, _u8p, _vp, _u64, _vp]),
    "swb_multi_set_reference": (_int, [_vp, _u8p, _u64]),
    "swb_multi_score_batch_vs_reference": (_int, [_vp, _u8p, _vp, _u64, _vp, _vp, _vp]),
    "swb_synth_device_ref": (_int, [_vp, _vp, _u64, _u64, _u64, _u32, _u32, _int, _vp, _vp, _vp, _vp, _vp]),
    "swb_ref_compat_align": (_int, [_vp, _u8p, _u64, _u8p, _u64, _u32, ctypes.POINTER(ctypes.c_int32)]),
    "swb_last_row_max": (_int, [_vp, _u8p, _u64, _u8p, _u64, ctypes.POINTER(ctypes.c_int32)]),
    "swb_pack2bit": (_int, [_vp, _u8p, _u64, _vp, _vp]),
    "swb_pack2bit_device": (_int, [_vp, _vp, _u64, _vp, _vp]),
    "swb_synth_device": (_int, [_vp, _u64, _u64, _u32, _u32, _int, _vp, _vp, _vp, _vp]),
    "swb_last_timings": (_int, [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_int)]),
    "swb_last_routing": (_int, [_vp, ctypes.POINTER(_u64)]),
    "swb_last_routing_ex": (_int, [_vp, ctypes.POINTER(_u64)]),
    "swb_set_mid_path": (_int, [_vp, _int]),
    "swb_set_short_variant": (_int, [_vp, _int]),
    "swb_traceback_batch": (_int, [_vp, _u8p, _vp, _u8p, _vp, _u64, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64)]),
    "swb_set_chunking": (_int, [_vp, _u64, _u64]),
    "swb_set_chunk_ramp": (_int, [_vp, _int]),
    "swb_fastq_bgzf_prefetch": (_int, [_vp, _u8p, _u64, _vp, _u64]),
    "swb_fastq_bgzf_cancel": (_int, [_vp, _u8p]),
    "swb_fastq_bgzf_score": (_int, [_vp, _u8p, _u64, _vp, _u64, _u8p, _u64, _int, _u64, _u64, _u32,
                                    ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(_u64),
                                    _u8p, _u64, ctypes.POINTER(_u64), ctypes.POINTER(_int)]),
    "swb_numa_prefer_device": (_int, [_int]),
    "swb_numa_reset": (None, []),
    "swb_stream": (_vp, [_vp]),
    "swb_last_error": (ctypes.c_char_p, []),
    "swb_version": (ctypes.c_char_p, []),
    "swb_malloc_device": (_int, [_vp, _u64, ctypes.POINTER(_vp)]),
    "swb_free_device": (_int, [_vp, _vp]),
    "swb_malloc_pinned": (_int, [_u64, ctypes.POINTER(_vp)]),
    "swb_free_pinned": (_int, [_vp]),
    "swb_memcpy_d2h": (_int, [_vp, _vp, _vp, _u64]),
    "swb_memcpy_h2d": (_int, [_vp, _vp, _vp, _u64]),
}

_lib = None


def bind(path):
    """ctypes handle of one build of the library with every declared symbol typed."""
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA library first (make, or __graft_entry__.build()). "
            "There is no CPU fallback in this package.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library():
    """Load libswb200.so.  Raises (never falls back) when the CUDA library is missing."""
    global _lib
    if _lib is None:
        _lib = bind(LIB_PATH)
    return _lib
