"""Python view of the host mirror (include/rustseq_host.h, csrc/rustseq_host.cpp): the reference's Rust host
functions for the alignment path under their own names.  Everything here calls the compiled C++ host; the only
Python logic is argument marshalling.

    reference (smith_waterman/src)                       here
    aligner.rs:9-15    get_chunk_size_reads()            get_chunk_size_reads()
    aligner.rs:107-178 process_fastq_file_in_chunks()    process_fastq_file_in_chunks(path, n, processor)
    aligner.rs:535-544 count_bases_in_fastq()            count_bases_in_fastq(path)
    aligner.rs:410-532 gpu_align()                       gpu_align(seq1, seq2, device)
    aligner.rs:365-373 gpu_align_chunk_self()            gpu_align_chunk_self(chunk, device)
    aligner.rs:376-407 gpu_align_pair()                  gpu_align_pair(file1, file2, device)
    aligner.rs:183-362 process_full_wgs_dataset()        process_full_wgs_dataset(device)
    gpu.rs:33-94       is_gpu_available/get_gpu_devices  is_gpu_available(), get_gpu_devices()
    main.rs:48-192     main()                            main(argv) / build/rustseq_mini
Errors (the Err(String) arm) are raised as AlignerError with the same text.
"""
import ctypes

import numpy as np

from ._lib import SwbResult, load_library

GPU_WORK_GROUP_SIZE = 1024        # gpu.rs:9
GPU_MAX_WORK_GROUPS = 1_000_000   # gpu.rs:10


class AlignerError(RuntimeError):
    pass


class GpuDevice(ctypes.Structure):            # gpu.rs:17-23
    _fields_ = [("name", ctypes.c_char * 256), ("memory_gb", ctypes.c_float),
                ("max_work_group_size", ctypes.c_uint64), ("ordinal", ctypes.c_int32)]


class FileCheckpoint(ctypes.Structure):       # aligner.rs:23-32 (+ score64)
    _fields_ = [("file_path", ctypes.c_char * 1024), ("file_index", ctypes.c_uint64), ("score", ctypes.c_int32),
                ("score64", ctypes.c_int64), ("processing_time_ms", ctypes.c_double), ("total_bases", ctypes.c_uint64),
                ("total_reads", ctypes.c_uint64), ("completed", ctypes.c_int32)]


class GpuAlignmentResult(ctypes.Structure):   # gpu.rs:26-30
    _fields_ = [("score", ctypes.c_int32), ("score64", ctypes.c_int64), ("processing_time_ms", ctypes.c_double),
                ("gpu_device", ctypes.c_char * 256), ("total_reads", ctypes.c_uint64), ("total_bases", ctypes.c_uint64)]


_CHUNK_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64)
_sig_done = False


def _lib():
    global _sig_done
    lib = load_library()
    if not _sig_done:
        u64p = ctypes.POINTER(ctypes.c_uint64)
        lib.rsm_last_error.restype = ctypes.c_char_p
        lib.rsm_get_gpu_devices.argtypes = [ctypes.POINTER(GpuDevice), ctypes.c_int]
        lib.rsm_get_chunk_size_reads.argtypes = [u64p]
        lib.rsm_get_chunk_size_bases.argtypes = [u64p]
        lib.rsm_process_fastq_file_in_chunks.argtypes = [ctypes.c_char_p, ctypes.c_uint64, _CHUNK_FN, ctypes.c_void_p]
        lib.rsm_count_bases_in_fastq.argtypes = [ctypes.c_char_p, u64p]
        lib.rsm_debug_gunzip.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, u64p,
                                         ctypes.POINTER(ctypes.c_int)]
        lib.rsm_debug_pgunzip.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, u64p,
                                          ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), u64p, u64p]
        lib.rsm_debug_bgzf_segments.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_uint64, ctypes.c_uint, u64p, u64p, u64p, u64p,
                                                ctypes.POINTER(ctypes.c_int)]
        lib.rsm_gpu_align.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                      ctypes.POINTER(GpuDevice), ctypes.POINTER(ctypes.c_int32)]
        lib.rsm_gpu_align_ex.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                         ctypes.POINTER(GpuDevice), ctypes.POINTER(SwbResult)]
        lib.rsm_gpu_align_chunk_self.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(GpuDevice),
                                                 ctypes.POINTER(ctypes.c_int32)]
        lib.rsm_gpu_align_pair.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(GpuDevice),
                                           ctypes.POINTER(GpuAlignmentResult)]
        lib.rsm_process_full_wgs_dataset.argtypes = [ctypes.POINTER(GpuDevice), ctypes.POINTER(GpuAlignmentResult),
                                                     ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        lib.rsm_wgs_file_list.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]
        lib.rsm_main.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p)]
        lib.rsm_checkpoint_save.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(FileCheckpoint), ctypes.c_int, ctypes.c_uint64]
        lib.rsm_checkpoint_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(FileCheckpoint), ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_int), u64p]
        _sig_done = True
    return lib


def _check(rc):
    if rc != 0:
        raise AlignerError(_lib().rsm_last_error().decode("utf-8", "replace"))


def _bytes(x):
    if isinstance(x, str):
        x = x.encode("utf-8")
    return np.frombuffer(bytes(x), dtype=np.uint8)


def is_gpu_available():
    return bool(_lib().rsm_is_gpu_available())


def get_gpu_devices():
    arr = (GpuDevice * 64)()
    n = _lib().rsm_get_gpu_devices(arr, 64)
    return [arr[i] for i in range(min(n, 64))]


def get_chunk_size_reads():
    v = ctypes.c_uint64()
    _check(_lib().rsm_get_chunk_size_reads(ctypes.byref(v)))
    return int(v.value)


def get_chunk_size_bases():
    v = ctypes.c_uint64()
    _check(_lib().rsm_get_chunk_size_bases(ctypes.byref(v)))
    return int(v.value)


def process_fastq_file_in_chunks(filepath, chunk_size_reads, processor):
    """processor(list_of_bytes) is called per chunk; raise inside it to abort (the `?` of the reference)."""
    err = []

    def cb(_user, bases, offs, n):
        try:
            o = np.ctypeslib.as_array(ctypes.cast(offs, ctypes.POINTER(ctypes.c_uint64)), shape=(n + 1,))
            total = int(o[n])
            b = bytes(np.ctypeslib.as_array(ctypes.cast(bases, ctypes.POINTER(ctypes.c_uint8)), shape=(max(total, 1),))[:total])
            processor([b[int(o[k]):int(o[k + 1])] for k in range(n)])
            return 0
        except Exception as e:  # noqa: BLE001
            err.append(e)
            return 1

    rc = _lib().rsm_process_fastq_file_in_chunks(str(filepath).encode(), int(chunk_size_reads), _CHUNK_FN(cb), None)
    if err:
        raise err[0]
    _check(rc)


def debug_gunzip(filepath, read_cap=1 << 20, use_zlib=False, out_cap=None):
    """rsm_debug_gunzip (test hook): the file through the host gzip reader -> (bytes, n_delivered, failed)."""
    import os
    cap = int(out_cap) if out_cap is not None else max(1 << 16, 64 * os.path.getsize(filepath) + (1 << 20))
    buf = np.empty(cap, dtype=np.uint8)
    n = ctypes.c_uint64(); failed = ctypes.c_int()
    _check(_lib().rsm_debug_gunzip(str(filepath).encode(), int(read_cap), int(bool(use_zlib)), buf.ctypes.data, cap, ctypes.byref(n),
                                   ctypes.byref(failed)))
    return buf[: min(cap, int(n.value))].tobytes(), int(n.value), bool(failed.value)


def debug_pgunzip(filepath, threads, chunk_bytes=0, read_cap=1 << 20, out_cap=None):
    """rsm_debug_pgunzip (test hook): the file through the parallel host gzip reader ->
    (bytes, n_delivered, failed, dict(parallel, accepted, serial_stretches))."""
    import os
    cap = int(out_cap) if out_cap is not None else max(1 << 16, 64 * os.path.getsize(filepath) + (1 << 20))
    buf = np.empty(cap, dtype=np.uint8)
    n, acc, ser = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    failed, par = ctypes.c_int(), ctypes.c_int()
    _check(_lib().rsm_debug_pgunzip(str(filepath).encode(), int(read_cap), int(threads), int(chunk_bytes), buf.ctypes.data, cap, ctypes.byref(n),
                                    ctypes.byref(failed), ctypes.byref(par), ctypes.byref(acc), ctypes.byref(ser)))
    return (buf[: min(cap, int(n.value))].tobytes(), int(n.value), bool(failed.value),
            {"parallel": bool(par.value), "accepted": int(acc.value), "serial_stretches": int(ser.value)})


def debug_bgzf_segments(filepath, readers, seg_bytes, pool_buffers):
    """rsm_debug_bgzf_segments (test hook, no GPU): the --full-wgs BGZF readers against a consumer that only takes the segments
    in order -> dict(segments, blocks, text_bytes, hash, status)."""
    a, b, c, h = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    st = ctypes.c_int()
    _check(_lib().rsm_debug_bgzf_segments(str(filepath).encode(), int(readers), int(seg_bytes), int(pool_buffers), ctypes.byref(a),
                                          ctypes.byref(b), ctypes.byref(c), ctypes.byref(h), ctypes.byref(st)))
    return {"segments": int(a.value), "blocks": int(b.value), "text_bytes": int(c.value), "hash": int(h.value), "status": int(st.value)}


def count_bases_in_fastq(filepath):
    v = ctypes.c_uint64()
    _check(_lib().rsm_count_bases_in_fastq(str(filepath).encode(), ctypes.byref(v)))
    return int(v.value)


def gpu_align(seq1, seq2, device):
    a, b = _bytes(seq1), _bytes(seq2)
    s = ctypes.c_int32()
    _check(_lib().rsm_gpu_align(a.ctypes.data, a.size, b.ctypes.data, b.size, ctypes.byref(device), ctypes.byref(s)))
    return int(s.value)


def gpu_align_ex(seq1, seq2, device):
    a, b = _bytes(seq1), _bytes(seq2)
    r = SwbResult()
    _check(_lib().rsm_gpu_align_ex(a.ctypes.data, a.size, b.ctypes.data, b.size, ctypes.byref(device), ctypes.byref(r)))
    return int(r.score), int(r.end_i), int(r.end_j)


def gpu_align_chunk_self(chunk, device):
    a = _bytes(chunk)
    s = ctypes.c_int32()
    _check(_lib().rsm_gpu_align_chunk_self(a.ctypes.data, a.size, ctypes.byref(device), ctypes.byref(s)))
    return int(s.value)


def gpu_align_pair(file1, file2, device):
    r = GpuAlignmentResult()
    _check(_lib().rsm_gpu_align_pair(str(file1).encode(), str(file2).encode(), ctypes.byref(device), ctypes.byref(r)))
    return r


def wgs_file_list():
    buf = ctypes.create_string_buffer(1 << 20)
    n = ctypes.c_int()
    _check(_lib().rsm_wgs_file_list(buf, len(buf), ctypes.byref(n)))
    return [p for p in buf.value.decode().split("\n") if p]


def process_full_wgs_dataset(device):
    arr = (GpuAlignmentResult * 4096)()
    n = ctypes.c_int()
    _check(_lib().rsm_process_full_wgs_dataset(ctypes.byref(device), arr, 4096, ctypes.byref(n)))
    return [arr[i] for i in range(n.value)]


def checkpoint_save(path, run_id, files, total_files):
    """CheckpointState::save (aligner.rs:52-72): `files` is a list of FileCheckpoint."""
    arr = (FileCheckpoint * max(len(files), 1))(*files)
    _check(_lib().rsm_checkpoint_save(str(path).encode(), run_id.encode(), arr, len(files), total_files))


def checkpoint_load(path):
    """CheckpointState::load (aligner.rs:74-83): None when the file does not exist, else (run_id, files, total_files)."""
    arr = (FileCheckpoint * 4096)()
    rid = ctypes.create_string_buffer(256)
    n, tot = ctypes.c_int(), ctypes.c_uint64()
    _check(_lib().rsm_checkpoint_load(str(path).encode(), rid, 256, arr, 4096, ctypes.byref(n), ctypes.byref(tot)))
    if n.value < 0:
        return None
    return rid.value.decode(), [arr[i] for i in range(n.value)], int(tot.value)


def main(argv):
    args = [b"rustseq_mini"] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (ctypes.c_char_p * len(args))(*args)
    return int(_lib().rsm_main(len(args), arr))
