"""Engine: one swb_ctx (one GPU).  Host-side convenience over the C ABI; all compute is in libswb200.so."""
import ctypes

import numpy as np

from ._lib import ALIGNMENT_DTYPE, RESULT_DTYPE, SwbResult, load_library


class SwbError(RuntimeError):
    """Mirrors the Err(String) arm of the reference's Result<i32, String> (aligner.rs:410)."""


def device_count():
    return int(load_library().swb_device_count())


def _as_bytes_array(x):
    if isinstance(x, np.ndarray):
        if x.dtype != np.uint8:
            raise TypeError("byte arrays must be uint8")
        return np.ascontiguousarray(x)
    if isinstance(x, str):
        x = x.encode("utf-8")
    return np.frombuffer(bytes(x), dtype=np.uint8)


def to_csr(seqs):
    """list of bytes/str -> (uint8 concatenation, uint64 offsets)."""
    arrs = [_as_bytes_array(s) for s in seqs]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    data = np.concatenate(arrs) if arrs and off[-1] > 0 else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8), off


def _check_csr(data, off, what):
    """The C ABI trusts that offsets stay inside the byte array it is handed (it only checks their order while it walks
    them): a safe wrapper must not let a bad offset array make the library read past the buffer."""
    if off.size == 0 or int(off[0]) != 0:
        raise ValueError(f"{what}: offsets must start at 0")
    if off.size > 1 and bool(np.any(off[1:] < off[:-1])):
        raise ValueError(f"{what}: offsets must be non-decreasing")
    if int(off[-1]) > data.size:
        raise ValueError(f"{what}: the last offset ({int(off[-1])}) is past the end of the {data.size}-byte array")


class Engine:
    def __init__(self, device=0, lib=None):
        self._lib = lib if lib is not None else load_library()      # lib: another build of the library (tests: the all-variants build)
        h = ctypes.c_void_p()
        if self._lib.swb_create(ctypes.byref(h), int(device), None) != 0:
            raise SwbError(self._err())
        self._h = h
        self.device = int(device)

    def _err(self):
        return self._lib.swb_last_error().decode("utf-8", "replace")

    def _check(self, rc):
        if rc != 0:
            raise SwbError(self._err())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.swb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- scoring ----
    def score_pair(self, s1, s2):
        a, b = _as_bytes_array(s1), _as_bytes_array(s2)
        res = SwbResult()
        self._check(self._lib.swb_score_pair(self._h, a.ctypes.data, a.size, b.ctypes.data, b.size, ctypes.byref(res)))
        return int(res.score), int(res.end_i), int(res.end_j)

    def score_batch_csr(self, q_bytes, q_off, r_bytes, r_off):
        q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8)
        r_bytes = np.ascontiguousarray(r_bytes, dtype=np.uint8)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        r_off = np.ascontiguousarray(r_off, dtype=np.uint64)
        n = q_off.size - 1
        if r_off.size - 1 != n:
            raise ValueError("q_off and r_off must describe the same number of pairs")
        if int(q_off[-1]) > q_bytes.size or int(r_off[-1]) > r_bytes.size or int(q_off[0]) != 0 or int(r_off[0]) != 0:
            _check_csr(q_bytes, q_off, "reads"); _check_csr(r_bytes, r_off, "windows")      # (order is checked by the library, with its own message)
        out = np.zeros(max(n, 0), dtype=RESULT_DTYPE)
        if n > 0:
            self._check(self._lib.swb_score_batch(self._h, q_bytes.ctypes.data, q_off.ctypes.data,
                                                  r_bytes.ctypes.data, r_off.ctypes.data, n, out.ctypes.data))
        return out

    def score_batch(self, reads, windows):
        q, qo = to_csr(reads)
        r, ro = to_csr(windows)
        return self.score_batch_csr(q, qo, r, ro)

    CIGAR_OPS = "MIDNSHP=X"

    def traceback_batch(self, q_bytes, q_off, r_bytes, r_off, results, cigar_cap=None):
        """Start cell and CIGAR behind the results of score_batch_csr on the same pairs (swb_traceback_batch).
        Returns (alignments[ALIGNMENT_DTYPE], ops): alignment k's operations are ops[cigar_off : cigar_off + cigar_len],
        words of length << 4 | op with op indexing CIGAR_OPS ('=' 7, 'X' 8, 'I' 1, 'D' 2)."""
        q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8)
        r_bytes = np.ascontiguousarray(r_bytes, dtype=np.uint8)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        r_off = np.ascontiguousarray(r_off, dtype=np.uint64)
        results = np.ascontiguousarray(results, dtype=RESULT_DTYPE)
        n = q_off.size - 1
        if r_off.size - 1 != n or results.size != n:
            raise ValueError("offsets and results must describe the same number of pairs")
        _check_csr(q_bytes, q_off, "reads"); _check_csr(r_bytes, r_off, "windows")
        out = np.zeros(max(n, 0), dtype=ALIGNMENT_DTYPE)
        cap = int(cigar_cap) if cigar_cap is not None else 8 * max(n, 1) + 1024
        used = ctypes.c_uint64()
        while n > 0:
            ops = np.zeros(cap, dtype=np.uint32)
            rc = self._lib.swb_traceback_batch(self._h, q_bytes.ctypes.data, q_off.ctypes.data, r_bytes.ctypes.data, r_off.ctypes.data, n,
                                               results.ctypes.data, out.ctypes.data, ops.ctypes.data, cap, ctypes.byref(used))
            if rc != 0 and used.value > cap and cigar_cap is None:
                cap = int(used.value)                              # the call says how much room the batch needs
                continue
            self._check(rc)
            return out, ops[:used.value]
        return out, np.zeros(0, dtype=np.uint32)

    def cigar_of(self, alignment, ops):
        """[(length, op), ...] of one alignment returned by traceback_batch."""
        a, n = int(alignment["cigar_off"]), int(alignment["cigar_len"])
        return [(int(v >> 4), self.CIGAR_OPS[int(v & 15)]) for v in ops[a:a + n]]

    def set_reference(self, ref):
        a = _as_bytes_array(ref)
        self._check(self._lib.swb_set_reference(self._h, a.ctypes.data, a.size))

    def score_batch_vs_reference(self, q_bytes, q_off, win_start, win_len):
        q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        win_start = np.ascontiguousarray(win_start, dtype=np.uint64)
        win_len = np.ascontiguousarray(win_len, dtype=np.uint32)
        n = q_off.size - 1
        if win_start.size != n or win_len.size != n:
            raise ValueError("q_off, win_start and win_len must describe the same number of pairs")
        _check_csr(q_bytes, q_off, "reads")
        out = np.zeros(max(n, 0), dtype=RESULT_DTYPE)
        if n > 0:
            self._check(self._lib.swb_score_batch_vs_reference(self._h, q_bytes.ctypes.data, q_off.ctypes.data, n,
                                                               win_start.ctypes.data, win_len.ctypes.data, out.ctypes.data))
        return out

    def score_batch_ranges(self, q_bytes, q_off, w_bytes, win_start, win_len):
        """Windows are ranges of ONE host buffer (they may overlap / repeat): swb_score_batch_ranges."""
        q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        w_bytes = np.ascontiguousarray(w_bytes, dtype=np.uint8)
        win_start = np.ascontiguousarray(win_start, dtype=np.uint64)
        win_len = np.ascontiguousarray(win_len, dtype=np.uint32)
        n = q_off.size - 1
        if win_start.size != n or win_len.size != n:
            raise ValueError("q_off, win_start and win_len must describe the same number of pairs")
        _check_csr(q_bytes, q_off, "reads")
        out = np.zeros(max(n, 0), dtype=RESULT_DTYPE)
        if n > 0:
            self._check(self._lib.swb_score_batch_ranges(self._h, q_bytes.ctypes.data, q_off.ctypes.data, n, w_bytes.ctypes.data, w_bytes.size,
                                                         win_start.ctypes.data, win_len.ctypes.data, out.ctypes.data))
        return out

    def guard_check(self):
        """SWB_GUARD=1 debugging aid (swb_debug_guard_check): (damaged guard zones, arenas checked, report); (-1, 0, "") when off."""
        buf = ctypes.create_string_buffer(4096)
        n = ctypes.c_uint64()
        bad = self._lib.swb_debug_guard_check(self._h, buf, 4096, ctypes.byref(n))
        return int(bad), int(n.value), buf.value.decode()

    def last_ranges_info(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self._check(self._lib.swb_last_ranges_info(self._h, ctypes.byref(a), ctypes.byref(b)))
        return {"bytes_uploaded": int(a.value), "window_bytes": int(b.value)}

    def fastq_bgzf_score(self, comp, blocks, carry=b"", final=True, file_index=0, first_read=0, window_len=500, carry_cap=1 << 16):
        """swb_fastq_bgzf_score: one segment of whole BGZF blocks -> (score sum, reads, bases, new carry, status).
        `blocks` is a sequence of (payload offset in comp, payload length, inflated length)."""
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        blk = np.zeros(len(blocks), dtype=np.dtype([("in_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4")]))
        for k, (o, n, m) in enumerate(blocks):
            blk[k] = (o, n, m)
        car = np.frombuffer(bytes(carry), dtype=np.uint8)
        cout = np.zeros(carry_cap, dtype=np.uint8)
        ssum, nr, nb, nl, cl, st = ctypes.c_int64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()
        self._check(self._lib.swb_fastq_bgzf_score(self._h, comp.ctypes.data if comp.size else None, comp.size,
                                                   blk.ctypes.data if len(blk) else None, len(blk),
                                                   car.ctypes.data if car.size else None, car.size, int(bool(final)),
                                                   int(file_index), int(first_read), int(window_len), ctypes.byref(ssum), ctypes.byref(nr),
                                                   ctypes.byref(nb), ctypes.byref(nl), cout.ctypes.data, carry_cap, ctypes.byref(cl), ctypes.byref(st)))
        return {"score_sum": int(ssum.value), "reads": int(nr.value), "bases": int(nb.value), "lines": int(nl.value), "carry": cout[: cl.value].tobytes(),
                "status": int(st.value)}

    def score_batch_device(self, d_q, d_qo, q_total, d_r, d_ro, r_total, n_pairs, max_q_len, max_r_len, d_out):
        """All pointers are device addresses (ints).  Asynchronous on the engine's stream."""
        self._check(self._lib.swb_score_batch_device(self._h, d_q, d_qo, q_total, d_r, d_ro, r_total,
                                                     n_pairs, max_q_len, max_r_len, d_out))

    def sync(self):
        self._check(self._lib.swb_sync(self._h))

    # ---- the reference's two literal behaviours ----
    def ref_compat_align(self, s1, s2, dev_max_work_group=1024):
        a, b = _as_bytes_array(s1), _as_bytes_array(s2)
        out = ctypes.c_int32()
        self._check(self._lib.swb_ref_compat_align(self._h, a.ctypes.data, a.size, b.ctypes.data, b.size,
                                                   dev_max_work_group, ctypes.byref(out)))
        return int(out.value)

    def last_row_max(self, s1, s2):
        a, b = _as_bytes_array(s1), _as_bytes_array(s2)
        out = ctypes.c_int32()
        self._check(self._lib.swb_last_row_max(self._h, a.ctypes.data, a.size, b.ctypes.data, b.size, ctypes.byref(out)))
        return int(out.value)

    # ---- stages ----
    def pack2bit(self, data):
        a = _as_bytes_array(data)
        nw = (a.size + 15) // 16
        words = np.zeros(nw, dtype=np.uint32)
        bitmap = np.zeros((nw + 31) // 32, dtype=np.uint32)
        if a.size:
            self._check(self._lib.swb_pack2bit(self._h, a.ctypes.data, a.size, words.ctypes.data, bitmap.ctypes.data))
        return words, bitmap

    def pack2bit_device(self, d_bytes, n, d_words, d_bitmap):
        self._check(self._lib.swb_pack2bit_device(self._h, d_bytes, n, d_words, d_bitmap))

    def synth_device(self, first_pair, n_pairs, read_len, window_len, distribution, d_q, d_qo, d_r, d_ro):
        self._check(self._lib.swb_synth_device(self._h, first_pair, n_pairs, read_len, window_len, distribution,
                                               d_q, d_qo, d_r, d_ro))

    def synth_device_ref(self, d_ref, ref_len, first_pair, n_pairs, read_len, window_len, distribution, d_q, d_qo, d_r, d_ro, d_ws):
        self._check(self._lib.swb_synth_device_ref(self._h, d_ref, ref_len, first_pair, n_pairs, read_len, window_len, distribution,
                                                   d_q, d_qo, d_r, d_ro, d_ws))

    def set_short_variant(self, v):
        self._check(self._lib.swb_set_short_variant(self._h, int(v)))

    def set_chunking(self, chunk_bytes, min_chunk_pairs=16384):
        """Chunk size of the pipelined host path (swb_set_chunking)."""
        self._check(self._lib.swb_set_chunking(self._h, int(chunk_bytes), int(min_chunk_pairs)))

    def set_chunk_ramp(self, ramp):
        """Chunk size schedule of the pipelined host path (swb_set_chunk_ramp): 0 equal, 1 auto (default), 2 always ramped."""
        self._check(self._lib.swb_set_chunk_ramp(self._h, int(ramp)))

    def last_timings(self):
        ms = (ctypes.c_float * 6)()
        k = ctypes.c_int()
        self._check(self._lib.swb_last_timings(self._h, ms, ctypes.byref(k)))
        names = ("pack_classify_ms", "short_ms", "generic_ms", "device_ms", "h2d_ms", "d2h_ms")
        d = {n: float(v) for n, v in zip(names, ms)}
        d["kernels"] = int(k.value)
        return d

    def last_routing(self):
        c = (ctypes.c_uint64 * 3)()
        self._check(self._lib.swb_last_routing(self._h, c))
        return {"short": int(c[0]), "generic": int(c[1]), "long": int(c[2])}

    def last_routing_ex(self):
        c = (ctypes.c_uint64 * 5)()
        self._check(self._lib.swb_last_routing_ex(self._h, c))
        return dict(zip(("short", "mid", "long", "bytes", "generic"), (int(v) for v in c)))

    def set_mid_path(self, on):
        self._check(self._lib.swb_set_mid_path(self._h, int(bool(on))))

    @property
    def stream(self):
        return self._lib.swb_stream(self._h)

    # ---- raw memory (for callers without torch) ----
    def malloc_device(self, nbytes):
        p = ctypes.c_void_p()
        self._check(self._lib.swb_malloc_device(self._h, int(nbytes), ctypes.byref(p)))
        return p.value

    def free_device(self, p):
        self._lib.swb_free_device(self._h, p)

    def d2h(self, dst_array, d_src, nbytes):
        self._check(self._lib.swb_memcpy_d2h(self._h, dst_array.ctypes.data, d_src, int(nbytes)))

    def h2d(self, d_dst, src_array, nbytes):
        self._check(self._lib.swb_memcpy_h2d(self._h, d_dst, src_array.ctypes.data, int(nbytes)))


class MultiEngine:
    """Several GPUs behind one handle (swb_create_multi): a host batch is split into contiguous slices, one per device."""

    def __init__(self, devices=None):
        self._lib = load_library()
        h = ctypes.c_void_p()
        if devices is None:
            rc = self._lib.swb_create_multi(ctypes.byref(h), None, 0, None)
        else:
            ids = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._lib.swb_create_multi(ctypes.byref(h), ids, len(devices), None)
        if rc != 0:
            raise SwbError(self._lib.swb_last_error().decode("utf-8", "replace"))
        self._h = h

    def _check(self, rc):
        if rc != 0:
            raise SwbError(self._lib.swb_last_error().decode("utf-8", "replace"))

    @property
    def n_devices(self):
        return int(self._lib.swb_multi_device_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.swb_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def score_batch_csr(self, q_bytes, q_off, r_bytes, r_off):
        q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8); r_bytes = np.ascontiguousarray(r_bytes, dtype=np.uint8)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64); r_off = np.ascontiguousarray(r_off, dtype=np.uint64)
        n = q_off.size - 1
        out = np.zeros(max(n, 0), dtype=RESULT_DTYPE)
        if n > 0:
            self._check(self._lib.swb_multi_score_batch(self._h, q_bytes.ctypes.data, q_off.ctypes.data, r_bytes.ctypes.data, r_off.ctypes.data, n, out.ctypes.data))
        return out

    def set_reference(self, ref):
        a = _as_bytes_array(ref)
        self._check(self._lib.swb_multi_set_reference(self._h, a.ctypes.data, a.size))

    def score_batch_vs_reference(self, q_bytes, q_off, win_start, win_len):
        q_bytes = np.ascontiguousarray(q_bytes, dtype=np.uint8); q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        win_start = np.ascontiguousarray(win_start, dtype=np.uint64); win_len = np.ascontiguousarray(win_len, dtype=np.uint32)
        n = q_off.size - 1
        out = np.zeros(max(n, 0), dtype=RESULT_DTYPE)
        if n > 0:
            self._check(self._lib.swb_multi_score_batch_vs_reference(self._h, q_bytes.ctypes.data, q_off.ctypes.data, n, win_start.ctypes.data,
                                                                     win_len.ctypes.data, out.ctypes.data))
        return out
