"""New entry points of round 2, through the C ABI against the oracle:
swb_score_batch_ranges (windows = ranges of one host buffer, overlapping / repeated / unordered),
swb_create_multi + swb_multi_score_batch* (one batch split over several devices),
the hard window bound of swb_score_batch_device, and the reference-window workload generator."""
import numpy as np
import pytest

import oracle_lib as ol
import mini_parallel_b200 as mp
from mini_parallel_b200 import synth
from mini_parallel_b200.engine import to_csr

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _expected(q, qo, buf, start, length):
    r, ro = to_csr([buf[int(s):int(s) + int(n)] for s, n in zip(start, length)])
    return ol.batch(q, qo, r, ro, threads=8, simd=True)


def _reads_for(rng, buf, start, length, lo=1, hi=160):
    reads = []
    for s, n in zip(start, length):
        ln = int(rng.integers(lo, hi + 1))
        if n >= ln:
            o = int(s) + int(rng.integers(0, int(n) - ln + 1))
            rd = buf[o:o + ln].copy()
            m = rng.random(ln) < 0.03
            rd[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
        else:
            rd = ACGT[rng.integers(0, 4, ln)]
        reads.append(rd)
    return to_csr(reads)


def test_ranges_overlapping_unordered_windows(engine):
    rng = np.random.default_rng(2101)
    buf = ACGT[rng.integers(0, 4, 300_000)]
    n = 7000
    start = rng.integers(0, 300_000 - 700, n).astype(np.uint64)          # random order, heavy overlap (3.5 M window bases in 0.3 M)
    length = rng.integers(1, 700, n).astype(np.uint32)
    start[5], length[5] = 0, 1                                            # first byte of the buffer
    start[6], length[6] = 300_000 - 333, 333                              # up to its last byte
    start[7], length[7] = 123_456, 0                                      # empty window
    start[8] = start[9]; length[8] = length[9]                            # a repeated window
    q, qo = _reads_for(rng, buf, start, length)
    exp = _expected(q, qo, buf, start, length)
    try:
        for chunk_bytes, min_pairs in ((1 << 15, 1), (1 << 18, 500), (64 << 20, 16384)):
            engine.set_chunking(chunk_bytes, min_pairs)
            got = engine.score_batch_ranges(q, qo, buf, start, length)
            assert np.array_equal(got, exp), (chunk_bytes, min_pairs)
        info = engine.last_ranges_info()
        assert info["window_bytes"] == int(length.astype(np.uint64).sum())
        assert info["bytes_uploaded"] <= buf.size and info["bytes_uploaded"] < info["window_bytes"] // 5
    finally:
        engine.set_chunking(32 << 20, 16384)
    assert tuple(got[7]) == (0, -1, -1)


def test_ranges_touch_only_part_of_the_buffer_and_any_bytes(engine):
    rng = np.random.default_rng(2102)
    buf = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, 2_000_000)].copy()
    buf[1_000_000:1_000_400] = ord("a")                                    # lower case: raw byte compare (cl:114)
    n = 3000
    start = (rng.integers(999_000, 1_050_000, n)).astype(np.uint64)       # windows live in 51 kB of a 2 MB buffer, unaligned start
    length = np.full(n, 500, dtype=np.uint32)
    q, qo = _reads_for(rng, buf, start, length, 100, 200)                  # reads up to 200 bp: long-pair and byte kernels too
    got = engine.score_batch_ranges(q, qo, buf, start, length)
    assert np.array_equal(got, _expected(q, qo, buf, start, length))
    assert engine.last_ranges_info()["bytes_uploaded"] < 60_000
    # same pairs as CSR windows give the same results
    r, ro = to_csr([buf[int(s):int(s) + 500] for s in start])
    assert np.array_equal(engine.score_batch_csr(q, qo, r, ro), got)
    with pytest.raises(mp.SwbError, match="outside the buffer"):
        engine.score_batch_ranges(q, qo, buf, start + np.uint64(1_000_000), length)
    assert np.array_equal(engine.score_batch_ranges(q, qo, buf, start, length), got)   # still usable


def test_device_generator_with_reference_matches_host_twin(engine):
    ref = synth.synth_reference(300_000)
    n, rl, wl = 1500, 150, 500
    d = {k: engine.malloc_device(v) for k, v in (("ref", ref.size), ("q", n * rl), ("qo", (n + 1) * 8), ("r", n * wl), ("ro", (n + 1) * 8), ("ws", n * 8))}
    try:
        engine.h2d(d["ref"], ref, ref.size)
        for dist in (0, 1):
            engine.synth_device_ref(d["ref"], ref.size, 777, n, rl, wl, dist, d["q"], d["qo"], d["r"], d["ro"], d["ws"])
            engine.sync()
            q = np.zeros(n * rl, dtype=np.uint8); r = np.zeros(n * wl, dtype=np.uint8); ws = np.zeros(n, dtype=np.uint64)
            engine.d2h(q, d["q"], q.nbytes); engine.d2h(r, d["r"], r.nbytes); engine.d2h(ws, d["ws"], ws.nbytes)
            eq, _, er, _, ews = synth.make_pairs_ref(ref, 777, n, rl, wl, dist)
            assert np.array_equal(ws, ews) and np.array_equal(r, er) and np.array_equal(q, eq)
    finally:
        for p in d.values():
            engine.free_device(p)
    # and the ranges path on that workload equals the CSR path
    q, qo, r, ro, ws = synth.make_pairs_ref(ref, 0, 4000, rl, wl, 0)
    a = engine.score_batch_ranges(q, qo, ref, ws, np.full(4000, wl, dtype=np.uint32))
    assert np.array_equal(a, engine.score_batch_csr(q, qo, r, ro))
    assert np.array_equal(a, ol.batch(q, qo, r, ro, threads=8, simd=True))
    assert a["score"].min() > 200


def test_device_batch_refuses_windows_longer_than_the_declared_bound(engine):
    """swb_score_batch_device sizes the long-pair kernels' boundary rows from max_r_len before it has seen the offsets
    (ADVICE r01): an under-reported bound must not corrupt anything -- the pairs are refused and the sync says so."""
    rng = np.random.default_rng(2103)
    reads = [ACGT[rng.integers(0, 4, 300)] for _ in range(40)]
    wins = [ACGT[rng.integers(0, 4, 900 if k % 4 else 2500)] for k in range(40)]
    q, qo = to_csr(reads); r, ro = to_csr(wins)
    exp = ol.batch(q, qo, r, ro, threads=4, simd=True)
    dq, dqo, dr, dro, dout = (engine.malloc_device(x) for x in (q.size, qo.nbytes, r.size, ro.nbytes, 40 * 12))
    try:
        engine.h2d(dq, q, q.size); engine.h2d(dqo, qo, qo.nbytes); engine.h2d(dr, r, r.size); engine.h2d(dro, ro, ro.nbytes)
        out = np.zeros(40, dtype=mp.RESULT_DTYPE)
        engine.score_batch_device(dq, dqo, q.size, dr, dro, r.size, 40, 300, 2500, dout)       # honest bound
        engine.sync()
        engine.d2h(out, dout, out.nbytes)
        assert np.array_equal(out, exp)
        engine.score_batch_device(dq, dqo, q.size, dr, dro, r.size, 40, 300, 1000, dout)       # bound too small for ten pairs
        with pytest.raises(mp.SwbError, match="10 pairs have a window longer than max_r_len"):
            engine.sync()
        engine.d2h(out, dout, out.nbytes)
        short = np.array([k % 4 != 0 for k in range(40)])
        assert np.array_equal(out[short], exp[short])                                           # the others are exact
        assert np.all(out["score"][~short] == np.iinfo(np.int32).min)
        engine.score_batch_device(dq, dqo, q.size, dr, dro, r.size, 40, 0, 2500, dout)         # read-length hint unknown: fine
        engine.sync()
        engine.d2h(out, dout, out.nbytes)
        assert np.array_equal(out, exp)
    finally:
        for p in (dq, dqo, dr, dro, dout):
            engine.free_device(p)


@pytest.mark.parametrize("devices", [(0,), (0, 1), None])
def test_multi_device_batch_split(engine, devices):
    """One call, several devices: contiguous slices of equal bytes, disjoint slices of one result array.  On a 1-GPU box
    the (0, 1) case is skipped and None (= every visible device) is the single device."""
    if devices is not None and max(devices) >= mp.device_count():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    rng = np.random.default_rng(2104)
    reads = [ACGT[rng.integers(0, 4, int(rng.integers(1, 161)))] for _ in range(9000)]
    wins = [ACGT[rng.integers(0, 4, int(rng.integers(1, 900)))] for _ in range(9000)]
    reads[100] = np.frombuffer(b"ACGTN" * 40, dtype=np.uint8); reads[8999] = ACGT[rng.integers(0, 4, 700)]   # byte kernel, long kernel
    q, qo = to_csr(reads); r, ro = to_csr(wins)
    exp = ol.batch(q, qo, r, ro, threads=8, simd=True)
    me = mp.MultiEngine(devices)
    try:
        assert me.n_devices == (len(devices) if devices is not None else mp.device_count())
        assert np.array_equal(me.score_batch_csr(q, qo, r, ro), exp)
        assert np.array_equal(me.score_batch_csr(q, qo, r, ro), exp)                              # twice: the worker threads persist
        ref = ACGT[rng.integers(0, 4, 150_000)]
        me.set_reference(ref)
        start = rng.integers(0, 150_000 - 600, 9000).astype(np.uint64); wlen = rng.integers(1, 600, 9000).astype(np.uint32)
        q2, qo2 = _reads_for(rng, ref, start, wlen)
        assert np.array_equal(me.score_batch_vs_reference(q2, qo2, start, wlen), _expected(q2, qo2, ref, start, wlen))
        bad = qo.copy(); bad[4000] = bad[3999] - 1
        with pytest.raises(mp.SwbError, match="non-decreasing"):
            me.score_batch_csr(q, bad, r, ro)
        assert np.array_equal(me.score_batch_csr(q[:int(qo[3])], qo[:4], r[:int(ro[3])], ro[:4]), exp[:3])   # fewer pairs than devices is fine
    finally:
        me.close()
    assert np.array_equal(engine.score_batch_csr(q, qo, r, ro), exp)                              # the single-device context is unaffected
