"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> libswb200.so), against the CPU
oracle on the same inputs.  Bit-exact: scores and end coordinates are integers."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import mini_parallel_b200 as mp
from mini_parallel_b200 import synth
from mini_parallel_b200.engine import to_csr

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sw_vectors.json")))["vectors"]
PRODUCT = "product"        # the default variant on the PRODUCT library; the numbered ones run on the test build
ALL_VARIANTS = (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11)
DEFAULT_VARIANT = 9          # sw_stream_kernel with the two-step tracker and dynamic couple distribution


def _vnum(variant):
    return DEFAULT_VARIANT if variant == PRODUCT else variant


def _rand(rng, n, alphabet=b"ACGT"):
    al = np.frombuffer(alphabet, dtype=np.uint8)
    return al[rng.integers(0, al.size, n)]


def _pairs(rng, n, rl, wl, related=True, alphabet=b"ACGT", mut=0.03):
    reads, wins = [], []
    for _ in range(n):
        n1, n2 = int(rng.integers(rl[0], rl[1] + 1)), int(rng.integers(wl[0], wl[1] + 1))
        w = _rand(rng, n2, alphabet)
        if related and n2 >= n1 > 0:
            o = int(rng.integers(0, n2 - n1 + 1))
            r = w[o:o + n1].copy()
            m = rng.random(n1) < mut
            r[m] = _rand(rng, int(m.sum()), alphabet)
        else:
            r = _rand(rng, n1, alphabet)
        reads.append(r)
        wins.append(w)
    return reads, wins


def _assert_parity(engine, reads, wins, threads=8):
    q, qo = to_csr(reads)
    r, ro = to_csr(wins)
    got = engine.score_batch_csr(q, qo, r, ro)
    exp = ol.batch(q, qo, r, ro, threads=threads, simd=False)
    bad = np.nonzero(got != exp)[0]
    assert bad.size == 0, f"{bad.size} pairs differ, first: pair {bad[0]} got {got[bad[0]]} expected {exp[bad[0]]}"
    return got


@pytest.mark.parametrize("v", GOLD, ids=[v["name"] for v in GOLD])
def test_golden_vectors(engine, v):
    a, b = v["seq1"].encode("latin1"), v["seq2"].encode("latin1")
    assert engine.score_pair(a, b) == (v["score"], v["end_i"], v["end_j"])
    assert engine.last_row_max(a, b) == v["last_row_max"]
    assert engine.ref_compat_align(a, b, 1024) == v["ref_compat_1024"]
    assert engine.ref_compat_align(a, b, 256) == v["ref_compat_256"]


def test_golden_vectors_as_one_batch(engine, engine_variants):
    reads = [v["seq1"].encode("latin1") for v in GOLD]
    wins = [v["seq2"].encode("latin1") for v in GOLD]
    for eng, variants in ((engine, (DEFAULT_VARIANT,)), (engine_variants, ALL_VARIANTS)):
        for variant in variants:
            eng.set_short_variant(variant)
            got = eng.score_batch(reads, wins)
            for g, v in zip(got, GOLD):
                assert (int(g["score"]), int(g["end_i"]), int(g["end_j"])) == (v["score"], v["end_i"], v["end_j"]), v["name"]
        eng.set_short_variant(DEFAULT_VARIANT)


def test_product_library_carries_the_default_variant_only(engine):
    for v in ALL_VARIANTS:
        if v != DEFAULT_VARIANT:
            with pytest.raises(mp.SwbError, match="not in this build"):
                engine.set_short_variant(v)
    engine.set_short_variant(DEFAULT_VARIANT)


@pytest.mark.parametrize("variant", (PRODUCT,) + ALL_VARIANTS)
def test_short_path_uniform_150x500(engine, engine_variants, variant):
    if variant != PRODUCT:
        engine = engine_variants
    rng = np.random.default_rng(100 + _vnum(variant))
    engine.set_short_variant(_vnum(variant))
    _assert_parity(engine, *_pairs(rng, 4001, (150, 150), (500, 500)))           # odd count: last group holds one pair
    assert engine.last_routing() == {"short": 4001, "generic": 0, "long": 0}
    _assert_parity(engine, *_pairs(rng, 2000, (150, 150), (500, 500), related=False))
    engine.set_short_variant(DEFAULT_VARIANT)


@pytest.mark.parametrize("variant", (PRODUCT,) + ALL_VARIANTS)
def test_short_path_ragged_lengths(engine, engine_variants, variant):
    if variant != PRODUCT:
        engine = engine_variants
    rng = np.random.default_rng(200 + _vnum(variant))
    engine.set_short_variant(_vnum(variant))
    _assert_parity(engine, *_pairs(rng, 6000, (1, 160), (1, 900)))
    _assert_parity(engine, *_pairs(rng, 1500, (140, 160), (1, 60), related=False))   # window shorter than the read
    engine.set_short_variant(DEFAULT_VARIANT)


@pytest.mark.parametrize("variant", (PRODUCT,) + ALL_VARIANTS)
def test_short_path_many_way_ties(engine, engine_variants, variant):
    if variant != PRODUCT:
        engine = engine_variants
    """Homopolymers and short repeats: every tie-break decision (min i, then min j) is exercised."""
    rng = np.random.default_rng(300 + _vnum(variant))
    engine.set_short_variant(_vnum(variant))
    _assert_parity(engine, *_pairs(rng, 1500, (1, 160), (1, 400), alphabet=b"A"))
    _assert_parity(engine, *_pairs(rng, 1500, (1, 160), (1, 400), related=False, alphabet=b"AC"))
    reads = [b"ACG" * 50] * 64 + [b"AT" * 80] * 64
    wins = [b"ACG" * 160] * 64 + [b"TA" * 250] * 64
    _assert_parity(engine, reads, wins)
    engine.set_short_variant(DEFAULT_VARIANT)


def test_short_path_128_row_instantiation(engine):
    """No read of the batch longer than 128 bp (2 x 100 / 2 x 125 bp runs): the 16 lanes x 8 rows instantiation of the stream
    kernel.  The choice follows the longest read of each chunk on the host paths and the caller's bound on the device path,
    where a read beyond the bound is routed past the kernel instead of losing its last rows."""
    rng = np.random.default_rng(128)
    for rl, wl in (((100, 100), (500, 500)), ((125, 125), (500, 500)), ((128, 128), (300, 700)), ((1, 128), (1, 900))):
        _assert_parity(engine, *_pairs(rng, 3001, rl, wl))
        assert engine.last_routing_ex()["short"] == 3001
    _assert_parity(engine, *_pairs(rng, 1500, (1, 128), (1, 600), alphabet=b"A"))                    # all ties
    _assert_parity(engine, *_pairs(rng, 1500, (1, 128), (1, 600), related=False, alphabet=b"AC"))
    got = _assert_parity(engine, [b"G" * 128] * 3 + [b"AC" * 64] * 30, [b"G" * 4096] * 3 + [b"CA" * 600] * 30)
    assert tuple(got[0]) == (256, 127, 127)
    r_a, w_a = _pairs(rng, 999, (90, 128), (200, 900))
    r_b, w_b = _pairs(rng, 1, (129, 129), (400, 400))                                            # one read of 129 bp: this batch runs on 160 rows
    _assert_parity(engine, r_a + r_b, w_a + w_b)
    assert engine.last_routing_ex()["short"] == 1000
    # device path with a bound that is too small: 150 bp reads declared as <= 100 are still scored exactly (by the long-pair kernel)
    reads, wins = _pairs(rng, 64, (150, 150), (500, 500))
    q, qo = to_csr(reads); r, ro = to_csr(wins)
    exp = ol.batch(q, qo, r, ro, threads=4, simd=True)
    dq, dqo, dr, dro, dout = (engine.malloc_device(x) for x in (q.size, qo.nbytes, r.size, ro.nbytes, 64 * 12))
    try:
        engine.h2d(dq, q, q.size); engine.h2d(dqo, qo, qo.nbytes); engine.h2d(dr, r, r.size); engine.h2d(dro, ro, ro.nbytes)
        out = np.zeros(64, dtype=mp.RESULT_DTYPE)
        for bound, path in ((100, "long"), (150, "short"), (0, "short")):
            engine.score_batch_device(dq, dqo, q.size, dr, dro, r.size, 64, bound, 500, dout)
            engine.sync()
            engine.d2h(out, dout, out.nbytes)
            assert np.array_equal(out, exp), bound
            assert engine.last_routing_ex()[path] == 64, (bound, engine.last_routing_ex())
    finally:
        for p in (dq, dqo, dr, dro, dout):
            engine.free_device(p)


def test_short_path_window_limits(engine):
    rng = np.random.default_rng(400)
    _assert_parity(engine, *_pairs(rng, 64, (150, 160), (4000, 4096)))               # longest window the short path takes
    assert engine.last_routing()["generic"] == 0
    _assert_parity(engine, *_pairs(rng, 16, (150, 160), (4097, 4200)))               # one past: long-pair kernel
    assert engine.last_routing() == {"short": 0, "generic": 0, "long": 16}


def test_generic_path_bytes_and_lengths(engine):
    rng = np.random.default_rng(500)
    _assert_parity(engine, *_pairs(rng, 1500, (1, 200), (1, 600), alphabet=b"ACGTN"))          # N == N matches (cl:114)
    _assert_parity(engine, *_pairs(rng, 400, (10, 150), (10, 500), alphabet=b"ACGTacgt"))       # case-sensitive
    _assert_parity(engine, *_pairs(rng, 100, (161, 900), (200, 3000)))                          # multi-band reads
    _assert_parity(engine, *_pairs(rng, 50, (255, 258), (255, 258)))                            # band boundary 256
    _assert_parity(engine, *_pairs(rng, 8, (2000, 2100), (5000, 5100)))
    reads = [bytes(rng.integers(0, 256, 200, dtype=np.uint8)) for _ in range(64)]               # arbitrary bytes
    wins = [bytes(rng.integers(0, 256, 300, dtype=np.uint8)) for _ in range(64)]
    _assert_parity(engine, reads, wins)


def test_mixed_batch_routing_and_empties(engine):
    rng = np.random.default_rng(600)
    r1, w1 = _pairs(rng, 300, (100, 160), (200, 600))
    r2, w2 = _pairs(rng, 60, (161, 400), (200, 900))
    r3, w3 = _pairs(rng, 40, (50, 150), (100, 400), alphabet=b"ACGTN")
    reads = r1 + r2 + r3 + [b"", b"ACGT", b""]
    wins = w1 + w2 + w3 + [b"ACGT", b"", b""]
    order = rng.permutation(len(reads))
    reads = [reads[k] for k in order]
    wins = [wins[k] for k in order]
    got = _assert_parity(engine, reads, wins)
    routing = engine.last_routing()
    assert routing["short"] + routing["generic"] + routing["long"] == len(reads) - 3
    ex = engine.last_routing_ex()                                                           # r2: reads of 161..320 bp -> mid, 321..400 -> long
    assert ex["short"] >= 200 and ex["bytes"] >= 25 and ex["mid"] >= 20 and ex["long"] >= 8  # N-flags are per 16-base word: neighbours of an N read may go to the byte kernel too
    assert ex["short"] + ex["mid"] == routing["short"] and ex["bytes"] + ex["generic"] == routing["generic"] and ex["long"] == routing["long"]
    for k, (a, b) in enumerate(zip(reads, wins)):
        if len(a) == 0 or len(b) == 0:
            assert tuple(got[k]) == (0, -1, -1)                                                 # aligner.rs:413-416


def test_long_kernel_band_and_stride_edges(engine):
    """sw_long_kernel: bands of 32*K rows streamed through one warp (K = 12 for ACGT pairs -> 384 rows, K = 10 for raw
    bytes -> 320, K = 6 -> 192 when every read of the batch is <= 192 bp).  Row counts around the band heights, windows
    shorter / longer than the minimum column stride (3 wavefronts), tall-and-narrow and short-and-wide pairs, many
    pairs per warp (work stealing), ties."""
    rng = np.random.default_rng(750)
    engine.set_mid_path(False)                                                      # reads of 161..320 bp through the long kernel too
    try:
        _long_kernel_edges(engine, rng)
    finally:
        engine.set_mid_path(True)


def _long_kernel_edges(engine, rng):
    _assert_parity(engine, *_pairs(rng, 40, (161, 170), (1, 400)))                  # one band, window < stride
    assert engine.last_routing()["long"] == 40
    _assert_parity(engine, *_pairs(rng, 60, (318, 323), (900, 1100)))               # band boundary 320, stride boundary 970
    _assert_parity(engine, *_pairs(rng, 60, (382, 387), (1100, 1200)))              # band boundary 384, stride boundary 1164
    _assert_parity(engine, *_pairs(rng, 30, (766, 771), (1150, 1180)))              # two / three bands of 384
    _assert_parity(engine, *_pairs(rng, 60, (318, 323), (900, 1100), alphabet=b"ACGTN"))   # the same edges on the byte kernel
    _assert_parity(engine, *_pairs(rng, 40, (161, 192), (1, 700), alphabet=b"ACGTN"))      # K = 6 byte kernel
    _assert_parity(engine, *_pairs(rng, 30, (639, 642), (950, 990)))                # two / three bands
    _assert_parity(engine, *_pairs(rng, 12, (1500, 2500), (1, 50), related=False))  # tall and narrow
    _assert_parity(engine, *_pairs(rng, 12, (161, 200), (6000, 9000)))              # short and wide
    _assert_parity(engine, *_pairs(rng, 3000, (161, 700), (100, 1500)))             # more pairs than resident warps
    assert engine.last_routing() == {"short": 0, "generic": 0, "long": 3000}
    _assert_parity(engine, *_pairs(rng, 200, (161, 1000), (1, 1200), alphabet=b"A"))                 # all ties
    _assert_parity(engine, *_pairs(rng, 200, (161, 1000), (1, 1200), related=False, alphabet=b"AC"))
    reads = [b"ACG" * 400] * 8 + [b"AT" * 500] * 8
    wins = [b"ACG" * 700] * 8 + [b"TA" * 900] * 8
    _assert_parity(engine, reads, wins)


def test_mid_path_reads_161_to_320(engine):
    """Reads of 161..320 bp: the 320-row instantiation of sw_stream_kernel (one group of 32 lanes x 10 rows per warp, value
    scale 32, 5 tag bits, blocks of 30 steps, keys with 9 row bits) -- the 2 x 250 / 2 x 300 bp read lengths."""
    rng = np.random.default_rng(760)
    _assert_parity(engine, *_pairs(rng, 3001, (250, 250), (500, 500)))              # odd count: the last couple holds one pair
    assert engine.last_routing_ex() == {"short": 0, "mid": 3001, "long": 0, "bytes": 0, "generic": 0}
    _assert_parity(engine, *_pairs(rng, 2000, (300, 300), (1000, 1000), related=False))
    _assert_parity(engine, *_pairs(rng, 5000, (161, 320), (1, 1200)))               # ragged: windows shorter than the read, than a wavefront
    assert engine.last_routing_ex()["mid"] == 5000
    # the list's longest read picks the instantiation on the device: <= 256 bp the 32 x 8-row one (256 rows), else 32 x 10 (320)
    _assert_parity(engine, *_pairs(rng, 4001, (161, 256), (1, 1200)))               # 256-row kernel, ragged
    assert engine.last_routing_ex()["mid"] == 4001
    _assert_parity(engine, *_pairs(rng, 1000, (256, 256), (300, 700)))              # its last row in use
    _assert_parity(engine, *_pairs(rng, 1000, (161, 256), (1, 600), alphabet=b"A"))             # all ties on 256 rows
    r_a, w_a = _pairs(rng, 999, (161, 256), (200, 900))
    r_b, w_b = _pairs(rng, 1, (257, 257), (400, 400))                               # one read of 257 bp: the whole list moves to 320 rows
    _assert_parity(engine, r_a + r_b, w_a + w_b)
    reads256 = [b"G" * 256] * 3 + [b"AC" * 128] * 30
    got256 = _assert_parity(engine, reads256, [b"G" * 4096] * 3 + [b"CA" * 600] * 30)
    assert tuple(got256[0]) == (512, 255, 255)                                       # the highest score the 256-row kernel can see
    _assert_parity(engine, *_pairs(rng, 300, (310, 320), (3900, 4096)))             # the longest windows the path takes
    assert engine.last_routing_ex()["mid"] == 300
    _assert_parity(engine, *_pairs(rng, 40, (310, 320), (4097, 4200)))              # one past: long-pair kernel
    assert engine.last_routing_ex() == {"short": 0, "mid": 0, "long": 40, "bytes": 0, "generic": 0}
    _assert_parity(engine, *_pairs(rng, 40, (321, 330), (500, 600)))                # one row past
    assert engine.last_routing_ex()["long"] == 40
    _assert_parity(engine, *_pairs(rng, 1500, (161, 320), (1, 700), alphabet=b"A"))                   # all ties: min i, then min j
    _assert_parity(engine, *_pairs(rng, 1500, (161, 320), (1, 700), related=False, alphabet=b"AC"))
    reads = [b"ACG" * 100] * 33 + [b"AT" * 160] * 33 + [b"G" * 320] * 3
    wins = [b"ACG" * 300] * 33 + [b"TA" * 400] * 33 + [b"G" * 4096] * 3
    got = _assert_parity(engine, reads, wins)
    assert tuple(got[-1]) == (640, 319, 319)                                        # the highest score the path can see
    # short, mid, long and byte-kernel pairs in one batch, shuffled
    r1, w1 = _pairs(rng, 700, (100, 160), (200, 600))
    r2, w2 = _pairs(rng, 700, (161, 320), (200, 900))
    r3, w3 = _pairs(rng, 50, (321, 600), (200, 900))
    r4, w4 = _pairs(rng, 50, (200, 300), (300, 600), alphabet=b"ACGTN")
    reads, wins = r1 + r2 + r3 + r4, w1 + w2 + w3 + w4
    order = rng.permutation(len(reads))
    _assert_parity(engine, [reads[k] for k in order], [wins[k] for k in order])
    ex = engine.last_routing_ex()
    # (the N-flags are per 16-base word of the concatenated arrays: a neighbour of an N read may take the byte kernel too)
    assert sum(ex.values()) == 1500 and ex["short"] >= 600 and ex["mid"] >= 600 and ex["long"] >= 40 and 50 <= ex["bytes"] <= 250 and ex["generic"] == 0
    # the same pairs through the long kernel give the same results (two independent implementations)
    a = engine.score_batch(r2, w2)
    engine.set_mid_path(False)
    try:
        b = engine.score_batch(r2, w2)
        assert engine.last_routing_ex()["long"] == 700
    finally:
        engine.set_mid_path(True)
    assert np.array_equal(a, b)


def test_long_identical_pair_needs_32bit(engine):
    rng = np.random.default_rng(700)
    s = bytes(_rand(rng, 16500))
    assert engine.score_pair(s, s) == (33000, 16499, 16499)


def test_long_pair_10k(engine):
    """Config 4 shape (10 kb x 10 kb), a few pairs: the oracle needs ~0.3 s per pair."""
    rng = np.random.default_rng(701)
    _assert_parity(engine, *_pairs(rng, 3, (10000, 10000), (10000, 10000), mut=0.1), threads=3)


def test_results_do_not_depend_on_batch_composition(engine):
    rng = np.random.default_rng(800)
    reads, wins = _pairs(rng, 999, (1, 160), (1, 700))
    whole = engine.score_batch(reads, wins)
    perm = rng.permutation(len(reads))
    shuffled = engine.score_batch([reads[k] for k in perm], [wins[k] for k in perm])
    assert np.array_equal(whole[perm], shuffled)
    parts = np.concatenate([engine.score_batch(reads[a:a + 100], wins[a:a + 100]) for a in range(0, 999, 100)])
    assert np.array_equal(whole, parts)
    again = engine.score_batch(reads, wins)
    assert np.array_equal(whole, again)                                                         # idempotent


def test_pack2bit_matches_numpy(engine):
    rng = np.random.default_rng(900)
    for n in (1, 15, 16, 17, 511, 512, 513, 100003):
        data = _rand(rng, n, b"ACGT").copy()
        if n > 40:
            data[rng.integers(0, n, 5)] = ord("N")
            data[rng.integers(0, n)] = ord("a")
        words, bitmap = engine.pack2bit(data)
        codes = (data >> 1) & 3
        pad = np.zeros((-n) % 16, dtype=np.uint8)
        c = np.concatenate([codes, pad]).reshape(-1, 16).astype(np.uint32)
        exp_words = (c << (2 * np.arange(16, dtype=np.uint32))[None, :]).sum(axis=1).astype(np.uint32)
        assert np.array_equal(words, exp_words)
        okb = np.isin(data, np.frombuffer(b"ACGT", dtype=np.uint8))
        bad_word = ~np.concatenate([okb, np.ones((-n) % 16, dtype=bool)]).reshape(-1, 16).all(axis=1)
        got_bad = ((bitmap[np.arange(words.size) // 32] >> (np.arange(words.size) % 32).astype(np.uint32)) & 1).astype(bool)
        assert np.array_equal(got_bad, bad_word)


@pytest.mark.parametrize("dist", [0, 1])
def test_device_generator_matches_host_twin(engine, dist):
    n, rl, wl = 777, 150, 500
    dq, dqo = engine.malloc_device(n * rl), engine.malloc_device((n + 1) * 8)
    dr, dro = engine.malloc_device(n * wl), engine.malloc_device((n + 1) * 8)
    try:
        engine.synth_device(12345, n, rl, wl, dist, dq, dqo, dr, dro)
        engine.sync()
        q = np.zeros(n * rl, dtype=np.uint8); r = np.zeros(n * wl, dtype=np.uint8)
        qo = np.zeros(n + 1, dtype=np.uint64); ro = np.zeros(n + 1, dtype=np.uint64)
        engine.d2h(q, dq, q.nbytes); engine.d2h(r, dr, r.nbytes); engine.d2h(qo, dqo, qo.nbytes); engine.d2h(ro, dro, ro.nbytes)
    finally:
        for p in (dq, dqo, dr, dro):
            engine.free_device(p)
    eq, eqo, er, ero = synth.make_pairs(12345, n, rl, wl, dist)
    assert np.array_equal(qo, eqo) and np.array_equal(ro, ero)
    assert np.array_equal(r, er)
    assert np.array_equal(q, eq)


def test_config2_full_size_device_resident(engine):
    """BASELINE.json configs[1] at full size (1 M pairs, 150 x 500), inputs generated and kept in HBM.  EVERY pair is
    compared with the CPU SIMD oracle (bit-exact with sw_linear, tests/test_oracle.py); plus the size-independent
    properties and the shard == whole check."""
    n, rl, wl = 1_000_000, 150, 500
    dq, dqo = engine.malloc_device(n * rl), engine.malloc_device((n + 1) * 8)
    dr, dro = engine.malloc_device(n * wl), engine.malloc_device((n + 1) * 8)
    dout = engine.malloc_device(n * 12)
    try:
        engine.synth_device(0, n, rl, wl, 0, dq, dqo, dr, dro)
        engine.score_batch_device(dq, dqo, n * rl, dr, dro, n * wl, n, rl, wl, dout)
        engine.sync()
        out = np.zeros(n, dtype=mp.RESULT_DTYPE)
        engine.d2h(out, dout, out.nbytes)
        assert engine.last_routing() == {"short": n, "generic": 0, "long": 0}
        assert out["score"].min() > 200 and out["score"].max() <= 2 * rl            # related reads score high
        assert np.all(out["end_i"] < rl) and np.all(out["end_j"] < wl) and np.all(out["end_i"] >= 0)
        q = np.zeros(n * rl, dtype=np.uint8); r = np.zeros(n * wl, dtype=np.uint8)
        engine.d2h(q, dq, q.nbytes); engine.d2h(r, dr, r.nbytes)
        qo = np.arange(n + 1, dtype=np.uint64) * rl; ro = np.arange(n + 1, dtype=np.uint64) * wl
        exp = ol.batch(q, qo, r, ro, threads=os.cpu_count() or 8, simd=True)         # all 1 M pairs, ~0.5 s of host cores
        bad = np.nonzero(out != exp)[0]
        assert bad.size == 0, f"{bad.size} of {n} pairs differ, first: pair {bad[0]} got {out[bad[0]]} expected {exp[bad[0]]}"
        # the unrelated distribution too (scores ~20: the floor path), all pairs
        engine.synth_device(0, n, rl, wl, 1, dq, dqo, dr, dro)
        engine.score_batch_device(dq, dqo, n * rl, dr, dro, n * wl, n, rl, wl, dout)
        engine.sync()
        out1 = np.zeros(n, dtype=mp.RESULT_DTYPE)
        engine.d2h(out1, dout, out1.nbytes)
        engine.d2h(q, dq, q.nbytes); engine.d2h(r, dr, r.nbytes)
        exp1 = ol.batch(q, qo, r, ro, threads=os.cpu_count() or 8, simd=True)
        bad = np.nonzero(out1 != exp1)[0]
        assert bad.size == 0, f"unrelated reads: {bad.size} of {n} pairs differ, first: pair {bad[0]} got {out1[bad[0]]} expected {exp1[bad[0]]}"
        # sharded == whole: score the second half alone (what a second rank would do) and compare
        h = n // 2
        engine.synth_device(h, n - h, rl, wl, 0, dq, dqo, dr, dro)
        engine.score_batch_device(dq, dqo, (n - h) * rl, dr, dro, (n - h) * wl, n - h, rl, wl, dout)
        engine.sync()
        half = np.zeros(n - h, dtype=mp.RESULT_DTYPE)
        engine.d2h(half, dout, half.nbytes)
        assert np.array_equal(half, out[h:])
    finally:
        for p in (dq, dqo, dr, dro, dout):
            engine.free_device(p)


def test_host_path_is_chunked_over_streams_and_stays_exact(engine):
    """swb_score_batch cuts a host batch into chunks pipelined over three streams; every chunk boundary
    (offset rebase, result placement, routing sums) must be invisible in the results."""
    rng = np.random.default_rng(900)
    reads, wins = _pairs(rng, 5000, (1, 160), (1, 700))
    reads[17] = b""                                               # empty read inside a chunk
    wins[4099] = b""                                              # empty window at a chunk boundary (1024-pair chunks)
    reads[2048] = b"ACGTN" * 30                                   # generic path inside a chunk
    q, qo = to_csr(reads)
    r, ro = to_csr(wins)
    exp = ol.batch(q, qo, r, ro, threads=8, simd=False)
    try:
        for ramp in (0, 2):                                       # equal chunks / sizes ramping up and down again
            engine.set_chunk_ramp(ramp)
            for chunk_bytes, min_pairs in ((1 << 14, 1), (1 << 16, 1024), (1 << 20, 999), (64 << 20, 16384)):
                engine.set_chunking(chunk_bytes, min_pairs)
                got = engine.score_batch_csr(q, qo, r, ro)
                assert np.array_equal(got, exp), (ramp, chunk_bytes, min_pairs)
                rt = engine.last_routing()
                assert rt["short"] + rt["generic"] + rt["long"] == 4998, rt      # the two empty pairs are routed nowhere
                assert rt["generic"] >= 1
    finally:
        engine.set_chunking(32 << 20, 16384)
        engine.set_chunk_ramp(1)


def test_uniform_length_chunks_generate_their_offsets_on_the_device(engine):
    """A chunk whose reads (or windows) all have one length uploads no offsets for that side; every mix of uniform and
    ragged sides, and a single odd pair that makes a chunk ragged, must give the oracle's results."""
    rng = np.random.default_rng(902)
    n = 6000
    cases = {
        "both uniform": ([_rand(rng, 150) for _ in range(n)], [_rand(rng, 500) for _ in range(n)]),
        "reads uniform": ([_rand(rng, 100) for _ in range(n)], [_rand(rng, int(rng.integers(1, 600))) for _ in range(n)]),
        "windows uniform": ([_rand(rng, int(rng.integers(1, 160))) for _ in range(n)], [_rand(rng, 333) for _ in range(n)]),
    }
    odd_r, odd_w = [_rand(rng, 150) for _ in range(n)], [_rand(rng, 500) for _ in range(n)]
    odd_r[4500] = _rand(rng, 149); odd_w[17] = _rand(rng, 501)         # one chunk of each side is ragged, the others are not
    cases["one odd pair"] = (odd_r, odd_w)
    try:
        engine.set_chunking(1 << 18, 500)                              # several chunks per batch
        for name, (reads, wins) in cases.items():
            q, qo = to_csr(reads); r, ro = to_csr(wins)
            exp = ol.batch(q, qo, r, ro, threads=8, simd=True)
            assert np.array_equal(engine.score_batch_csr(q, qo, r, ro), exp), name
        ref = _rand(rng, 100_000)
        engine.set_reference(ref)
        start = rng.integers(0, 100_000 - 500, n).astype(np.uint64); wlen = np.full(n, 500, dtype=np.uint32)
        reads = [ref[int(s) + 100:int(s) + 250].copy() for s in start]
        q, qo = to_csr(reads); r, ro = to_csr([ref[int(s):int(s) + 500] for s in start])
        assert np.array_equal(engine.score_batch_vs_reference(q, qo, start, wlen), ol.batch(q, qo, r, ro, threads=8, simd=True))
    finally:
        engine.set_chunking(32 << 20, 16384)


def test_reference_windows_chunked(engine):
    rng = np.random.default_rng(901)
    ref = _rand(rng, 200_000)
    n = 6000
    start = rng.integers(0, 200_000 - 600, n).astype(np.uint64)
    wlen = rng.integers(1, 600, n).astype(np.uint32)
    reads = []
    for k in range(n):
        o = int(start[k]) + int(rng.integers(0, max(1, int(wlen[k]) - 100)))
        rd = ref[o:o + int(rng.integers(1, 160))].copy()
        reads.append(rd)
    q, qo = to_csr(reads)
    wins = [ref[int(s):int(s) + int(l)] for s, l in zip(start, wlen)]
    r, ro = to_csr(wins)
    exp = ol.batch(q, qo, r, ro, threads=8, simd=False)
    engine.set_reference(ref)
    try:
        for ramp in (0, 1):                                       # 1 (the default) ramps batches against the resident reference
            engine.set_chunk_ramp(ramp)
            for chunk_bytes, min_pairs in ((1 << 15, 1), (1 << 17, 300), (64 << 20, 16384)):
                engine.set_chunking(chunk_bytes, min_pairs)
                got = engine.score_batch_vs_reference(q, qo, start, wlen)
                assert np.array_equal(got, exp), (ramp, chunk_bytes, min_pairs)
    finally:
        engine.set_chunking(32 << 20, 16384)
        engine.set_chunk_ramp(1)


def test_config4_full_size_long_pairs(engine):
    """BASELINE.json configs[3] at full size: 10 000 pairs of 10 kb x 10 kb (1e12 cells) through sw_long_kernel,
    inputs generated and kept in HBM.  EVERY pair is compared with the CPU SIMD oracle (1e12 cells, seconds on the box's
    host cores); plus the size-independent properties: an identical pair scores 2n and ends in the last cell, the score
    is symmetric in its arguments, bounds."""
    n, L = 10_000, 10_000
    dq, dqo = engine.malloc_device(n * L), engine.malloc_device((n + 1) * 8)
    dr, dro = engine.malloc_device(n * L), engine.malloc_device((n + 1) * 8)
    dout = engine.malloc_device(n * 12)
    try:
        engine.synth_device(0, n, L, L, 0, dq, dqo, dr, dro)            # related: read = window with ~1.2 % edits
        engine.score_batch_device(dq, dqo, n * L, dr, dro, n * L, n, L, L, dout)
        engine.sync()
        out = np.zeros(n, dtype=mp.RESULT_DTYPE)
        engine.d2h(out, dout, out.nbytes)
        assert engine.last_routing() == {"short": 0, "generic": 0, "long": n}
        assert out["score"].min() > 18_000 and out["score"].max() <= 2 * L
        assert np.all(out["end_i"] < L) and np.all(out["end_j"] < L) and np.all(out["end_i"] > 9_000)
        q = np.zeros(n * L, dtype=np.uint8); r = np.zeros(n * L, dtype=np.uint8)
        engine.d2h(q, dq, q.nbytes); engine.d2h(r, dr, r.nbytes)
        off = np.arange(n + 1, dtype=np.uint64) * L
        exp = ol.batch(q, off, r, off, threads=os.cpu_count() or 8, simd=True)      # all 10 000 pairs
        bad = np.nonzero(out != exp)[0]
        assert bad.size == 0, f"{bad.size} of {n} pairs differ, first: pair {bad[0]} got {out[bad[0]]} expected {exp[bad[0]]}"
        for k in range(2):                                                           # and the scalar oracle on two of them
            assert tuple(out[k]) == ol.sw_linear(q[k * L:(k + 1) * L].tobytes(), r[k * L:(k + 1) * L].tobytes())
        # symmetry: swapping reads and windows leaves every score unchanged (coordinates follow the tie-break, not checked)
        engine.score_batch_device(dr, dro, n * L, dq, dqo, n * L, n, L, L, dout)
        engine.sync()
        swapped = np.zeros(n, dtype=mp.RESULT_DTYPE)
        engine.d2h(swapped, dout, swapped.nbytes)
        assert np.array_equal(swapped["score"], out["score"])
        # identity: every window against itself scores 2n and ends in the last cell
        engine.score_batch_device(dr, dro, n * L, dr, dro, n * L, n, L, L, dout)
        engine.sync()
        ident = np.zeros(n, dtype=mp.RESULT_DTYPE)
        engine.d2h(ident, dout, ident.nbytes)
        assert np.all(ident["score"] == 2 * L) and np.all(ident["end_i"] == L - 1) and np.all(ident["end_j"] == L - 1)
    finally:
        for p in (dq, dqo, dr, dro, dout):
            engine.free_device(p)


def test_error_in_a_later_chunk_leaves_the_engine_usable(engine):
    """A malformed offset found while chunk 3 is being validated aborts the call after chunks 0-2 were enqueued; the call
    must drain them (inputs are borrowed for the call only) and the context must keep working."""
    rng = np.random.default_rng(950)
    reads, wins = _pairs(rng, 4000, (100, 160), (200, 600))
    q, qo = to_csr(reads)
    r, ro = to_csr(wins)
    bad = qo.copy()
    bad[3500] = bad[3499] - 1                                     # non-monotone offset inside the last chunk
    try:
        engine.set_chunking(1 << 17, 1000)
        with pytest.raises(mp.SwbError, match="non-decreasing"):
            engine.score_batch_csr(q, bad, r, ro)
        got = engine.score_batch_csr(q, qo, r, ro)
        assert np.array_equal(got, ol.batch(q, qo, r, ro, threads=8, simd=True))
    finally:
        engine.set_chunking(32 << 20, 16384)


def test_repeated_runs_are_bit_identical(engine):
    """Stand-in for racecheck (compute-sanitizer is closed on the GPU pool): the kernels exchange data through shared memory
    (window ring, parked trackers, couple ids) and through per-warp global scratch rows; a missing barrier or a stale read
    shows up as run-to-run variation long before it shows up as a wrong answer on one fixed schedule.  The same mixed batch,
    20 times over, alone and while a second batch keeps other lanes busy: every run equals the first, which equals the oracle."""
    rng = np.random.default_rng(2027)
    r1, w1 = _pairs(rng, 6000, (1, 160), (1, 700))
    r2, w2 = _pairs(rng, 1500, (161, 320), (100, 900))
    r3, w3 = _pairs(rng, 30, (321, 1400), (300, 2600))
    r4, w4 = _pairs(rng, 60, (1, 500), (1, 900), alphabet=b"ACGTNacgt")
    reads, wins = r1 + r2 + r3 + r4, w1 + w2 + w3 + w4
    order = rng.permutation(len(reads))
    reads, wins = [reads[k] for k in order], [wins[k] for k in order]
    q, qo = to_csr(reads); r, ro = to_csr(wins)
    first = engine.score_batch_csr(q, qo, r, ro)
    assert np.array_equal(first, ol.batch(q, qo, r, ro, threads=8, simd=False))
    engine.set_chunking(256 << 10, 64)                      # many small chunks over the three lanes: kernels of different chunks overlap
    try:
        for rep in range(20):
            got = engine.score_batch_csr(q, qo, r, ro)
            assert np.array_equal(got, first), f"run {rep} differs from the first in {int((got != first).sum())} pairs"
    finally:
        engine.set_chunking(32 << 20)
    al0, ops0 = engine.traceback_batch(q, qo, r, ro, first)
    c0 = [engine.cigar_of(al0[k], ops0) for k in range(0, len(reads), 7)]
    for rep in range(5):
        al, ops = engine.traceback_batch(q, qo, r, ro, first)
        assert np.array_equal(al["start_i"], al0["start_i"]) and np.array_equal(al["start_j"], al0["start_j"])
        assert [engine.cigar_of(al[k], ops) for k in range(0, len(reads), 7)] == c0
