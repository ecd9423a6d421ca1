"""Host-side twin of the device workload generator (csrc/swb_kernels.cu: synth_window_kernel / synth_read_kernel).

SURVEY.md 8d: counter RNG.  The k-th draw of stream ``seed`` for pair ``p`` is splitmix64 evaluated at state
``(seed ^ p*G) + (k+1)*G``; stream 0xB200 makes the windows, 0xB201 the reads.  Any shard, GPU or host regenerates
identical bytes independently.  distribution 0 = related (read cut from its window, ~1 % substitutions, 0.1 %
1-base insertions, 0.1 % 1-base deletions), 1 = unrelated.
"""
import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def splitmix_at(seed, p, k):
    with np.errstate(over="ignore"):
        p = np.asarray(p, dtype=np.uint64)
        k = np.asarray(k, dtype=np.uint64)
        z = (np.uint64(seed) ^ (p * _G)) + (k + np.uint64(1)) * _G
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _window_code(p, j):
    j = np.asarray(j, dtype=np.uint64)
    return (splitmix_at(0xB200, p, j >> np.uint64(5)) >> (np.uint64(2) * (j & np.uint64(31)))) & np.uint64(3)


def _reads_from_windows(pv, wcodes, read_len, window_len, distribution):
    """Reads of pairs pv (uint64 vector) made from their windows' 2-bit codes (n, window_len)."""
    n_pairs = pv.size
    p = pv[:, None]
    q = np.zeros((n_pairs, read_len), dtype=np.uint8)
    if distribution == 1:
        i = np.arange(read_len, dtype=np.uint64)[None, :]
        codes = (splitmix_at(0xB201, p, np.uint64(1) + (i >> np.uint64(5))) >> (np.uint64(2) * (i & np.uint64(31)))) & np.uint64(3)
        q[:] = _ACGT[codes.astype(np.intp)]
        return q
    span = np.uint64(window_len - read_len + 1 if window_len >= read_len else 1)
    c = splitmix_at(0xB201, pv, 0) % span
    rows = np.arange(n_pairs, dtype=np.intp)
    for i in range(read_len):
        x = splitmix_at(0xB201, pv, 1 + i)
        ev = x % np.uint64(1000)
        ins = ev == 0
        c = c + ((ev == 1) & ~ins).astype(np.uint64)
        inside = c < np.uint64(window_len)
        wc = wcodes[rows, np.minimum(c, np.uint64(max(window_len - 1, 0))).astype(np.intp)].astype(np.uint64) if window_len else np.zeros(n_pairs, dtype=np.uint64)
        code = np.where(inside, wc, (x >> np.uint64(34)) & np.uint64(3))
        sub = ((x >> np.uint64(10)) % np.uint64(100)) == 0
        code = np.where(sub, (code + np.uint64(1) + ((x >> np.uint64(20)) % np.uint64(3))) & np.uint64(3), code)
        code = np.where(ins, (x >> np.uint64(32)) & np.uint64(3), code)
        c = c + (~ins).astype(np.uint64)
        q[:, i] = _ACGT[code.astype(np.intp)]
    return q


def make_pairs(first_pair, n_pairs, read_len, window_len, distribution=0):
    """Returns (q_bytes, q_off, r_bytes, r_off): uint8 ASCII + uint64 CSR offsets, identical to swb_synth_device."""
    pv = np.uint64(first_pair) + np.arange(n_pairs, dtype=np.uint64)
    # windows: one draw of stream 0xB200 per 32 bases, 2 bits per base
    n_draws = (window_len + 31) // 32
    draws = splitmix_at(0xB200, pv[:, None], np.arange(n_draws, dtype=np.uint64)[None, :])          # (n, n_draws)
    sh = (np.uint64(2) * np.arange(32, dtype=np.uint64))[None, None, :]
    wcodes = ((draws[:, :, None] >> sh) & np.uint64(3)).astype(np.uint8).reshape(n_pairs, n_draws * 32)[:, :window_len]
    r = _ACGT[wcodes]
    q = _reads_from_windows(pv, wcodes, read_len, window_len, distribution)
    q_off = np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(read_len)
    r_off = np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(window_len)
    return q.reshape(-1), q_off, r.reshape(-1), r_off


def make_pairs_ref(reference, first_pair, n_pairs, read_len, window_len, distribution=0):
    """Twin of swb_synth_device_ref: window p is reference[ws : ws + window_len] with ws = draw(0xB202, 0) mod
    (len(reference) - window_len + 1); the read is made from it by the same rule as make_pairs.
    Returns (q_bytes, q_off, r_bytes, r_off, win_start)."""
    reference = np.ascontiguousarray(reference, dtype=np.uint8)
    pv = np.uint64(first_pair) + np.arange(n_pairs, dtype=np.uint64)
    ws = splitmix_at(0xB202, pv, 0) % np.uint64(reference.size - window_len + 1)
    r = reference[ws[:, None].astype(np.intp) + np.arange(window_len, dtype=np.intp)[None, :]]
    t = (r >> 1) & 3
    wcodes = (t ^ (t >> 1)).astype(np.uint8)                                                        # A C G T -> 0 1 2 3
    q = _reads_from_windows(pv, wcodes, read_len, window_len, distribution)
    q_off = np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(read_len)
    r_off = np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(window_len)
    return q.reshape(-1), q_off, r.reshape(-1), r_off, ws


def synth_reference(n):
    """n bases of the synthetic reference (same bytes as the --full-wgs driver's default, rustseq_host.cpp load_reference)."""
    k = np.arange((n + 31) // 32, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(0xB2F0) + k + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    sh = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    codes = ((x[:, None] >> sh) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n]
    return _ACGT[codes]
