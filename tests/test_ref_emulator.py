"""The restatement against the reference's OWN kernel source, compiled unmodified by oracle/Makefile into
oracle/_ref/libref_cl.so (built in the container that has /root/reference; the .so travels to the GPU box).
Skipped when the .so is absent."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.skipif(ol.ref_cl() is None, reason="oracle/_ref/libref_cl.so not built")


def test_detailed_kernel_equals_last_row_max():
    rng = np.random.default_rng(11)
    al = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for _ in range(150):
        a = al[rng.integers(0, 5, int(rng.integers(1, 180)))]
        b = al[rng.integers(0, 5, int(rng.integers(1, 257)))]
        assert ol.ref_detailed(a, b, 256) == ol.last_row_max(a, b)


def _prefix_case(args):
    """One case of the prefix pin (runs in a worker process: the emulator keeps its state in globals).  Returns an error
    string or None.  full=True: every row prefix; else the prefix that ends at the oracle's end row, its neighbours, the
    whole read and a few random ones -- each must stay <= the score, and the end-row prefix must reach it exactly."""
    seed, alphabet, max1, max2, full = args
    rng = np.random.default_rng(seed)
    al = np.frombuffer(alphabet, dtype=np.uint8)
    a = al[rng.integers(0, al.size, int(rng.integers(1, max1 + 1)))]
    b = al[rng.integers(0, al.size, int(rng.integers(1, max2 + 1)))]
    if rng.random() < 0.5 and b.size > a.size:                      # related pair: the read is a mutated slice of the window
        o = int(rng.integers(0, b.size - a.size + 1))
        a = b[o:o + a.size].copy()
        m = rng.random(a.size) < 0.08
        a[m] = al[rng.integers(0, al.size, int(m.sum()))]
    score, end_i, _ = ol.sw_linear(a, b)
    wg = 1 << max(0, int(b.size - 1).bit_length())                  # power of two >= len2 (tree reduction, cl:139), <= 256
    if full:
        ks = range(1, a.size + 1)
    else:
        ks = {a.size, int(rng.integers(1, a.size + 1)), int(rng.integers(1, a.size + 1)), int(rng.integers(1, a.size + 1))}
        if score > 0:
            ks |= {end_i + 1, max(1, end_i), min(a.size, end_i + 2)}
    vals = {k: ol.ref_detailed(a[:k], b, wg) for k in ks}
    if max(vals.values()) > score:
        return f"seed {seed}: a prefix of the reference kernel scores {max(vals.values())} > sw_linear {score}"
    if full and max(vals.values()) != score:
        return f"seed {seed}: max over prefixes {max(vals.values())} != sw_linear {score}"
    if score > 0 and vals[end_i + 1] != score:
        return f"seed {seed}: prefix ending at the oracle's end row gives {vals[end_i + 1]}, sw_linear {score}"
    if score == 0 and any(vals.values()):
        return f"seed {seed}: sw_linear 0 but the reference kernel is positive"
    return None


def test_global_max_from_reference_kernel_on_prefixes():
    """sw_linear's score (and its end ROW) pinned through the reference kernel alone: the maximum of the dead kernel's
    last-row reduction over row prefixes.  420 small cases sweep every prefix; 180 cases up to 160 x 256 (the kernel's own
    column limit, cl:93-94) check the deciding prefixes; alphabets ACGT, ACGTN (N == N, cl:114), mixed case, homopolymer
    ties.  The emulator is single-threaded global state, so cases run in a process pool."""
    import multiprocessing as mpc
    import os
    alphabets = [b"ACGT", b"ACGTN", b"ACGTacgt", b"AC", b"A"]
    cases = [(1000 + k, alphabets[k % 5], 40, 128, True) for k in range(420)]
    cases += [(5000 + k, alphabets[k % 3], 160, 256, False) for k in range(180)]
    with mpc.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        errs = [e for e in pool.map(_prefix_case, cases, chunksize=8) if e]
    assert not errs, errs[:5]


@pytest.mark.parametrize("wg", [32, 64, 256])
def test_live_kernel_equals_ref_compat(wg):
    rng = np.random.default_rng(13)
    for _ in range(40):
        n1, n2 = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        a = rng.integers(0, 4, n1).astype(np.uint8) + 65
        b = rng.integers(0, 4, n2).astype(np.uint8) + (65 if rng.random() < 0.8 else 97)
        assert ol.ref_gpu_align(a, b, wg) == ol.ref_compat_align(a, b, wg)


def test_live_kernel_explicit_ndrange_multi_iteration():
    """Fewer groups than the host would launch => work-items walk several strided positions (cl:39-53)."""
    import ctypes
    rng = np.random.default_rng(14)
    a = rng.integers(0, 2, 4000).astype(np.uint8) + 65
    b = rng.integers(0, 2, 4000).astype(np.uint8) + 65
    out = ctypes.c_int32()
    assert ol.ref_cl().refcl_run_align(a.ctypes.data, b.ctypes.data, 4000, 16, 5, ctypes.byref(out)) == 0
    # restate that NDRange by hand
    chunk = (4000 + 4) // 5
    best = 0
    for g in range(5):
        for lid in range(16):
            cur = 0
            for i in range(g * chunk + lid, min((g + 1) * chunk, 4000), 16):
                cur = max(cur + (2 if a[i] == b[i] else -1), 0)
                best = max(best, cur)
    assert out.value == best
