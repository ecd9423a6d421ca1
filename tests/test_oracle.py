"""CPU tests of the CHECKER: the C oracle against the committed golden vectors (generated with the
reference's own kernels executed through oracle/_ref), against an independent NumPy restatement, and
against the properties Smith-Waterman scores must satisfy."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle_lib as ol
import sw_numpy

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sw_vectors.json")))["vectors"]


@pytest.mark.parametrize("v", GOLD, ids=[v["name"] for v in GOLD])
def test_oracle_matches_golden(v):
    a, b = v["seq1"].encode("latin1"), v["seq2"].encode("latin1")
    assert ol.sw_linear(a, b) == (v["score"], v["end_i"], v["end_j"])
    assert ol.last_row_max(a, b) == v["last_row_max"]
    assert ol.ref_compat_align(a, b, 1024) == v["ref_compat_1024"]
    assert ol.ref_compat_align(a, b, 256) == v["ref_compat_256"]
    # columns produced by the reference's own kernels (make_golden.py): they must agree with the restatement
    if v["ref_detailed"] is not None:
        assert v["ref_detailed"] == v["last_row_max"]
    if v["ref_gpu_align"] is not None:
        assert v["ref_gpu_align"] == v["ref_compat_256"]


def test_survey_table_literal():
    """SURVEY.md 8c golden table, spelled out (row 1 is README.md:7-11 of the reference)."""
    table = [(b"ATCGT", b"ATTGG", 5, (3, 3), 4, 2), (b"ACGT", b"ACGT", 8, (3, 3), 8, 2), (b"AAAA", b"TTTT", 0, (-1, -1), 0, 0),
             (b"ACGTACGT", b"ACGACGT", 12, (7, 6), 12, 2), (b"GATTACA", b"GCATGCU", 4, (2, 3), 4, 2),
             (b"TGTTACGG", b"GGTTGACTA", 8, (5, 6), 6, 2)]
    for a, b, s, end, lrm, rc in table:
        assert ol.sw_linear(a, b) == (s, *end)
        assert ol.last_row_max(a, b) == lrm
        assert ol.ref_compat_align(a, b) == rc


dna = st.binary(min_size=0, max_size=48).map(lambda x: bytes(b"ACGTN"[c % 5] for c in x))


@settings(max_examples=300, deadline=None)
@given(dna, dna)
def test_oracle_equals_numpy_twin(a, b):
    assert ol.sw_linear(a, b) == sw_numpy.sw_linear(a, b)
    assert ol.last_row_max(a, b) == sw_numpy.last_row_max(a, b)


@settings(max_examples=200, deadline=None)
@given(dna, dna, dna)
def test_score_properties(a, b, c):
    s, i, j = ol.sw_linear(a, b)
    assert 0 <= s <= 2 * min(len(a), len(b))
    assert ol.sw_linear(a, a)[0] == 2 * len(a)
    assert ol.sw_linear(b, a)[0] == s                                   # score is symmetric
    assert ol.sw_linear(a + c, b)[0] >= s and ol.sw_linear(a, b + c)[0] >= s   # appending never lowers it
    if s > 0:
        assert 0 <= i < len(a) and 0 <= j < len(b) and a[i] == b[j]     # a maximum ends on a match
        assert ol.sw_linear(a[:i + 1], b[:j + 1]) == (s, i, j)          # ... and is already there in the prefix
    else:
        assert (i, j) == (-1, -1)


@settings(max_examples=150, deadline=None)
@given(dna, dna)
def test_global_max_is_max_of_last_row_reductions(a, b):
    """Ties the full-matrix maximum to the reference's last-row reduction (cl:130-134):
    max over row prefixes of last_row_max == global maximum."""
    s = ol.sw_linear(a, b)[0]
    assert s == max([ol.last_row_max(a[:k], b) for k in range(1, len(a) + 1)] + [0])


def _random_batch(rng, n, rl, wl, alphabet=b"ACGT"):
    al = np.frombuffer(alphabet, dtype=np.uint8)
    reads = [al[rng.integers(0, al.size, int(rng.integers(rl[0], rl[1] + 1)))] for _ in range(n)]
    wins = [al[rng.integers(0, al.size, int(rng.integers(wl[0], wl[1] + 1)))] for _ in range(n)]
    from mini_parallel_b200.engine import to_csr
    return to_csr(reads) + to_csr(wins)


@pytest.mark.parametrize("isa", [2, 1])
def test_simd_port_equals_scalar(isa):
    rng = np.random.default_rng(7)
    q, qo, r, ro = _random_batch(rng, 700, (0, 220), (0, 420), b"ACGTN")
    ol.oracle().sw_simd_force_isa(isa)
    try:
        simd = ol.batch(q, qo, r, ro, threads=3, simd=True)
    finally:
        ol.oracle().sw_simd_force_isa(2)
    scalar = ol.batch(q, qo, r, ro, threads=2, simd=False)
    assert np.array_equal(simd, scalar)


def test_ref_compat_closed_form():
    """On the NVIDIA geometry (wgs=1024 => chunk <= wgs) each work-item sees at most one position, so the live
    kernel returns 2 iff some aligned position matches."""
    rng = np.random.default_rng(3)
    for _ in range(200):
        n1, n2 = int(rng.integers(1, 5000)), int(rng.integers(1, 5000))
        a = rng.integers(0, 4, n1).astype(np.uint8) + 65
        b = rng.integers(0, 4, n2).astype(np.uint8) + (65 if rng.random() < 0.7 else 97)
        L = min(n1, n2)
        assert ol.ref_compat_align(a, b, 1024) == (2 if np.any(a[:L] == b[:L]) else 0)
