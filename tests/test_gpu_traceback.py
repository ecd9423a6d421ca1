"""swb_traceback_batch (start cell + CIGAR behind a score) against the oracle's traceback, pair by pair."""
import numpy as np
import pytest

import oracle_lib as ol
from mini_parallel_b200.engine import SwbError, to_csr
from test_traceback_oracle import replay

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate(rng, piece, n_edits):
    piece = bytearray(piece)
    for _ in range(n_edits):
        if not piece:
            piece = bytearray(b"A")
        p = int(rng.integers(0, len(piece))); c = int(rng.integers(0, 3))
        if c == 0: piece[p] = int(ACGT[rng.integers(0, 4)])
        elif c == 1: del piece[p]
        else: piece.insert(p, int(ACGT[rng.integers(0, 4)]))
    return bytes(piece) or b"A"


def _check(engine, reads, wins):
    q, qo = to_csr(reads); r, ro = to_csr(wins)
    res = engine.score_batch_csr(q, qo, r, ro)
    al, ops = engine.traceback_batch(q, qo, r, ro, res)
    assert int(al["cigar_len"].sum()) == ops.size
    for k, (a, b) in enumerate(zip(reads, wins)):
        exp = ol.traceback(a, b, int(res[k]["end_i"]), int(res[k]["end_j"]))
        got = (int(al[k]["start_i"]), int(al[k]["start_j"]), engine.cigar_of(al[k], ops))
        assert al[k]["status"] == 0 and got == exp, (k, a, b, got, exp)
    return res, al, ops


def test_random_pairs_match_the_oracle(engine):
    rng = np.random.default_rng(41)
    reads, wins = [], []
    for k in range(1500):
        m = int(rng.choice([1, 7, 33, 100, 500, 501, 900])); w = ACGT[rng.integers(0, int(rng.integers(1, 5)), m)].tobytes()
        kind = k % 4
        if kind == 0:                                                   # related read with substitutions and indels
            n = int(rng.integers(1, 161)); o = int(rng.integers(0, m))
            rd = _mutate(rng, w[o:o + n] or w[:1], int(rng.integers(0, 6)))
        elif kind == 1:                                                 # unrelated
            rd = ACGT[rng.integers(0, 4, int(rng.integers(1, 200)))].tobytes()
        elif kind == 2:                                                 # homopolymers and short repeats: ties everywhere
            rd = (b"A" * int(rng.integers(1, 50)) + b"AC" * int(rng.integers(0, 20)))[:160]
        else:                                                           # any bytes: N, lower case (raw byte equality, cl:114)
            rd = bytes(np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, int(rng.integers(1, 120)))])
            w = bytes(np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, m)])
        reads.append(rd); wins.append(w)
    reads[3] = b""; wins[5] = b""                                       # empty sides: score 0
    res, al, ops = _check(engine, reads, wins)
    zero = res["score"] == 0
    assert zero.any() and (al["start_i"][zero] == -1).all() and (al["cigar_len"][zero] == 0).all()


def test_long_pairs_and_replay(engine):
    """Pairs far beyond the short-read limits: the rectangle rule (at most twice as many columns as rows) and the scratch
    sizing; every CIGAR replays to the score and ends in the end cell."""
    rng = np.random.default_rng(42)
    reads, wins = [], []
    for n, m in ((3000, 5000), (2500, 2500), (700, 9000), (4000, 300)):
        w = ACGT[rng.integers(0, 4, m)].tobytes()
        o = int(rng.integers(0, max(1, m - n)))
        reads.append(_mutate(rng, w[o:o + n], 40)); wins.append(w)
    res, al, ops = _check(engine, reads, wins)
    for k in range(len(reads)):
        cigar = engine.cigar_of(al[k], ops)
        assert replay(reads[k], wins[k], int(al[k]["start_i"]), int(al[k]["start_j"]), cigar) == tuple(int(x) for x in res[k])


def test_config_shape_batch(engine):
    """BASELINE.json configs[1] shape (150 x 500, related reads): 20 000 pairs, a sample checked against the oracle, all of
    them against the properties a CIGAR must have."""
    from mini_parallel_b200 import synth
    q, qo, r, ro = synth.make_pairs(0, 20_000, 150, 500, 0)
    res = engine.score_batch_csr(q, qo, r, ro)
    al, ops = engine.traceback_batch(q, qo, r, ro, res)
    assert (al["status"] == 0).all() and int(al["cigar_len"].sum()) == ops.size
    lens = (ops >> 4).astype(np.int64); kinds = ops & 15
    rows_per_op = np.where(np.isin(kinds, (7, 8, 1)), lens, 0); cols_per_op = np.where(np.isin(kinds, (7, 8, 2)), lens, 0)
    order = np.argsort(al["cigar_off"])                                 # slices are in no particular order
    owner = np.empty(ops.size, dtype=np.int64)
    owner[:] = np.repeat(order, al["cigar_len"][order].astype(np.int64))
    rows = np.bincount(owner, weights=rows_per_op, minlength=al.size).astype(np.int64)
    cols = np.bincount(owner, weights=cols_per_op, minlength=al.size).astype(np.int64)
    assert (rows == res["end_i"] - al["start_i"] + 1).all() and (cols == res["end_j"] - al["start_j"] + 1).all()
    for k in range(0, 20_000, 97):
        a = q[int(qo[k]):int(qo[k + 1])].tobytes(); b = r[int(ro[k]):int(ro[k + 1])].tobytes()
        assert (int(al[k]["start_i"]), int(al[k]["start_j"]), engine.cigar_of(al[k], ops)) == ol.traceback(a, b, int(res[k]["end_i"]), int(res[k]["end_j"]))


def test_bad_results_and_small_buffers(engine):
    reads, wins = [b"ACGTACGT", b"TTTT", b"ACGT"], [b"ACGACGT", b"TTTTT", b"ACGT"]
    q, qo = to_csr(reads); r, ro = to_csr(wins)
    res = engine.score_batch_csr(q, qo, r, ro)
    bad = res.copy()
    bad[0]["end_j"] = 2                                                # a cell whose value is not the score
    bad[1]["end_i"] = 99                                               # outside the pair
    al, ops = engine.traceback_batch(q, qo, r, ro, bad)
    assert list(al["status"]) == [1, 1, 0] and engine.cigar_of(al[2], ops) == [(4, "=")]
    with pytest.raises(SwbError, match="operations"):                  # room for one operation, the batch needs five
        engine.traceback_batch(q, qo, r, ro, res, cigar_cap=1)
    al, ops = engine.traceback_batch(q, qo, r, ro, res)                # the engine is usable afterwards
    assert engine.cigar_of(al[0], ops) == [(3, "="), (1, "I"), (4, "=")]
