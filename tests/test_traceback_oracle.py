"""The oracle's traceback (oracle/sw_oracle.c:sw_traceback, this repository's definition of the alignment behind a score:
SURVEY.md 8f rank 4) against the alignments SURVEY.md 8c describes, an independent pure-Python twin, and the properties
any correct CIGAR has."""
import numpy as np

import oracle_lib as ol


def _py_traceback(a, b, ei, ej):
    """Independent twin: full matrix in Python lists, predecessor order diagonal, up, left."""
    n, m = len(a), len(b)
    H = [[0] * (m + 1) for _ in range(n + 1)]
    for i in range(n):
        for j in range(m):
            s = 2 if a[i] == b[j] else -1
            H[i + 1][j + 1] = max(0, H[i][j] + s, H[i][j + 1] - 2, H[i + 1][j] - 2)
    ops, i, j, si, sj = [], ei, ej, -1, -1
    while i >= 0 and j >= 0 and H[i + 1][j + 1] > 0:
        si, sj = i, j
        s = 2 if a[i] == b[j] else -1
        if H[i + 1][j + 1] == H[i][j] + s:
            op = "=" if a[i] == b[j] else "X"; i -= 1; j -= 1
        elif H[i + 1][j + 1] == H[i][j + 1] - 2:
            op = "I"; i -= 1
        else:
            op = "D"; j -= 1
        if ops and ops[-1][1] == op:
            ops[-1][0] += 1
        else:
            ops.append([1, op])
    return si, sj, [(l, o) for l, o in reversed(ops)]


def replay(a, b, si, sj, cigar):
    """Score of the alignment a CIGAR describes and the cell it ends in; checks '=' / 'X' against the bytes."""
    i, j, score = si, sj, 0
    for length, op in cigar:
        for _ in range(length):
            if op in "=X":
                assert (a[i] == b[j]) == (op == "="), (i, j, op)
                score += 2 if op == "=" else -1; i += 1; j += 1
            elif op == "I":
                score -= 2; i += 1
            else:
                assert op == "D"; score -= 2; j += 1
    return score, i - 1, j - 1


def test_known_alignments():
    # ACGTACGT / ACGACGT = 7 matches and one gap; TGTTACGG / GGTTGACTA = GTT-AC over GTTGAC
    s, ei, ej = ol.sw_linear("ACGTACGT", "ACGACGT")
    assert ol.traceback("ACGTACGT", "ACGACGT", ei, ej) == (0, 0, [(3, "="), (1, "I"), (4, "=")])
    s, ei, ej = ol.sw_linear("TGTTACGG", "GGTTGACTA")
    assert ol.traceback("TGTTACGG", "GGTTGACTA", ei, ej) == (1, 1, [(3, "="), (1, "D"), (2, "=")])
    s, ei, ej = ol.sw_linear("ATCGT", "ATTGG")                       # README.md:7-11 read as a pair
    assert (s, ei, ej) == (5, 3, 3) and ol.traceback("ATCGT", "ATTGG", ei, ej) == (0, 0, [(2, "="), (1, "X"), (1, "=")])
    assert ol.traceback("AAAA", "TTTT", -1, -1) == (-1, -1, [])      # score 0: nothing aligned
    assert ol.traceback("NNNN", "NNNN", 3, 3) == (0, 0, [(4, "=")])  # raw byte equality (cl:114)


def test_matches_the_python_twin_and_replays_to_the_score():
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    for t in range(400):
        n, m = int(rng.integers(1, 40)), int(rng.integers(1, 60))
        k = int(rng.integers(1, 5))                                   # small alphabets: many ties
        b = alpha[rng.integers(0, k, m)].tobytes()
        if t % 2:
            a = alpha[rng.integers(0, k, n)].tobytes()
        else:                                                         # a mutated piece of b: real alignments with indels
            o = int(rng.integers(0, m)); piece = bytearray(b[o:o + n] or b[:1])
            for _ in range(int(rng.integers(0, 4))):
                p = int(rng.integers(0, len(piece)))
                c = int(rng.integers(0, 3))
                if c == 0: piece[p] = int(alpha[rng.integers(0, 4)])
                elif c == 1: del piece[p]
                else: piece.insert(p, int(alpha[rng.integers(0, 4)]))
                if not piece: piece = bytearray(b"A")
            a = bytes(piece)
        s, ei, ej = ol.sw_linear(a, b)
        got = ol.traceback(a, b, ei, ej)
        assert got == _py_traceback(a, b, ei, ej), (a, b)
        if s:
            si, sj, cigar = got
            assert replay(a, b, si, sj, cigar) == (s, ei, ej)
            rows = sum(l for l, o in cigar if o in "=XI"); cols = sum(l for l, o in cigar if o in "=XD")
            assert rows == ei - si + 1 and cols == ej - sj + 1 and cols <= 2 * rows
            assert cigar[0][1] == "=" and cigar[-1][1] == "="           # a local alignment starts and ends on a match


def _diagonal_rule(a, b, score, ei, ej):
    """What traceback_diag_kernel does on the GPU: walking back along the diagonal from the end cell, the first k cells whose
    substitution scores add up to the score ARE the alignment (csrc/swb_traceback.cu has the argument); None when a partial
    sum passes the score or the diagonal runs out (a gap: the matrix has to decide)."""
    s, runs = 0, []
    for t in range(min(ei, ej) + 1):
        eq = a[ei - t] == b[ej - t]
        s += 2 if eq else -1
        if runs and runs[-1][1] == ("=" if eq else "X"):
            runs[-1][0] += 1
        else:
            runs.append([1, "=" if eq else "X"])
        if s == score:
            return ei - t, ej - t, [(l, o) for l, o in reversed(runs)]
        if s > score:
            return None
    return None


def test_gapless_alignments_follow_the_diagonal_rule():
    rng = np.random.default_rng(9)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    handled = total = 0
    for t in range(6000):
        n, m, k = int(rng.integers(1, 40)), int(rng.integers(1, 60)), int(rng.integers(1, 5))
        b = alpha[rng.integers(0, k, m)].tobytes()
        if t % 3 == 0:
            a = alpha[rng.integers(0, k, n)].tobytes()
        else:
            o = int(rng.integers(0, m)); piece = bytearray(b[o:o + n] or b[:1])
            for _ in range(int(rng.integers(0, 3))):
                p = int(rng.integers(0, len(piece))); c = int(rng.integers(0, 3))
                if c == 0: piece[p] = int(alpha[rng.integers(0, 4)])
                elif c == 1 and len(piece) > 1: del piece[p]
                else: piece.insert(p, int(alpha[rng.integers(0, 4)]))
            a = bytes(piece)
        s, ei, ej = ol.sw_linear(a, b)
        if s <= 0:
            continue
        total += 1
        got = _diagonal_rule(a, b, s, ei, ej)
        if got is not None:
            handled += 1
            assert got == ol.traceback(a, b, ei, ej), (a, b)
    assert handled > total // 2                                       # most alignments of such pairs have no gap
