#!/usr/bin/env python
"""Time-bounded randomised parity run: random batch shapes, alphabets, chunkings and kernel variants through the C ABI,
every result compared with the CPU oracle (SIMD port, itself pinned to the scalar oracle by the CPU tests).
Exit code 1 and a reproducer line on the first difference."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import mini_parallel_b200 as mp
    from mini_parallel_b200.engine import to_csr
    import oracle_lib as ol
    eng = mp.Engine(0)
    rng = np.random.default_rng(args.seed)
    alphabets = [b"ACGT"] * 6 + [b"A", b"AC", b"ACGTN", b"ACGTacgtN", bytes(range(256))]
    t_end = time.time() + args.seconds
    rounds, pairs_total, routes = 0, 0, {"short": 0, "mid": 0, "long": 0, "bytes": 0, "generic": 0}
    tb_checked = 0
    while time.time() < t_end:
        shape = rng.integers(0, 7)
        if shape == 0:   n, rl, wl = int(rng.integers(1, 30000)), (1, 160), (1, 900)
        elif shape == 1: n, rl, wl = int(rng.integers(1, 30000)), (150, 150), (500, 500)
        elif shape == 2: n, rl, wl = int(rng.integers(1, 3000)), (1, 200), (1, 4200)
        elif shape == 3: n, rl, wl = int(rng.integers(1, 400)), (100, 1500), (1, 2500)
        elif shape == 4: n, rl, wl = int(rng.integers(1, 40)), (300, 6000), (300, 6000)
        elif shape == 6: n, rl, wl = int(rng.integers(1, 6000)), (150, 330), (1, 1300)       # round 2: the 320-row int16x2 path and its edges
        else:            n, rl, wl = int(rng.integers(1, 8000)), (140, 160), (400, 1100)
        al = np.frombuffer(alphabets[int(rng.integers(0, len(alphabets)))], dtype=np.uint8)
        related = bool(rng.integers(0, 2))
        n1 = rng.integers(rl[0], rl[1] + 1, n); n2 = rng.integers(wl[0], wl[1] + 1, n)
        if rng.integers(0, 8) == 0:
            n1[rng.integers(0, n)] = 0
        wins = [al[rng.integers(0, al.size, int(m))] for m in n2]
        reads = []
        for k in range(n):
            a, b = int(n1[k]), int(n2[k])
            if related and b >= a > 0:
                o = int(rng.integers(0, b - a + 1)); r = wins[k][o:o + a].copy()
                m = rng.random(a) < 0.03; r[m] = al[rng.integers(0, al.size, int(m.sum()))]
            else:
                r = al[rng.integers(0, al.size, a)]
            reads.append(r)
        q, qo = to_csr(reads); r, ro = to_csr(wins)
        variant = 9                                        # the product library carries the default stream kernel only
        chunk = (int(rng.choice([1 << 14, 1 << 17, 1 << 20, 32 << 20])), int(rng.choice([1, 256, 16384])))
        eng.set_chunking(*chunk)
        eng.set_mid_path(bool(rng.integers(0, 4)))         # one batch in four: reads of 161..320 bp through the 32-bit kernel instead
        api = int(rng.integers(0, 3))
        if api == 0 or r.size == 0:
            got = eng.score_batch_csr(q, qo, r, ro)
        elif api == 1:                                     # the same windows as ranges of one host buffer
            got = eng.score_batch_ranges(q, qo, r, ro[:-1].copy(), np.diff(ro).astype(np.uint32))
        else:                                              # ... and of the resident reference
            eng.set_reference(r)
            got = eng.score_batch_vs_reference(q, qo, ro[:-1].copy(), np.diff(ro).astype(np.uint32))
        exp = ol.batch(q, qo, r, ro, threads=os.cpu_count() or 8, simd=True)
        bad = np.nonzero(got != exp)[0]
        if bad.size:
            k = int(bad[0])
            print(f"MISMATCH seed={args.seed} round={rounds} shape={shape} n={n} variant={variant} chunk={chunk} pair={k} "
                  f"n1={n1[k]} n2={n2[k]} got={got[k]} exp={exp[k]} ({bad.size} pairs differ)")
            return 1
        rt = eng.last_routing_ex()
        for kk in routes:
            routes[kk] += rt[kk]
        if rounds % 5 == 0:                                # the alignments behind a sample of the scores
            m = min(n, 200)
            al_, ops = eng.traceback_batch(q[:int(qo[m])], qo[:m + 1], r[:int(ro[m])], ro[:m + 1], got[:m])
            for k in range(0, m, 9):
                a1 = q[int(qo[k]):int(qo[k + 1])].tobytes(); b1 = r[int(ro[k]):int(ro[k + 1])].tobytes()
                e = ol.traceback(a1, b1, int(got[k]["end_i"]), int(got[k]["end_j"]))
                g = (int(al_[k]["start_i"]), int(al_[k]["start_j"]), eng.cigar_of(al_[k], ops))
                if g != e:
                    print(f"TRACEBACK MISMATCH seed={args.seed} round={rounds} pair={k} got={g} exp={e}")
                    return 1
                tb_checked += 1
        rounds += 1; pairs_total += n
    print(f"fuzz ok: {rounds} batches, {pairs_total} pairs, routing {routes}, {tb_checked} tracebacks, seed {args.seed}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
