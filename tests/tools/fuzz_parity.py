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
    rounds, pairs_total, routes = 0, 0, {"short": 0, "generic": 0, "long": 0}
    while time.time() < t_end:
        shape = rng.integers(0, 6)
        if shape == 0:   n, rl, wl = int(rng.integers(1, 30000)), (1, 160), (1, 900)
        elif shape == 1: n, rl, wl = int(rng.integers(1, 30000)), (150, 150), (500, 500)
        elif shape == 2: n, rl, wl = int(rng.integers(1, 3000)), (1, 200), (1, 4200)
        elif shape == 3: n, rl, wl = int(rng.integers(1, 400)), (100, 1500), (1, 2500)
        elif shape == 4: n, rl, wl = int(rng.integers(1, 40)), (300, 6000), (300, 6000)
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
        variant = int(rng.choice([4, 4, 4, 5, 6, 1]))
        chunk = (int(rng.choice([1 << 14, 1 << 17, 1 << 20, 32 << 20])), int(rng.choice([1, 256, 16384])))
        eng.set_short_variant(variant); eng.set_chunking(*chunk)
        got = eng.score_batch_csr(q, qo, r, ro)
        exp = ol.batch(q, qo, r, ro, threads=os.cpu_count() or 8, simd=True)
        bad = np.nonzero(got != exp)[0]
        if bad.size:
            k = int(bad[0])
            print(f"MISMATCH seed={args.seed} round={rounds} shape={shape} n={n} variant={variant} chunk={chunk} pair={k} "
                  f"n1={n1[k]} n2={n2[k]} got={got[k]} exp={exp[k]} ({bad.size} pairs differ)")
            return 1
        rt = eng.last_routing()
        for kk in routes:
            routes[kk] += rt[kk]
        rounds += 1; pairs_total += n
    print(f"fuzz ok: {rounds} batches, {pairs_total} pairs, routing {routes}, seed {args.seed}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
