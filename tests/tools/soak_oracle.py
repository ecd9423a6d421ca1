#!/usr/bin/env python
"""Soak of the CHECKER itself: the CPU SIMD port (oracle/sw_simd.c, AVX-512BW and AVX2 paths) -- the checker of the
full-population GPU parity tests and the CPU arm of bench.py -- against the scalar restatement (oracle/sw_oracle.c) on random
batches: the SURVEY.md 8d generator at random shapes (related and unrelated reads), ragged batches over several alphabets
(ACGT, with N, mixed case, two letters = many ties, homopolymers), empty reads and windows, reads up to 320 bp, windows up to
4 096, a few pairs beyond the int16 range.  Score AND end cell must agree on every pair.  No GPU.

    python tests/tools/soak_oracle.py [--seconds 300] [--seed 1]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol                                    # noqa: E402
from mini_parallel_b200 import synth                       # noqa: E402
from mini_parallel_b200.engine import to_csr               # noqa: E402


def ragged(rng, n, rmax, wmax, alphabet):
    al = np.frombuffer(alphabet, dtype=np.uint8)
    reads, wins = [], []
    for _ in range(n):
        w = al[rng.integers(0, al.size, int(rng.integers(0, wmax + 1)))]
        if rng.random() < 0.5 and w.size > 1:                 # a mutated slice of the window
            ln = int(rng.integers(1, min(rmax, w.size) + 1)); o = int(rng.integers(0, w.size - ln + 1))
            r = w[o:o + ln].copy()
            m = rng.random(ln) < 0.05
            r[m] = al[rng.integers(0, al.size, int(m.sum()))]
        else:
            r = al[rng.integers(0, al.size, int(rng.integers(0, rmax + 1)))]
        reads.append(r); wins.append(w)
    return to_csr(reads) + to_csr(wins)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    threads = os.cpu_count() or 1
    alphabets = [b"ACGT", b"ACGTN", b"ACGTacgtNn", b"AC", b"A", b"ACGT"]
    t0, batches, pairs, cells, bad = time.time(), 0, 0, 0.0, 0
    kinds = {}
    while time.time() - t0 < args.seconds:
        kind = int(rng.integers(0, 4))
        if kind == 0:                                          # the workload generator at a random shape
            rl, wl = int(rng.integers(1, 321)), int(rng.integers(1, 2049))
            n = max(1, min(200_000, int(4e8 / (rl * wl))))
            q, qo, r, ro = synth.make_pairs(int(rng.integers(0, 1 << 40)), n, rl, wl, int(rng.integers(0, 2)))
            tag = "generator"
        elif kind == 1:
            q, qo, r, ro = ragged(rng, 3000, 320, int(rng.choice([64, 600, 4096])), alphabets[int(rng.integers(0, 6))]); tag = "ragged"
        elif kind == 2:
            q, qo, r, ro = ragged(rng, 4000, 40, 40, alphabets[int(rng.integers(3, 5))]); tag = "ties"
        else:                                                  # identical long pairs: scores beyond int16 (the 32-bit rescore path)
            al = np.frombuffer(b"ACGT", dtype=np.uint8)
            seqs = [al[rng.integers(0, 4, int(rng.integers(16_000, 20_000)))] for _ in range(3)]
            q, qo, r, ro = to_csr(seqs) + to_csr([s.copy() for s in seqs]); tag = "beyond int16"
        qo = np.asarray(qo, dtype=np.uint64); ro = np.asarray(ro, dtype=np.uint64)
        exp = ol.batch(q, qo, r, ro, threads=threads, simd=False)
        for isa in (2, 1):
            ol.oracle().sw_simd_force_isa(isa)
            try:
                got = ol.batch(q, qo, r, ro, threads=threads, simd=True)
            finally:
                ol.oracle().sw_simd_force_isa(2)
            if not np.array_equal(got, exp):
                k = int(np.flatnonzero(got != exp)[0])
                print(f"DIFFERENCE (isa {isa}, {tag}) at pair {k}: simd {got[k]} scalar {exp[k]}", flush=True)
                bad += 1
        n = qo.size - 1
        batches += 1; pairs += n; cells += float(np.sum(np.diff(qo).astype(np.float64) * np.diff(ro).astype(np.float64)))
        kinds[tag] = kinds.get(tag, 0) + n
    print(f"{batches} batches, {pairs} pairs ({kinds}), {cells:.3e} cells, each through AVX-512BW and AVX2: {bad} differences, {time.time() - t0:.0f} s, {threads} threads")
    print("ok: the SIMD port equals the scalar restatement on every pair" if not bad else "FAILED")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
