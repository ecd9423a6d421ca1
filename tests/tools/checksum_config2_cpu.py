#!/usr/bin/env python
"""BASELINE.json configs[2] at FULL size on the CPU checker: the 64-bit checksum bench.py's `config2_strong` leg reports for the
GPU's results over all 100 M pairs (a wrapping sum of hash(pair index, score, end_i, end_j), the same at every N), recomputed
from the CPU SIMD port's results over the same 100 M pairs (SURVEY.md 8d generator, host twin mini_parallel_b200/synth.py).
Equal checksums = every one of the 100 M GPU results equals the checker's (up to a 2^-64 collision), not only the 1 M pairs
per rank the bench leg compares directly.  No GPU; ~2 core-hours.

    python tests/tools/checksum_config2_cpu.py [--pairs 100000000] [--expect 5ec5834cd0ec6b2a] [--procs N]"""
import argparse
import json
import multiprocessing as mpc
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
RL, WL, SLICE = 150, 500, 100_000


def _i64(x):
    return np.int64(x - (1 << 64) if x >= (1 << 63) else x)


C1, C2, C3, C4, C5 = (_i64(0x9E3779B97F4A7C15), _i64(0xBF58476D1CE4E5B9), _i64(0x94D049BB133111EB), _i64(0xD6E8FEB86659FD93), _i64(0xFF51AFD7ED558CCD))


def checksum64(res, first_index):
    """bench.py's checksum64 in numpy: int64 arithmetic that wraps, an arithmetic shift, a wrapping sum."""
    with np.errstate(over="ignore"):
        idx = np.arange(first_index, first_index + res.size, dtype=np.int64)
        h = idx * C1 + res["score"].astype(np.int64) * C2 + res["end_i"].astype(np.int64) * C3 + res["end_j"].astype(np.int64) * C4
        h = (h ^ (h >> np.int64(29))) * C5
        return int(h.view(np.uint64).sum(dtype=np.uint64))


def one_slice(job):
    first, n, dist, rl, wl = job
    import oracle_lib as ol
    from mini_parallel_b200 import synth
    q, qo, r, ro = synth.make_pairs(first, n, rl, wl, dist)
    res = ol.batch(q, np.asarray(qo, dtype=np.uint64), r, np.asarray(ro, dtype=np.uint64), threads=1, simd=True)
    plain = int(res["score"].sum(dtype=np.int64)) + int(res["end_i"].sum(dtype=np.int64)) + int(res["end_j"].sum(dtype=np.int64))
    return checksum64(res, first), int(res["score"].sum(dtype=np.int64)), n, plain


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100_000_000)
    ap.add_argument("--dist", type=int, default=0)
    ap.add_argument("--read", type=int, default=RL, help="other shapes of the same generator (tools/probe_rows128.py records the plain sum)")
    ap.add_argument("--window", type=int, default=WL)
    ap.add_argument("--expect", default="", help="the GPU's checksum64 (hex) from the bench line")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    jobs = [(a, min(SLICE, args.pairs - a), args.dist, args.read, args.window) for a in range(0, args.pairs, SLICE)]
    t0, csum, ssum, done, plain = time.time(), 0, 0, 0, 0
    with mpc.get_context("fork").Pool(args.procs) as pool:
        for k, (c, s_, n, pl) in enumerate(pool.imap_unordered(one_slice, jobs, chunksize=1)):
            csum = (csum + c) & ((1 << 64) - 1); ssum += s_; done += n; plain += pl
            if (k + 1) % 100 == 0:
                print(f"  {done} pairs, {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
    out = {"workload": f"{'BASELINE.json configs[2]: ' if (args.read, args.window) == (RL, WL) else ''}{args.pairs} synthetic {args.read} bp reads vs {args.window} bp windows (SURVEY.md 8d generator, distribution {args.dist})",
           "checker": "oracle/sw_simd.c (CPU SIMD port, bit-exact with the scalar restatement)", "pairs": done, "checksum64": f"{csum:016x}",
           "sum_of_score_end_i_end_j": plain, "mean_score": round(ssum / max(done, 1), 3), "seconds": round(time.time() - t0, 1), "procs": args.procs}
    if args.expect:
        out["gpu_checksum64"] = args.expect.lower()
        out["equal"] = out["checksum64"] == args.expect.lower()
    print(json.dumps(out))
    return 0 if (not args.expect or out["equal"]) else 1


if __name__ == "__main__":
    sys.exit(main())
