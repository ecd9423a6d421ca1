#!/usr/bin/env python
"""BASELINE.json configs[0] -- "--test-wgs shape: 10k synthetic 150bp reads vs 1 kb reference windows, reference CPU SIMD path
(runs without a GPU)" -- timed on the host cores: the CPU SIMD port (oracle/sw_simd.c, the stand-in for the CPU SIMD path the
reference does not have) over the SURVEY.md 8d workload of that shape, every result compared with the scalar
restatement (oracle/sw_oracle.c).  No GPU.  Prints one JSON line.

    python tests/tools/bench_config0_cpu.py [--pairs 10000] [--window 1000] [--threads N] [--reps 9]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=10_000)
    ap.add_argument("--read", type=int, default=150)
    ap.add_argument("--window", type=int, default=1000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--reps", type=int, default=9)
    args = ap.parse_args()
    import oracle_lib as ol
    from mini_parallel_b200 import synth
    out = {"workload": f"BASELINE.json configs[0]: {args.pairs} synthetic {args.read} bp reads vs {args.window} bp windows (SURVEY.md 8d generator)",
           "host_cores": os.cpu_count(), "threads": args.threads, "isa": ol.simd_isa(), "rows": []}
    for dist, tag in ((0, "related reads (1 % subst, 0.1 % ins, 0.1 % del)"), (1, "unrelated reads")):
        q, qo, r, ro = synth.make_pairs(0, args.pairs, args.read, args.window, dist)
        qo = qo.astype(np.uint64); ro = ro.astype(np.uint64)
        cells = float(args.pairs) * args.read * args.window
        exp = ol.batch(q, qo, r, ro, threads=args.threads, simd=False)            # scalar restatement, all pairs
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter(); got = ol.batch(q, qo, r, ro, threads=args.threads, simd=True); ts.append(time.perf_counter() - t0)
        same = all(np.array_equal(got[k], exp[k]) for k in ("score", "end_i", "end_j"))
        t0 = time.perf_counter(); ol.batch(q[: 500 * args.read], qo[:501], r[: 500 * args.window], ro[:501], threads=1, simd=False); t1 = time.perf_counter() - t0
        best, med = min(ts), sorted(ts)[len(ts) // 2]
        out["rows"].append({"distribution": tag, "ms_best": round(best * 1e3, 2), "ms_median": round(med * 1e3, 2), "gcups_best": round(cells / best / 1e9, 1),
                            "gcups_median": round(cells / med / 1e9, 1), "reads_per_s_best": round(args.pairs / best, 0),
                            "scalar_oracle_gcups_1_core": round(500.0 * args.read * args.window / t1 / 1e9, 3),
                            "equals_scalar_oracle_on_all_pairs": bool(same), "mean_score": round(float(got["score"].mean()), 2)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
