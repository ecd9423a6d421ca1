#!/usr/bin/env python
"""swb_traceback_batch on BASELINE.json configs[1]-shaped pairs (150 x 500, related reads): time per batch through the
host API (H2D of the pairs and their results, kernel, D2H of alignments + CIGARs) and agreement of a sample with the oracle."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

import mini_parallel_b200 as mp
import oracle_lib as ol
from mini_parallel_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
eng = mp.Engine(0)
q, qo, r, ro = synth.make_pairs(0, n, 150, 500, 0)
res = eng.score_batch_csr(q, qo, r, ro)
al, ops = eng.traceback_batch(q, qo, r, ro, res, cigar_cap=16 * n)
t0 = time.perf_counter()
for _ in range(3):
    al, ops = eng.traceback_batch(q, qo, r, ro, res, cigar_cap=16 * n)
dt = (time.perf_counter() - t0) / 3
kernel_ms = eng.last_timings()["device_ms"]
ok = True
for k in range(0, n, max(1, n // 200)):
    a = q[int(qo[k]):int(qo[k + 1])].tobytes(); b = r[int(ro[k]):int(ro[k + 1])].tobytes()
    ok = ok and (int(al[k]["start_i"]), int(al[k]["start_j"]), eng.cigar_of(al[k], ops)) == ol.traceback(a, b, int(res[k]["end_i"]), int(res[k]["end_j"]))
cells = float(((res["end_i"].astype(np.int64) + 1) * np.minimum(res["end_j"].astype(np.int64) + 1, 2 * (res["end_i"].astype(np.int64) + 1))).sum())
print(json.dumps({"pairs": n, "ms_per_batch_host_api": round(dt * 1e3, 2), "kernel_ms": round(kernel_ms, 3),
                  "alignments_per_s_kernel": round(n / (kernel_ms * 1e-3), 1), "rectangle_cells": cells,
                  "gcups_recomputed_kernel": round(cells / (kernel_ms * 1e-3) / 1e9, 1), "operations": int(ops.size), "ops_per_alignment": round(ops.size / n, 2),
                  "sample_equals_oracle": bool(ok)}))
