#!/usr/bin/env python
"""BASELINE.json configs[3]: long-read pairs (10 kb x 10 kb) through the device-resident API.
Prints one JSON line: GCUPS of the whole step and of the long/generic kernel alone, plus an oracle
spot-check of a few pairs (the full-size oracle run is a parity test, not part of the timing)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4736)
    ap.add_argument("--len", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dist", type=int, default=0)
    ap.add_argument("--check", type=int, default=4, help="pairs compared with the CPU oracle")
    args = ap.parse_args()
    import torch
    import mini_parallel_b200 as mp
    eng = mp.Engine(0)
    dev = torch.device("cuda", 0)
    n, L = args.pairs, args.len
    d_q = torch.empty(n * L, dtype=torch.uint8, device=dev)
    d_r = torch.empty(n * L, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * 3, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    eng.synth_device(0, n, L, L, args.dist, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
    eng.sync()
    ms_all, ms_long = [], []
    for s in range(args.steps + 1):
        eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * L, d_r.data_ptr(), d_ro.data_ptr(), n * L, n, L, L, d_out.data_ptr())
        t = eng.last_timings()
        if s:
            ms_all.append(t["device_ms"]); ms_long.append(t["generic_ms"] + t["short_ms"])
    cells = float(n) * L * L
    out = d_out.cpu().numpy().reshape(n, 3)
    ok = None
    if args.check:
        import oracle_lib as ol
        q = d_q[: args.check * L].cpu().numpy(); r = d_r[: args.check * L].cpu().numpy()
        ok = all(tuple(out[k]) == ol.sw_linear(q[k * L:(k + 1) * L].tobytes(), r[k * L:(k + 1) * L].tobytes()) for k in range(args.check))
    print(json.dumps({"workload": f"{n} pairs {L}x{L}, dist {args.dist}", "ms_per_step": round(float(np.mean(ms_all)), 3),
                      "gcups": round(cells / (np.mean(ms_all) * 1e-3) / 1e9, 1),
                      "kernel_ms": round(float(np.mean(ms_long)), 3), "kernel_gcups": round(cells / (np.mean(ms_long) * 1e-3) / 1e9, 1),
                      "routing": eng.last_routing(), "mean_score": float(out[:, 0].mean()), "oracle_check": ok}))


if __name__ == "__main__":
    main()
