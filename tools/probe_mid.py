#!/usr/bin/env python
"""Reads of 250 / 300 bp against 500 / 1000 bp windows: the 320-row int16x2 instantiation vs the 32-bit long-pair kernel."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mini_parallel_b200 as mp
eng = mp.Engine(0); dev = torch.device("cuda", 0)
out = {}
for rl, wl in ((250, 500), (256, 500), (200, 500), (300, 1000), (161, 500), (320, 500)):
    n = 500_000
    d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_o = torch.empty(n * 3, dtype=torch.int32, device=dev)
    eng.synth_device(0, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr()); eng.sync()
    row = {}
    keep = None
    for mid in (1, 0):
        eng.set_mid_path(mid)
        ms = []
        for s in range(4):
            eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * rl, d_r.data_ptr(), d_ro.data_ptr(), n * wl, n, rl, wl, d_o.data_ptr())
            t = eng.last_timings()
            if s: ms.append(t["device_ms"])
        m = sum(ms) / len(ms)
        row["int16x2 320-row stream kernel" if mid else "32-bit long-pair kernel"] = {"ms": round(m, 3), "gcups": round(n * rl * wl / m / 1e6, 1), "routing": eng.last_routing_ex()}
        cur = d_o.clone()
        if keep is not None: row["identical_results"] = bool(torch.equal(keep, cur))
        keep = cur
    eng.set_mid_path(1)
    out[f"{n} pairs {rl} x {wl}"] = row
print(json.dumps(out, indent=1))
