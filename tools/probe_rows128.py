#!/usr/bin/env python
"""Reads of 100 / 125 / 128 bp against 500 bp windows: the 128-row instantiation of the stream kernel vs the 160-row one (SWB_ROWS128=0)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import numpy as np, torch
    import mini_parallel_b200 as mp
    eng = mp.Engine(0); dev = torch.device("cuda", 0)
    n, wl = 1_000_000, 500
    out = {}
    for rl in (100, 125, 128):
        d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
        d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
        d_out = torch.empty(n * 3, dtype=torch.int32, device=dev)
        eng.synth_device(0, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr()); eng.sync()
        ms = []
        for s in range(5):
            eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * rl, d_r.data_ptr(), d_ro.data_ptr(), n * wl, n, rl, wl, d_out.data_ptr()); eng.sync()
            ms.append(eng.last_timings()["short_ms"])
        out[f"{rl} x {wl}"] = {"short_ms": round(min(ms[1:]), 3), "gcups": round(n * rl * wl / min(ms[1:]) / 1e6, 1), "checksum": int(d_out.to(torch.int64).sum().item())}
    print(json.dumps(out))
else:
    res = {}
    for rows, env in (("128 rows", {}), ("160 rows", {"SWB_ROWS128": "0"})):
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
        res[rows] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-500:]
    print(json.dumps(res, indent=1))
