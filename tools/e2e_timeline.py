#!/usr/bin/env python
"""Per-chunk timeline of one swb_score_batch call (SWB_DEBUG_TIMELINE): where the host pipeline has bubbles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SWB_DEBUG_TIMELINE"] = "1"
import torch
import mini_parallel_b200 as mp

n, rl, wl = 1_000_000, 150, 500
dev = torch.device("cuda", 0)
eng = mp.Engine(0); lib = mp.load_library()
d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
eng.synth_device(0, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr()); eng.sync()
h_q = d_q.cpu().pin_memory(); h_r = d_r.cpu().pin_memory(); h_qo = d_qo.cpu().pin_memory(); h_ro = d_ro.cpu().pin_memory()
h_out = torch.empty(n * 3, dtype=torch.int32).pin_memory()
for mb, ramp in ((32, 0), (32, 2), (64, 2)):
    eng.set_chunking(mb << 20, 16384); eng.set_chunk_ramp(ramp)
    for k in range(3):
        lib.swb_score_batch(eng._h, h_q.data_ptr(), h_qo.data_ptr(), h_r.data_ptr(), h_ro.data_ptr(), n, h_out.data_ptr())
    print(f"=== chunk {mb} MiB ramp {ramp}", file=sys.stderr)
    eng.last_timings()
