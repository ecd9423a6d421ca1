#!/usr/bin/env python
"""Per-chunk timeline of swb_score_batch_ranges vs swb_score_batch_vs_reference on the bench workload (SWB_DEBUG_TIMELINE)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SWB_DEBUG_TIMELINE"] = "1"
import numpy as np, torch
import mini_parallel_b200 as mp
from mini_parallel_b200 import synth
eng = mp.Engine(0); lib = mp.load_library()
n, rl, wl, R = 1_000_000, 150, 500, 16_000_000
dev = torch.device("cuda", 0)
h_ref = torch.from_numpy(synth.synth_reference(R)).pin_memory(); d_ref = h_ref.to(dev)
d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev); d_ws = torch.empty(n, dtype=torch.int64, device=dev)
eng.synth_device_ref(d_ref.data_ptr(), R, 0, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr(), d_ws.data_ptr()); eng.sync()
h_q = d_q.cpu().pin_memory(); h_qo = d_qo.cpu().pin_memory(); h_ws = d_ws.cpu().pin_memory(); h_wl = torch.full((n,), wl, dtype=torch.int32).pin_memory()
h_out = torch.empty(n * 3, dtype=torch.int32).pin_memory()
def ranges(): assert lib.swb_score_batch_ranges(eng._h, h_q.data_ptr(), h_qo.data_ptr(), n, h_ref.data_ptr(), R, h_ws.data_ptr(), h_wl.data_ptr(), h_out.data_ptr()) == 0
def resident(): assert lib.swb_score_batch_vs_reference(eng._h, h_q.data_ptr(), h_qo.data_ptr(), n, h_ws.data_ptr(), h_wl.data_ptr(), h_out.data_ptr()) == 0
lib.swb_set_reference(eng._h, h_ref.data_ptr(), R)
for name, fn in (("ranges", ranges), ("resident", resident), ("ranges", ranges), ("resident", resident)):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = (time.perf_counter() - t0) / 5
    print(f"== {name}: {dt * 1e3:.3f} ms per call (wall)", file=sys.stderr)
    fn(); eng.last_timings()
