#!/usr/bin/env python
"""Does the device-resident step slow down with batch size or with time under load?  (round 2: the 100 M strong-scaling
leg ran at 15 ms per million pairs, the 1 M-pair step at 9.65.)  Prints per-configuration stage times and, for a long
back-to-back run, the step time and SM clock over time."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mini_parallel_b200 as mp

eng = mp.Engine(0)
dev = torch.device("cuda", 0)
rl, wl = 150, 500
out = {}
for n in (1_000_000, 2_000_000, 5_000_000, 10_000_000):
    d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * 3, dtype=torch.int32, device=dev)
    eng.synth_device(0, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr()); eng.sync()
    rows = []
    for s in range(4):
        eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * rl, d_r.data_ptr(), d_ro.data_ptr(), n * wl, n, rl, wl, d_out.data_ptr())
        t = eng.last_timings()
        rows.append({k: round(t[k] / (n / 1e6), 3) for k in ("pack_classify_ms", "short_ms", "device_ms")})
        time.sleep(0.5)
    out[f"n={n} (ms per million pairs, 4 launches 0.5 s apart)"] = rows
    if n == 1_000_000:
        keep = (d_q, d_r, d_qo, d_ro, d_out)
    else:
        del d_q, d_r, d_qo, d_ro, d_out
d_q, d_r, d_qo, d_ro, d_out = keep
n = 1_000_000
clk = []
def sample():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    t0 = time.time()
    for line in p.stdout:
        clk.append((round(time.time() - t0, 2), line.strip()))
        if time.time() - t0 > 6: break
    p.terminate()
th = threading.Thread(target=sample, daemon=True); th.start()
time.sleep(0.5)
steps = []
t0 = time.time()
for s in range(300):
    eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * rl, d_r.data_ptr(), d_ro.data_ptr(), n * wl, n, rl, wl, d_out.data_ptr())
    t = eng.last_timings()
    steps.append((round(time.time() - t0, 3), round(t["short_ms"], 3)))
th.join(timeout=8)
out["300 steps of 1M back to back: (t_s, short_ms) every 20th"] = steps[::20]
out["nvidia-smi during that (t_s, 'sm MHz, W, sw_power_cap, hw_slowdown, sw_thermal')"] = clk[::3]
print(json.dumps(out, indent=1))
