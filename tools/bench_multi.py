#!/usr/bin/env python
"""One process, N GPUs behind swb_create_multi: a batch of N x 1 M pairs from ONE set of pinned host arrays is split
contiguously over the devices (north_star: "each chunk is partitioned across the 8 GPUs of one box with per-GPU
streams").  Reads + window starts travel, the reference is resident on every device.  Prints one JSON line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import mini_parallel_b200 as mp
from mini_parallel_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0, help="0 = every visible device")
    ap.add_argument("--pairs-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ref-bases", type=int, default=16_000_000)
    args = ap.parse_args()
    ng = args.gpus or mp.device_count()
    n, rl, wl = ng * args.pairs_per_gpu, 150, 500
    lib = mp.load_library()
    ref = synth.synth_reference(args.ref_bases)
    h_q = torch.empty(n * rl, dtype=torch.uint8).pin_memory(); h_ws = torch.empty(n, dtype=torch.int64).pin_memory()
    h_qo = (torch.arange(n + 1, dtype=torch.int64) * rl).pin_memory(); h_wl = torch.full((n,), wl, dtype=torch.int32).pin_memory()
    h_out = torch.empty(n * 3, dtype=torch.int32).pin_memory()
    # the workload, generated on device 0 in slices of 1 M pairs
    eng = mp.Engine(0)
    dev = torch.device("cuda", 0)
    sl = args.pairs_per_gpu
    d_ref = torch.from_numpy(ref).to(dev)
    d_q = torch.empty(sl * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(sl * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(sl + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(sl + 1, dtype=torch.int64, device=dev); d_ws = torch.empty(sl, dtype=torch.int64, device=dev)
    d_out = torch.empty(sl * 3, dtype=torch.int32, device=dev)
    first = None
    for k in range(ng):
        eng.synth_device_ref(d_ref.data_ptr(), args.ref_bases, k * sl, sl, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr(), d_ws.data_ptr())
        eng.sync()
        h_q[k * sl * rl:(k + 1) * sl * rl].copy_(d_q); h_ws[k * sl:(k + 1) * sl].copy_(d_ws)
        if k == 0:
            eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), sl * rl, d_r.data_ptr(), d_ro.data_ptr(), sl * wl, sl, rl, wl, d_out.data_ptr())
            eng.sync()
            first = d_out.cpu().clone()
    torch.cuda.synchronize()
    del d_q, d_r, d_qo, d_ro, d_ws, d_out, d_ref
    eng.close()
    t0 = time.perf_counter()
    me = mp.MultiEngine(list(range(ng)))
    create_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    me.set_reference(ref)
    setref_s = time.perf_counter() - t0

    def call():
        rc = lib.swb_multi_score_batch_vs_reference(me._h, h_q.data_ptr(), h_qo.data_ptr(), n, h_ws.data_ptr(), h_wl.data_ptr(), h_out.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.swb_last_error().decode())
    for _ in range(2):
        call()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        call()
    dt = (time.perf_counter() - t0) / args.reps
    same = bool(torch.equal(h_out[: sl * 3], first))
    print(json.dumps({"what": "one process, swb_create_multi: one host batch split contiguously over the devices (reads + window starts from pinned host memory, "
                              "reference resident on every device)", "n_gpus": ng, "pairs": n, "ms_per_call": round(dt * 1e3, 3),
                      "gcups": round(n * rl * wl / dt / 1e9, 1), "reads_per_s": round(n / dt, 1), "h2d_bytes_per_call": n * (rl + 8), "d2h_bytes_per_call": n * 12,
                      "first_slice_equals_single_device_results": same, "create_contexts_s": round(create_s, 3), "set_reference_s": round(setref_s, 3),
                      "host_cpus": os.cpu_count()}), flush=True)
    me.close()


if __name__ == "__main__":
    main()
