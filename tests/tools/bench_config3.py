#!/usr/bin/env python
"""BASELINE.json configs[2]: 100 M synthetic 150 bp reads (one lane-equivalent) against 500 bp windows.
One rank scores its shard of the 100 M-pair counter-RNG stream in device-resident slices (the whole shard would fit
the 180 GB of HBM; slices keep the run short to allocate).  Under torchrun every rank takes a contiguous shard, no
collective on the data path.  Prints one JSON line: reads/s, GCUPS, and checks every slice through size-independent
properties plus one oracle-checked sample per rank."""
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
    ap.add_argument("--pairs", type=int, default=100_000_000)
    ap.add_argument("--slice", type=int, default=10_000_000)
    ap.add_argument("--sample", type=int, default=20_000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import mini_parallel_b200 as mp
    import oracle_lib as ol
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        saved = os.dup(1); os.dup2(2, 1)                                         # NCCL's version banner goes to stderr
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        dist.barrier(); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    dev = torch.device("cuda", lr)
    eng = mp.Engine(lr)
    rl, wl = 150, 500
    lo, hi = rank * args.pairs // world, (rank + 1) * args.pairs // world       # contiguous shard
    sl = min(args.slice, hi - lo)
    d_q = torch.empty(sl * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(sl * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(sl + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(sl + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(sl * 3, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    score_ms, synth_s, total, ssum = 0.0, 0.0, 0, 0
    ok = True
    t_wall = time.perf_counter()
    for a in range(lo, hi, sl):
        n = min(sl, hi - a)
        t0 = time.perf_counter()
        eng.synth_device(a, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        eng.sync()
        synth_s += time.perf_counter() - t0
        eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * rl, d_r.data_ptr(), d_ro.data_ptr(), n * wl, n, rl, wl, d_out.data_ptr())
        score_ms += eng.last_timings()["device_ms"]
        out = d_out[: n * 3].view(n, 3)
        ok &= bool((out[:, 0] > 150).all() and (out[:, 0] <= 300).all() and (out[:, 1] < rl).all() and (out[:, 2] < wl).all() and (out[:, 1] >= 0).all())
        ok &= eng.last_routing() == {"short": n, "generic": 0, "long": 0}
        ssum += int(out[:, 0].sum(dtype=torch.int64).item())
        if a == lo and args.sample:
            m = min(args.sample, n)
            q = d_q[: m * rl].cpu().numpy(); r = d_r[: m * wl].cpu().numpy()
            exp = ol.batch(q, np.arange(m + 1, dtype=np.uint64) * rl, r, np.arange(m + 1, dtype=np.uint64) * wl, threads=os.cpu_count() or 8, simd=True)
            got = out[:m].cpu().numpy()
            ok &= bool(np.array_equal(got[:, 0], exp["score"]) and np.array_equal(got[:, 1], exp["end_i"]) and np.array_equal(got[:, 2], exp["end_j"]))
        total += n
    wall = time.perf_counter() - t_wall
    t = torch.tensor([score_ms, float(total), float(ssum), float(ok)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        score_ms, total, ssum, ok = float(mx[0]), float(sm[1]), float(sm[2]), bool(mn[3] > 0)
    if rank == 0:
        print(json.dumps({"workload": f"BASELINE.json configs[2]: {int(total)} synthetic 150 bp reads x 500 bp windows over {world} GPU(s), slices of {sl} pairs resident in HBM",
                          "n_gpus": world, "device_seconds_max_over_ranks": round(score_ms / 1e3, 3), "reads_per_s": round(total / (score_ms / 1e3), 1),
                          "gcups": round(total * rl * wl / (score_ms / 1e3) / 1e9, 1), "checks_ok": ok, "mean_score": round(ssum / total, 3),
                          "synth_seconds_rank0": round(synth_s, 2), "wall_seconds_rank0": round(wall, 2)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
