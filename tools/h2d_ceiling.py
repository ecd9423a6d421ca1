#!/usr/bin/env python
"""The box's plain host->device copy ceiling with N ranks copying at once.

Every rank does nothing but cudaMemcpyAsync of one pinned buffer (default: the 650 MB of ASCII one bench.py step
uploads) to its own GPU, REPS times back to back, all ranks between the same two barriers.  Timed with CUDA events on
the copying stream, max over ranks.  Rank 0 prints one JSON line: aggregate and per-GPU GB/s.

    python tools/h2d_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=650.0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--d2h-mb", type=float, default=12.0, help="bytes copied back per rep on a second stream (0 = none)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
            dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    n = int(args.mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(65)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    nb = int(args.d2h_mb * 1e6)
    hb = torch.empty(max(nb, 1), dtype=torch.uint8).pin_memory()
    db = torch.zeros(max(nb, 1), dtype=torch.uint8, device="cuda")
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def run(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_up)
        for _ in range(reps):
            with torch.cuda.stream(s_up):
                d.copy_(h, non_blocking=True)
            if nb:
                with torch.cuda.stream(s_down):
                    hb.copy_(db, non_blocking=True)
        e1.record(s_up)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    run(3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = run(args.reps)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        gbs = world * n * args.reps / (ms * 1e-3) / 1e9
        print(json.dumps({"what": "plain pinned H2D copy, all ranks at once", "n_gpus": world, "mb_per_copy": args.mb, "reps": args.reps,
                          "d2h_mb_alongside": args.d2h_mb, "ms_per_copy_max_over_ranks": round(ms / args.reps, 3),
                          "aggregate_h2d_gbs": round(gbs, 2), "per_gpu_h2d_gbs": round(gbs / world, 2), "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
