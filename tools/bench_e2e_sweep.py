#!/usr/bin/env python
"""Sweep of the host-path chunk schedule (swb_set_chunking / swb_set_chunk_ramp) on BASELINE.json configs[1]:
1 M pairs 150 x 500 from pinned host memory through swb_score_batch and swb_score_batch_vs_reference.  Prints one JSON
line per setting; the first line is the plain pinned H2D copy rate of the same bytes (the PCIe floor of the step)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--chunk-mb", default="16,32,64,128")
    ap.add_argument("--lanes", default="3", help="comma list of SWB_LANES values (a new engine per value)")
    args = ap.parse_args()
    import torch
    import mini_parallel_b200 as mp

    n, rl, wl = args.pairs, 150, 500
    dev = torch.device("cuda", 0)
    eng = mp.Engine(0)
    lib = mp.load_library()
    state = {"eng": eng}
    d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
    eng.synth_device(0, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
    eng.sync()
    h_q = torch.empty(n * rl, dtype=torch.uint8).pin_memory(); h_r = torch.empty(n * wl, dtype=torch.uint8).pin_memory()
    h_qo = torch.empty(n + 1, dtype=torch.int64).pin_memory(); h_ro = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    h_out = torch.empty(n * 3, dtype=torch.int32).pin_memory()
    h_q.copy_(d_q); h_r.copy_(d_r); h_qo.copy_(d_qo); h_ro.copy_(d_ro)
    torch.cuda.synchronize()

    # PCIe floor: the step's bytes as two plain pinned copies
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(4):
        if k == 1:
            e0.record()
        d_q.copy_(h_q, non_blocking=True); d_r.copy_(h_r, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    copy_ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"plain_h2d_ms": round(copy_ms, 3), "gb_per_s": round(n * (rl + wl) / copy_ms / 1e6, 1)}), flush=True)

    h_ws = (torch.arange(n, dtype=torch.int64) * wl).pin_memory()
    h_wl = torch.full((n,), wl, dtype=torch.int32).pin_memory()
    if lib.swb_set_reference(eng._h, h_r.data_ptr(), n * wl) != 0:
        raise RuntimeError(lib.swb_last_error().decode())

    def host():
        if lib.swb_score_batch(state["eng"]._h, h_q.data_ptr(), h_qo.data_ptr(), h_r.data_ptr(), h_ro.data_ptr(), n, h_out.data_ptr()) != 0:
            raise RuntimeError(lib.swb_last_error().decode())

    def vs_ref():
        if lib.swb_score_batch_vs_reference(state["eng"]._h, h_q.data_ptr(), h_qo.data_ptr(), n, h_ws.data_ptr(), h_wl.data_ptr(), h_out.data_ptr()) != 0:
            raise RuntimeError(lib.swb_last_error().decode())

    first = None
    for lanes in [int(x) for x in args.lanes.split(",")]:
      os.environ["SWB_LANES"] = str(lanes)
      eng = mp.Engine(0); state["eng"] = eng
      if lib.swb_set_reference(eng._h, h_r.data_ptr(), n * wl) != 0:
          raise RuntimeError(lib.swb_last_error().decode())
      for mb in [int(x) for x in args.chunk_mb.split(",")]:
        for ramp in (0, 2):
            eng.set_chunking(mb << 20, 16384); eng.set_chunk_ramp(ramp)
            row = {"lanes": lanes, "chunk_mb": mb, "ramp": ramp}
            for tag, fn in (("host", host), ("vs_reference", vs_ref)):
                fn(); fn()
                e0.record()
                for _ in range(args.steps):
                    fn()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                row[tag] = {"ms": round(ms, 3), "gcups": round(n * rl * wl / ms / 1e6, 1)}
                res = h_out.clone()
                if first is None:
                    first = res
                row[tag]["same_results"] = bool(torch.equal(res, first))
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
