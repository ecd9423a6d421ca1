#!/usr/bin/env python
"""bench.py -- headline benchmark of the Smith-Waterman scoring path (BASELINE.json metric:
GCUPS and reads/s on 150 bp reads, 1/2/4/8 B200, beside the CPU SIMD path).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle SIMD port, all host cores)

Workload (config.workload): BASELINE.json configs[1] -- 1,000,000 synthetic 150 bp reads, each against its own
500 bp window, on ONE B200; with N GPUs every rank scores its own 1 M-pair shard of the same counter-RNG stream
(weak scaling, no data-path collective: pairs are independent).  One step = one pass of the hot
path (2-bit pack + classify + int16x2 DPX kernel + generic kernel) over the whole batch.

  value   : whole-job GCUPS with the ASCII inputs already resident in HBM when the timed region starts
  e2e     : the same through swb_score_batch with HOST (pinned) buffers: H2D + kernels + D2H inside the region
  roofline: the short-read kernel against the DPX integer-issue roofline; second object for
            the HBM-bound packing kernel
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

READ_LEN, WINDOW_LEN = 150, 500
DPX_INSTR_PER_CLK_PER_SM_FALLBACK = 64.0     # measured: profiles/issue_rate_r01.json
INT_ISSUE_PER_CELL = 2.0                     # 4 thread-instructions per int16x2 cell pair


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def load_issue_rate():
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True) if os.path.isdir(os.path.join(ROOT, "profiles")) else []:
        if name.startswith("issue_rate_") and name.endswith(".json"):
            try:
                d = json.load(open(os.path.join(ROOT, "profiles", name)))
                return float(d["rates"]["viaddmnmx_s16x2"]["thread_instr_per_clk_per_sm"]), name
            except Exception:
                pass
    return DPX_INSTR_PER_CLK_PER_SM_FALLBACK, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_ncu_traffic(kernel_tag):
    """dram read+write bytes per launch of the dominant kernel from the committed `ncu --set full` summary."""
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir), reverse=True) if os.path.isdir(pdir) else []:
        if name.startswith(f"ncu_{kernel_tag}_") and name.endswith(".csv"):
            tot, scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for line in open(os.path.join(pdir, name)):
                f = line.strip().split(",")
                if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[1] in scale:
                    tot += float(f[2]) * scale[f[1]]
            if tot > 0:
                return tot, name
    return None, None


def cpu_inputs(first_pair, n_pairs, dist):
    """The first n_pairs of the workload as host arrays (generated in slices: the numpy twin of the device generator
    needs ~4 KB of temporaries per pair)."""
    from mini_parallel_b200 import synth
    q = np.empty(n_pairs * READ_LEN, dtype=np.uint8); r = np.empty(n_pairs * WINDOW_LEN, dtype=np.uint8)
    for a in range(0, n_pairs, 100_000):
        m = min(100_000, n_pairs - a)
        cq, _, cr, _ = synth.make_pairs(first_pair + a, m, READ_LEN, WINDOW_LEN, dist)
        q[a * READ_LEN:(a + m) * READ_LEN] = cq; r[a * WINDOW_LEN:(a + m) * WINDOW_LEN] = cr
    qo = np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(READ_LEN); ro = np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(WINDOW_LEN)
    return q, qo, r, ro


def cpu_simd_time(inputs, threads, steps, warmup=1):
    """Seconds per pass of oracle/sw_simd.c (all pairs of `inputs`, `threads` host threads), list of `steps` timings."""
    import oracle_lib as ol
    q, qo, r, ro = inputs
    for _ in range(warmup):
        ol.batch(q, qo, r, ro, threads=threads, simd=True)
    out = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ol.batch(q, qo, r, ro, threads=threads, simd=True)
        out.append(time.perf_counter() - t0)
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference has no CPU scoring path and cannot be built here (no
    Rust / OpenCL); the arm times this repo's CPU port of the same scoring function (oracle/sw_simd.c,
    bit-exact with the oracle) on all host cores.  One step = one pass over the SAME pairs one GPU scores per step
    (--ref-pairs, default = --pairs); the inputs are generated once, outside the timed region, as they are for the GPU."""
    if rank != 0:
        return
    import oracle_lib as ol
    threads = os.cpu_count() or 1
    n = args.ref_pairs if args.ref_pairs > 0 else args.pairs
    inputs = cpu_inputs(0, n, args.dist)
    times = cpu_simd_time(inputs, threads, args.steps, warmup=max(1, min(args.warmup, 2)))
    t_total = sum(times)
    gcups = n * args.steps * READ_LEN * WINDOW_LEN / t_total / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": round(gcups, 3), "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * t_total / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "reads_per_s": round(n * args.steps / t_total, 1),
        "config": workload_config(args, 1, n),
        "cpu_baseline": {"value": round(gcups, 3), "unit": "GCUPS", "cores": threads, "kind": "port", "isa": ol.simd_isa(),
                         "sample": f"pairs 0..{n - 1} of the configs[1] stream, every step a full pass ({args.steps} steps, "
                                   f"{min(times):.3f}-{max(times):.3f} s each); oracle/sw_simd.c, {threads} threads sharing one work cursor"},
        "e2e": {"value": round(gcups, 3), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, pairs_per_step):
    """config.workload is the same string for both arms (the driver compares it)."""
    return {"workload": ("BASELINE.json configs[1]: 1M synthetic 150bp reads vs 500bp windows per GPU (inter-task int16x2 DPX kernel), "
                         "related reads (1% subst, 0.1% ins, 0.1% del)") if args.dist == 0 else "BASELINE.json configs[1] shape, unrelated reads",
            "pairs_per_gpu": pairs_per_step, "read_len": READ_LEN, "window_len": WINDOW_LEN, "distribution": args.dist}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--dist", type=int, default=0, help="0 = related reads, 1 = unrelated")
    ap.add_argument("--ref-pairs", type=int, default=200_000, help="pairs per step of the CPU arm")
    ap.add_argument("--cpu-budget-s", type=float, default=12.0)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary measurements (unrelated reads, 1 % N, configs[0] shape)")
    ap.add_argument("--long-pairs", type=int, default=2368, help="pairs of the auxiliary 10 kb x 10 kb measurement (0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    import mini_parallel_b200 as mp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; stdout carries exactly one JSON line
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    warmup = max(args.warmup, 3)
    n, rl, wl = args.pairs, READ_LEN, WINDOW_LEN
    from mini_parallel_b200 import sharding
    first_pair, _ = sharding.shard_range(rank, world, world * n)   # this rank's contiguous shard of the counter-RNG stream
    eng = mp.Engine(local_rank)
    if args.variant >= 0:
        eng.set_short_variant(args.variant)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))

    # ---- device-resident inputs (ASCII, as a FASTQ chunk would arrive) ----
    dev = torch.device("cuda", local_rank)
    d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev)
    d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * 3, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    eng.synth_device(first_pair, n, rl, wl, args.dist, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
    eng.sync()

    def step_device():
        eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * rl, d_r.data_ptr(), d_ro.data_ptr(), n * wl, n, rl, wl,
                               d_out.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return sharding.reduce_scalars([x], "max")[0]

    # ---- timed region 1: device-resident ----
    for _ in range(warmup):
        step_device()
    eng.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    short_ms, pack_ms, generic_ms, launches = [], [], [], 0
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
            t = eng.last_timings()                          # CUDA events around each kernel, on the launching stream
            short_ms.append(t["short_ms"]); pack_ms.append(t["pack_classify_ms"]); generic_ms.append(t["generic_ms"])
            launches += t["kernels"]
        e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    routing = eng.last_routing()

    # ---- the packing kernel alone (HBM roofline of pack2bit_kernel; classify and launch gaps excluded) ----
    pk_words = torch.empty((n * wl + 15) // 16 + 16, dtype=torch.int32, device=dev)
    pk_bits = torch.empty(pk_words.numel() // 32 + 16, dtype=torch.int32, device=dev)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(3):
            eng.pack2bit_device(d_r.data_ptr(), n * wl, pk_words.data_ptr(), pk_bits.data_ptr())
        p0.record(stream)
        for _ in range(10):
            eng.pack2bit_device(d_r.data_ptr(), n * wl, pk_words.data_ptr(), pk_bits.data_ptr())
        p1.record(stream)
    torch.cuda.synchronize()
    pack_alone_ms = p0.elapsed_time(p1) / 10
    del pk_words, pk_bits

    # ---- timed region 2: end to end through the host API (pinned host buffers) ----
    numa_node = mp.load_library().swb_numa_prefer_device(local_rank)   # pinned buffers on the GPU's own NUMA node (a preference)
    h_q = torch.empty(n * rl, dtype=torch.uint8).pin_memory()
    h_r = torch.empty(n * wl, dtype=torch.uint8).pin_memory()
    h_qo = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    h_ro = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    h_out = torch.empty(n * 3, dtype=torch.int32).pin_memory()
    h_q.copy_(d_q); h_r.copy_(d_r); h_qo.copy_(d_qo); h_ro.copy_(d_ro)
    torch.cuda.synchronize()
    lib = mp.load_library()
    lib.swb_numa_reset()

    def step_host():
        rc = lib.swb_score_batch(eng._h, h_q.data_ptr(), h_qo.data_ptr(), h_r.data_ptr(), h_ro.data_ptr(), n, h_out.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.swb_last_error().decode())

    for _ in range(2):
        step_host()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        g0.record(stream)
        for _ in range(e2e_steps):
            step_host()
        g1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(g0.elapsed_time(g1))
    host_same = bool(torch.equal(h_out.to(dev), d_out))
    e2e_t = eng.last_timings()

    # ---- timed region 3: reads from the host against windows of a DEVICE-RESIDENT reference ----
    # (north_star: windows must not stream over PCIe.)  The reference is the concatenation of the shard's windows,
    # uploaded and packed once outside the timed region like a genome would be; per step only the reads, their
    # offsets and one (start, len) per read cross PCIe.
    h_ws = (torch.arange(n, dtype=torch.int64) * wl).pin_memory()
    h_wl = torch.full((n,), wl, dtype=torch.int32).pin_memory()
    h_out2 = torch.empty(n * 3, dtype=torch.int32).pin_memory()
    rc = lib.swb_set_reference(eng._h, h_r.data_ptr(), n * wl)
    if rc != 0:
        raise RuntimeError(lib.swb_last_error().decode())

    def step_ref():
        rc = lib.swb_score_batch_vs_reference(eng._h, h_q.data_ptr(), h_qo.data_ptr(), n, h_ws.data_ptr(), h_wl.data_ptr(),
                                              h_out2.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.swb_last_error().decode())

    for _ in range(2):
        step_ref()
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        r0.record(stream)
        for _ in range(e2e_steps):
            step_ref()
        r1.record(stream)
    barrier()
    ref_ms = max_over_ranks(r0.elapsed_time(r1))
    ref_same = bool(torch.equal(h_out2, h_out))

    # ---- auxiliary device-resident measurements on rank 0 (both distributions, the 1 % N variant,
    #      the configs[0] shape); three steps each, the first one is warm-up ----
    aux = {}
    if rank == 0 and not args.no_aux:
        def timed(tag, nn, qlen, wlen, dq, dqo, dr, dro, note):
            ms = []
            for s_ in range(3):
                eng.score_batch_device(dq.data_ptr(), dqo.data_ptr(), nn * qlen, dr.data_ptr(), dro.data_ptr(), nn * wlen, nn, qlen, wlen,
                                       d_out.data_ptr())
                t = eng.last_timings()
                if s_:
                    ms.append(t["device_ms"])
            m = statistics.mean(ms)
            aux[tag] = {"workload": note, "ms_per_step": round(m, 3), "gcups": round(float(nn) * qlen * wlen / (m * 1e-3) / 1e9, 1),
                        "reads_per_s": round(nn / (m * 1e-3), 1), "routing": eng.last_routing()}
        # unrelated reads (distribution U): scores ~20, stresses the floor path
        eng.synth_device(first_pair, n, rl, wl, 1, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        timed("unrelated_reads", n, rl, wl, d_q, d_qo, d_r, d_ro, f"{n} pairs {rl}x{wl}, reads independent of their windows")
        # 1 % of the reads carry one 'N': byte-compare routing (smith_waterman.cl:114 compares raw bytes)
        eng.synth_device(first_pair, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        eng.sync()
        g = torch.Generator(device=dev); g.manual_seed(0xB200)
        sel = torch.randperm(n, device=dev, generator=g)[: n // 100].to(torch.int64)
        d_q[sel * rl + torch.randint(0, rl, (sel.numel(),), device=dev, generator=g)] = ord("N")
        torch.cuda.synchronize()
        timed("one_percent_N", n, rl, wl, d_q, d_qo, d_r, d_ro, f"{n} pairs {rl}x{wl}, related reads, 1 % of the reads contain one N")
        # BASELINE.json configs[0] shape: 10 k reads against 1 kb windows
        c1 = 10_000
        eng.synth_device(0, c1, rl, 1000, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        timed("config0_shape", c1, rl, 1000, d_q, d_qo, d_r, d_ro, "BASELINE.json configs[0] shape: 10000 reads of 150 bp x 1 kb windows (one small launch)")
        eng.synth_device(first_pair, n, rl, wl, args.dist, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        eng.sync()

    # ---- auxiliary: BASELINE.json configs[3] shape (10 kb x 10 kb pairs, sw_long_kernel), one resident wave of pairs ----
    aux_long = None
    if args.long_pairs > 0 and rank == 0:
        ln, ll = args.long_pairs, 10_000
        del d_q, d_r, d_qo, d_ro
        l_q = torch.empty(ln * ll, dtype=torch.uint8, device=dev); l_r = torch.empty(ln * ll, dtype=torch.uint8, device=dev)
        l_qo = torch.empty(ln + 1, dtype=torch.int64, device=dev); l_ro = torch.empty(ln + 1, dtype=torch.int64, device=dev)
        l_out = torch.empty(ln * 3, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        eng.synth_device(0, ln, ll, ll, 0, l_q.data_ptr(), l_qo.data_ptr(), l_r.data_ptr(), l_ro.data_ptr())
        lms = []
        for s_ in range(3):
            eng.score_batch_device(l_q.data_ptr(), l_qo.data_ptr(), ln * ll, l_r.data_ptr(), l_ro.data_ptr(), ln * ll, ln, ll, ll,
                                   l_out.data_ptr())
            t = eng.last_timings()
            if s_:
                lms.append(t["device_ms"])
        lcells = float(ln) * ll * ll
        aux_long = {"workload": f"BASELINE.json configs[3] shape: {ln} pairs 10000 x 10000 (one resident wave of warps), related reads",
                    "kernel": "sw_long_kernel (32-bit banded wavefront, one warp per pair)", "ms_per_step": round(statistics.mean(lms), 3),
                    "gcups": round(lcells / (statistics.mean(lms) * 1e-3) / 1e9, 1), "routing": eng.last_routing(),
                    "peak_gcups_at_4_instr_per_cell": None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    uniform_offsets = os.environ.get("SWB_UNIFORM_OFFSETS", "1") != "0"
    cells_step = float(n) * rl * wl
    gcups = world * cells_step * args.steps / (ms_total * 1e-3) / 1e9
    e2e_gcups = world * cells_step * e2e_steps / (e2e_ms * 1e-3) / 1e9
    ref_gcups = world * cells_step * e2e_steps / (ref_ms * 1e-3) / 1e9
    peaks = load_peaks()
    rate, rate_src = load_issue_rate()
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_gcups = sms * peaks["sm_max_mhz"] * 1e6 * rate / INT_ISSUE_PER_CELL / 1e9
    traffic, traffic_src = load_ncu_traffic("stream_kernel")
    if aux_long:
        aux_long["peak_gcups_at_4_instr_per_cell"] = round(sms * peaks["sm_max_mhz"] * 1e6 * rate / 4.0 / 1e9, 1)
    k_ms = statistics.mean(short_ms)
    k_gcups = cells_step / (k_ms * 1e-3) / 1e9
    p_ms = statistics.mean(pack_ms)
    pack_bytes = 1.25 * n * (rl + wl)                      # 1 B read + 0.25 B written per base
    cpu_val, cpu_pairs, cpu_s = cpu_simd_gcups(0, 2_000_000, args.dist, os.cpu_count() or 1, args.cpu_budget_s)
    import oracle_lib as ol
    from mini_parallel_b200 import synth as _synth
    sq, sqo, sr, sro = _synth.make_pairs(0, 2000, rl, wl, args.dist)
    t0 = time.perf_counter()
    ol.batch(sq, sqo, sr, sro, threads=1, simd=False)
    scalar_gcups = 2000 * rl * wl / (time.perf_counter() - t0) / 1e9

    line = {
        "metric": "GCUPS", "value": round(gcups, 2), "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "reads_per_s": round(world * n * args.steps / (ms_total * 1e-3), 1),
        "config": {"workload": "BASELINE.json configs[1]: 1M synthetic 150bp reads vs 500bp windows per GPU (inter-task int16x2 DPX kernel), "
                               "related reads (1% subst, 0.1% ins, 0.1% del)" if args.dist == 0 else
                               "BASELINE.json configs[1] shape, unrelated reads",
                   "pairs_per_gpu": n, "read_len": rl, "window_len": wl, "distribution": args.dist, "sharding": f"{world} x independent shard",
                   "l2_policy": f"inputs larger than L2 ({(n * (rl + wl)) >> 20} MiB ASCII per step vs 126 MB L2)",
                   "short_variant": args.variant, "routing": routing, "host_path_equals_device_path": host_same},
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": round(e2e_gcups, 2), "unit": "GCUPS", "reads_per_s": round(world * n * e2e_steps / (e2e_ms * 1e-3), 1),
                "steps": e2e_steps, "ms_per_step": round(e2e_ms / e2e_steps, 3),
                # reads and windows all have one length: the library sends no offsets for such chunks (it writes k * length
                # on the device), unless SWB_UNIFORM_OFFSETS=0
                "h2d_bytes_per_step": int(h_q.numel() + h_r.numel() + (8 * (h_qo.numel() + h_ro.numel()) if not uniform_offsets else 0)),
                "d2h_bytes_per_step": int(4 * h_out.numel()),
                "api": "swb_score_batch (ASCII reads + windows from pinned host memory, chunks pipelined over 3 streams; offsets of "
                       "uniform-length chunks are generated on the device)",
                "pinned_numa_node": numa_node,
                "stage_ms_sum_over_chunks": {k: round(v, 3) for k, v in e2e_t.items() if k.endswith("_ms")}},
        "e2e_resident_reference": {"value": round(ref_gcups, 2), "unit": "GCUPS",
                                   "reads_per_s": round(world * n * e2e_steps / (ref_ms * 1e-3), 1), "steps": e2e_steps,
                                   "ms_per_step": round(ref_ms / e2e_steps, 3),
                                   "h2d_bytes_per_step": int(h_q.numel() + (8 * h_qo.numel() if not uniform_offsets else 0) + 8 * h_ws.numel() + 4 * h_wl.numel()),
                                   "d2h_bytes_per_step": int(4 * h_out2.numel()), "equals_e2e_results": ref_same,
                                   "api": "swb_score_batch_vs_reference (reads from pinned host memory, windows of a reference "
                                          "uploaded once)"},
        "roofline": {"bound": "int_issue", "kernel": "sw_stream_kernel" if args.variant < 0 or args.variant >= 4 else "sw_short_kernel", "achieved": round(k_gcups, 1), "peak": round(peak_gcups, 1),
                     "unit": "GCUPS", "frac": round(k_gcups / peak_gcups, 4),
                     "traffic": None if traffic is None else int(traffic),       # dram read+write bytes of one launch (ncu --set full)
                     "traffic_source": None if traffic is None else f"profiles/{traffic_src}",
                     "algorithmic_bytes_per_launch": int(n * (rl / 4 + wl / 4 + 32 + 12)),
                     "kernel_ms": round(k_ms, 4), "kernel_share_of_step": round(k_ms / (ms_total / args.steps), 4),
                     "peak_is": f"{sms} SMs x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} MEASURED_PEAKS.json) x {rate:.2f} DPX s16x2 "
                                f"thread-instr/clk/SM (measured, {rate_src}) / {INT_ISSUE_PER_CELL} instr per cell"},
        "roofline_pack": {"bound": "hbm", "kernel": "pack2bit_kernel x2 + classify_kernel", "achieved": round(pack_bytes / (p_ms * 1e-3) / 1e9, 1),
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(pack_bytes / (p_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                          "kernel_ms": round(p_ms, 4), "traffic": None, "peak_is": f"{peaks['source']} copy bandwidth",
                          "pack2bit_kernel_alone": {"bytes": int(1.25 * n * wl), "ms": round(pack_alone_ms, 4),
                                                    "achieved": round(1.25 * n * wl / (pack_alone_ms * 1e-3) / 1e9, 1),
                                                    "frac": round(1.25 * n * wl / (pack_alone_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                                                    "what": "the window array of the step (500 MB in, 125 MB out), 10 back-to-back launches"}},
        "aux_long_pairs": aux_long,
        "aux": aux,
        "cpu_baseline": {"value": round(cpu_val, 3), "unit": "GCUPS", "cores": os.cpu_count() or 1, "kind": "port", "isa": ol.simd_isa(),
                         "scalar_oracle_gcups_1_core": round(scalar_gcups, 3),
                         "sample": f"first {cpu_pairs} pairs of the same counter-RNG stream, {cpu_s:.1f} s wall on {os.cpu_count() or 1} threads = "
                                   f"{cpu_s * (os.cpu_count() or 1):.0f} core-seconds (oracle/sw_simd.c)"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
