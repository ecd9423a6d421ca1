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


def cpu_inputs(first_pair, n_pairs, dist, ref_bases):
    """The first n_pairs of the workload as host arrays (generated in slices: the numpy twin of the device generator
    needs ~4 KB of temporaries per pair)."""
    from mini_parallel_b200 import synth
    ref = synth.synth_reference(ref_bases)
    q = np.empty(n_pairs * READ_LEN, dtype=np.uint8); r = np.empty(n_pairs * WINDOW_LEN, dtype=np.uint8)
    for a in range(0, n_pairs, 100_000):
        m = min(100_000, n_pairs - a)
        cq, _, cr, _, _ = synth.make_pairs_ref(ref, first_pair + a, m, READ_LEN, WINDOW_LEN, dist)
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
    inputs = cpu_inputs(0, n, args.dist, args.ref_bases)
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
            "pairs_per_gpu": pairs_per_step, "read_len": READ_LEN, "window_len": WINDOW_LEN, "distribution": args.dist,
            "windows": f"cut from a synthetic {args.ref_bases} base reference at counter-RNG positions (they overlap, as candidate windows of a genome do)"}


def checksum64(out_i32, first_index):
    """64-bit checksum of (pair index, score, end_i, end_j) over a slice of results (torch int32 tensor (n,3) on the
    device): a wrapping sum of per-pair hashes, so it does not depend on how the pairs are sharded."""
    import torch
    def c(x):                                             # Python int -> two's complement int64
        return x - (1 << 64) if x >= (1 << 63) else x
    n = out_i32.shape[0]
    idx = torch.arange(first_index, first_index + n, dtype=torch.int64, device=out_i32.device)
    o = out_i32.to(torch.int64)
    h = idx * c(0x9E3779B97F4A7C15) + o[:, 0] * c(0xBF58476D1CE4E5B9) + o[:, 1] * c(0x94D049BB133111EB) + o[:, 2] * c(0xD6E8FEB86659FD93)
    h = (h ^ (h >> 29)) * c(0xFF51AFD7ED558CCD)
    return int(h.sum().item()) & ((1 << 64) - 1)


# configs[2] at full size on the CPU checker: the checksum64 of ITS results over all 100 M pairs (tests/tools/checksum_config2_cpu.py,
# profiles/config2_full_checksum_cpu_r02.json).  The GPU leg's checksum equal to it = full-population parity for configs[2].
CONFIG2_CPU_CHECKSUM = {(100_000_000, 0, 150, 500): "5ec5834cd0ec6b2a"}


def annotate_strong(res, total, dist_, rl, wl):
    """Adds the CPU checker's full-size checksum beside the GPU's, where one is on record for this exact workload."""
    ref = CONFIG2_CPU_CHECKSUM.get((int(total), int(dist_), int(rl), int(wl)))
    if ref is not None:
        res["cpu_checker_checksum64"] = ref
        res["cpu_checker_checksum_source"] = "profiles/config2_full_checksum_cpu_r02.json (tests/tools/checksum_config2_cpu.py: oracle/sw_simd.c over all pairs)"
        res["equals_cpu_checker_on_all_pairs"] = res.get("checksum64") == ref
    return res


def leg_strong(args, eng, dev, rank, world, barrier, reduce_scalars):
    """BASELINE.json configs[2]: 100 M reads (one lane-equivalent, aligner.rs:214-215 counts 51.8 M per lane file) sharded
    over the ranks -- STRONG scaling: the total is fixed, every rank scores total/N pairs in device-resident slices.  Timed:
    the scoring of every slice (CUDA events on the engine's stream); the generator that fills the slice is not."""
    import torch
    import oracle_lib as ol
    from mini_parallel_b200 import sharding
    total, rl, wl = args.strong_pairs, READ_LEN, WINDOW_LEN
    lo, hi = sharding.shard_range(rank, world, total)
    sl = max(1, min(args.strong_slice, hi - lo))
    d_q = torch.empty(sl * rl, dtype=torch.uint8, device=dev); d_r = torch.empty(sl * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(sl + 1, dtype=torch.int64, device=dev); d_ro = torch.empty(sl + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(sl * 3, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    ev, csum, ssum, launches, checked, ok, lib_ms = [], 0, 0, 0, 0, True, 0.0
    # warm-up: the library's arenas grow to the slice size here (cudaMalloc / cudaFree are host calls that would otherwise sit
    # between the timed region's first event and its first kernel)
    m0 = min(sl, hi - lo)
    eng.synth_device(lo, m0, rl, wl, args.dist, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
    for _ in range(2):
        eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), m0 * rl, d_r.data_ptr(), d_ro.data_ptr(), m0 * wl, m0, rl, wl, d_out.data_ptr())
    eng.sync()
    barrier()
    t_wall = time.perf_counter()
    for a in range(lo, hi, sl):
        m = min(sl, hi - a)
        eng.synth_device(a, m, rl, wl, args.dist, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), m * rl, d_r.data_ptr(), d_ro.data_ptr(), m * wl, m, rl, wl, d_out.data_ptr())
            e1.record(stream)
        ev.append((e0, e1))
        eng.sync()
        lt = eng.last_timings()
        launches += lt["kernels"]; lib_ms += lt["device_ms"]
        out = d_out[: m * 3].view(m, 3)
        csum = (csum + checksum64(out, a)) & ((1 << 64) - 1)
        ssum += int(out[:, 0].sum(dtype=torch.int64).item())
        if a == lo and args.strong_check > 0:             # every pair of the shard's first `strong_check` against the CPU oracle
            k = min(args.strong_check, m)
            hq = d_q[: k * rl].cpu().numpy(); hr = d_r[: k * wl].cpu().numpy()
            qo = np.arange(k + 1, dtype=np.uint64) * np.uint64(rl); ro = np.arange(k + 1, dtype=np.uint64) * np.uint64(wl)
            exp = ol.batch(hq, qo, hr, ro, threads=os.cpu_count() or 1, simd=True)
            got = out[:k].cpu().numpy()
            ok &= bool(np.array_equal(got, np.stack([exp["score"], exp["end_i"], exp["end_j"]], axis=1)))
            checked = k
    torch.cuda.synchronize()
    wall_local = time.perf_counter() - t_wall
    ms_local = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    barrier()
    ms = reduce_scalars([ms_local], "max")[0]
    wall = reduce_scalars([wall_local], "max")[0]
    # sums over ranks: 64-bit values travel as two 32-bit halves (exact in the float64 the reduction uses)
    parts = reduce_scalars([float(csum & 0xFFFFFFFF), float(csum >> 32), float(ssum), float(checked), 0.0 if ok else 1.0, float(launches)], "sum")
    csum_all = (int(parts[0]) + (int(parts[1]) << 32)) & ((1 << 64) - 1)
    del d_q, d_r, d_qo, d_ro, d_out
    cells = float(total) * rl * wl
    res = {"workload": f"BASELINE.json configs[2]: {total} synthetic {rl} bp reads vs {wl} bp windows, one lane-equivalent, sharded over {world} GPU(s)",
            "scaling": "strong", "pairs_total": total, "pairs_per_gpu": (total + world - 1) // world, "slice_pairs": sl,
            "score_ms_max_over_ranks": round(ms, 3), "gcups": round(cells / (ms * 1e-3) / 1e9, 1), "reads_per_s": round(total / (ms * 1e-3), 1),
            "score_ms_by_library_events_this_rank": round(lib_ms, 3), "wall_s_with_generation_and_checks": round(wall, 3),
            "checksum64": f"{csum_all:016x}", "checksum_is": "wrapping sum over ALL pairs of hash(pair index, score, end_i, end_j): equal at every N",
            "mean_score": round(parts[2] / total, 3), "oracle_checked_pairs": int(parts[3]), "oracle_checked_pairs_per_rank": checked,
            "oracle_equal": parts[4] == 0.0, "gpu_launches": int(parts[5])}
    try:
        annotate_strong(res, total, args.dist, rl, wl)
    except Exception:                                       # an annotation never costs the line
        pass
    return res


def leg_long(args, eng, dev):
    """BASELINE.json configs[3]: 10 000 pairs of 10 kb x 10 kb through sw_long_kernel (and the one-resident-wave figure)."""
    import torch
    ll = 10_000
    res = {}
    for tag, ln in (("full", args.long_pairs), ("one_wave", args.long_wave)):
        if ln <= 0:
            continue
        l_q = torch.empty(ln * ll, dtype=torch.uint8, device=dev); l_r = torch.empty(ln * ll, dtype=torch.uint8, device=dev)
        l_qo = torch.empty(ln + 1, dtype=torch.int64, device=dev); l_ro = torch.empty(ln + 1, dtype=torch.int64, device=dev)
        l_out = torch.empty(ln * 3, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        eng.synth_device(0, ln, ll, ll, 0, l_q.data_ptr(), l_qo.data_ptr(), l_r.data_ptr(), l_ro.data_ptr())
        lms, k = [], 0
        for s_ in range(3):
            eng.score_batch_device(l_q.data_ptr(), l_qo.data_ptr(), ln * ll, l_r.data_ptr(), l_ro.data_ptr(), ln * ll, ln, ll, ll, l_out.data_ptr())
            t = eng.last_timings()
            k = t["kernels"]
            if s_:
                lms.append(t["device_ms"])
        m = statistics.mean(lms)
        res[tag] = {"pairs": ln, "ms_per_step": round(m, 3), "gcups": round(float(ln) * ll * ll / (m * 1e-3) / 1e9, 1), "routing": eng.last_routing(),
                    "mean_score": round(float(l_out.view(ln, 3)[:, 0].to(torch.float64).mean().item()), 1), "kernels_per_step": k}
        del l_q, l_r, l_qo, l_ro, l_out
    return res


def _bgzf_part(job):
    """One synthetic FASTQ part as BGZF bytes (worker process)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_wgs
    path, fi, n_reads, ref_len = job
    bench_wgs.make_file((path, fi, n_reads, ref_len, READ_LEN, WINDOW_LEN, 1, True, False))
    return path


def leg_bgzf(args, eng, lib, mp):
    """FASTQ ingest on the GPU: BGZF-compressed FASTQ bytes in pinned host memory ->
    swb_fastq_bgzf_score (H2D of the compressed bytes, inflate, index, pack, score against the resident reference, score
    sum back): reads/s from compressed bytes.  Parts of --bgzf-reads/8 reads, the next part prefetched while one is scored."""
    import tempfile
    from concurrent.futures import ProcessPoolExecutor
    import torch
    from mini_parallel_b200 import bgzf, synth
    ref_len, parts = args.ref_bases, 8
    per = max(1, args.bgzf_reads // parts)
    tmp = tempfile.mkdtemp(prefix="swb_bgzf_")
    jobs = [(os.path.join(tmp, f"part{k}.fastq.gz"), k, per, ref_len) for k in range(parts)]
    t0 = time.perf_counter()
    import multiprocessing
    with ProcessPoolExecutor(max_workers=min(parts, os.cpu_count() or 1), mp_context=multiprocessing.get_context("spawn")) as ex:   # no fork of a CUDA process
        paths = list(ex.map(_bgzf_part, jobs))
    gen_s = time.perf_counter() - t0
    segs = []
    for path in paths:
        raw = open(path, "rb").read()
        os.unlink(path)
        blocks, used = bgzf.walk(raw)
        lo, hi = blocks[0][0], blocks[-1][0] + blocks[-1][1]
        comp = torch.frombuffer(bytearray(raw[lo:hi]), dtype=torch.uint8).pin_memory()
        blk = np.zeros(len(blocks), dtype=np.dtype([("in_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4")]))
        blk["in_off"] = [b[0] - lo for b in blocks]; blk["in_len"] = [b[1] for b in blocks]; blk["out_len"] = [b[2] for b in blocks]
        segs.append((comp, torch.from_numpy(blk.view(np.uint8).copy()).pin_memory(), len(blocks), sum(b[2] for b in blocks)))
    os.rmdir(tmp)
    eng.set_reference(synth.synth_reference(ref_len))
    import ctypes
    cout = np.zeros(1 << 16, dtype=np.uint8)
    ssum, nr, nb, nl, cl, st = ctypes.c_int64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()

    def one_pass():
        tot_s, tot_r, launches = 0, 0, 0
        for k, (comp, blk, nblk, _) in enumerate(segs):
            if k + 1 < len(segs):
                nx = segs[k + 1]
                lib.swb_fastq_bgzf_prefetch(eng._h, nx[0].data_ptr(), nx[0].numel(), nx[1].data_ptr(), nx[2])
            rc = lib.swb_fastq_bgzf_score(eng._h, comp.data_ptr(), comp.numel(), blk.data_ptr(), nblk, None, 0, 1, k, 0, WINDOW_LEN,
                                          ctypes.byref(ssum), ctypes.byref(nr), ctypes.byref(nb), ctypes.byref(nl), cout.ctypes.data, cout.size,
                                          ctypes.byref(cl), ctypes.byref(st))
            if rc != 0 or st.value != 0:
                raise RuntimeError("swb_fastq_bgzf_score: " + lib.swb_last_error().decode() + f" status {st.value}")
            tot_s += ssum.value; tot_r += nr.value; launches += eng.last_timings()["kernels"]
        return tot_s, tot_r, launches
    one_pass()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        tot_s, tot_r, launches = one_pass()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    gz = sum(s_[0].numel() for s_ in segs); text = sum(s_[3] for s_ in segs)
    return {"workload": f"{parts} BGZF parts x {per} reads of {READ_LEN} bp (gzip -1, constant qualities), each read vs a {WINDOW_LEN} bp window of a "
                        f"{ref_len} base resident reference; compressed bytes start in pinned host memory",
            "api": "swb_fastq_bgzf_prefetch + swb_fastq_bgzf_score (inflate_bgzf_kernel, fq_* kernels, pack, sw_stream_kernel)",
            "reads": int(tot_r), "ms_per_pass": round(dt * 1e3, 3), "reads_per_s": round(tot_r / dt, 1), "gcups": round(tot_r * READ_LEN * WINDOW_LEN / dt / 1e9, 1),
            "compressed_mb": round(gz / 1e6, 1), "text_mb": round(text / 1e6, 1), "text_gb_per_s": round(text / dt / 1e9, 2),
            "mean_score_per_read": round(tot_s / max(tot_r, 1), 2), "kernels_per_pass": launches, "generate_s": round(gen_s, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--dist", type=int, default=0, help="0 = related reads, 1 = unrelated")
    ap.add_argument("--ref-pairs", type=int, default=0, help="pairs per step of the CPU arm (0 = --pairs: the same pairs one GPU scores per step)")
    ap.add_argument("--cpu-passes", type=int, default=3, help="passes of the cpu_baseline leg over the same pairs")
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--ref-bases", type=int, default=16_000_000, help="synthetic reference the windows of the workload are cut from")
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary measurements (unrelated reads, 1 % N, configs[0] shape)")
    ap.add_argument("--long-pairs", type=int, default=10_000, help="pairs of the configs[3] leg, 10 kb x 10 kb (0 = skip)")
    ap.add_argument("--long-wave", type=int, default=2368, help="pairs of the one-resident-wave figure of the same leg (0 = skip)")
    ap.add_argument("--strong-pairs", type=int, default=100_000_000, help="configs[2]: total pairs of the strong-scaling leg (0 = skip)")
    ap.add_argument("--strong-slice", type=int, default=5_000_000)
    ap.add_argument("--strong-check", type=int, default=1_000_000, help="pairs per rank of that leg compared with the CPU oracle")
    ap.add_argument("--bgzf-reads", type=int, default=2_000_000, help="reads of the FASTQ(BGZF) ingest leg (0 = skip)")
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
    from mini_parallel_b200 import sharding, synth
    first_pair, _ = sharding.shard_range(rank, world, world * n)   # this rank's contiguous shard of the counter-RNG stream
    eng = mp.Engine(local_rank)
    lib = mp.load_library()
    if args.variant >= 0:
        eng.set_short_variant(args.variant)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return sharding.reduce_scalars([x], "max")[0]

    # ---- the workload: configs[1] shape, the windows cut from a synthetic reference (they overlap like candidate
    #      windows of a genome do); device-resident ASCII reads + materialised ASCII windows, as a FASTQ chunk arrives ----
    numa_node = lib.swb_numa_prefer_device(local_rank)             # pinned buffers on the GPU's own NUMA node (a preference)
    h_ref = torch.from_numpy(synth.synth_reference(args.ref_bases)).pin_memory()
    d_ref = h_ref.to(dev)
    d_q = torch.empty(n * rl, dtype=torch.uint8, device=dev)
    d_r = torch.empty(n * wl, dtype=torch.uint8, device=dev)
    d_qo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_ro = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_ws = torch.empty(n, dtype=torch.int64, device=dev)
    d_out = torch.empty(n * 3, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def gen_workload(dist_):
        eng.synth_device_ref(d_ref.data_ptr(), args.ref_bases, first_pair, n, rl, wl, dist_, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(),
                             d_ro.data_ptr(), d_ws.data_ptr())
        eng.sync()
    gen_workload(args.dist)

    def step_device():
        eng.score_batch_device(d_q.data_ptr(), d_qo.data_ptr(), n * rl, d_r.data_ptr(), d_ro.data_ptr(), n * wl, n, rl, wl,
                               d_out.data_ptr())

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

    # ---- host copies of the same pairs (pinned) ----
    h_q = torch.empty(n * rl, dtype=torch.uint8).pin_memory()
    h_r = torch.empty(n * wl, dtype=torch.uint8).pin_memory()
    h_qo = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    h_ro = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    h_ws = torch.empty(n, dtype=torch.int64).pin_memory()
    h_wl = torch.full((n,), wl, dtype=torch.int32).pin_memory()
    h_out = torch.empty(n * 3, dtype=torch.int32).pin_memory()
    h_q.copy_(d_q); h_r.copy_(d_r); h_qo.copy_(d_qo); h_ro.copy_(d_ro); h_ws.copy_(d_ws)
    torch.cuda.synchronize()
    lib.swb_numa_reset()
    e2e_steps = max(3, min(args.steps, 10))

    def timed_host(step_fn):
        for _ in range(2):
            step_fn()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            g0.record(stream)
            for _ in range(e2e_steps):
                step_fn()
            g1.record(stream)
        barrier()
        return max_over_ranks(g0.elapsed_time(g1)), eng.last_timings()

    def check(rc):
        if rc != 0:
            raise RuntimeError(lib.swb_last_error().decode())

    # ---- timed region 2 (the headline e2e): everything from HOST memory every step -- the reads, their window coordinates
    #      and the buffer the windows are ranges of; the library uploads the part of the buffer the windows touch once per call ----
    def step_ranges():
        check(lib.swb_score_batch_ranges(eng._h, h_q.data_ptr(), h_qo.data_ptr(), n, h_ref.data_ptr(), args.ref_bases, h_ws.data_ptr(),
                                         h_wl.data_ptr(), h_out.data_ptr()))
    rng_ms, rng_t = timed_host(step_ranges)
    rng_same = bool(torch.equal(h_out.to(dev), d_out))
    rinfo = eng.last_ranges_info()
    e2e_launches = rng_t["kernels"] * e2e_steps

    # ---- timed region 3: the same pairs with every window as its own copy (CSR): 650 MB per step over PCIe ----
    def step_csr():
        check(lib.swb_score_batch(eng._h, h_q.data_ptr(), h_qo.data_ptr(), h_r.data_ptr(), h_ro.data_ptr(), n, h_out.data_ptr()))
    csr_ms, csr_t = timed_host(step_csr)
    csr_same = bool(torch.equal(h_out.to(dev), d_out))

    # ---- timed region 4: reads from the host against a reference uploaded ONCE (outside the timed region, like a genome) ----
    check(lib.swb_set_reference(eng._h, h_ref.data_ptr(), args.ref_bases))
    def step_ref():
        check(lib.swb_score_batch_vs_reference(eng._h, h_q.data_ptr(), h_qo.data_ptr(), n, h_ws.data_ptr(), h_wl.data_ptr(), h_out.data_ptr()))
    ref_ms, _ = timed_host(step_ref)
    ref_same = bool(torch.equal(h_out.to(dev), d_out))

    # ---- cpu_baseline inputs: the same pairs rank 0 scores per step (copied before the buffers are reused) ----
    cpu_inputs_host = None
    if rank == 0:
        cpu_inputs_host = (h_q.numpy().copy(), h_qo.numpy().astype(np.uint64), h_r.numpy().copy(), h_ro.numpy().astype(np.uint64))
        gpu_first = np.stack([h_out.numpy().reshape(n, 3)[:, k].copy() for k in range(3)], axis=1)
    del h_r, h_ro

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
        gen_workload(1)
        timed("unrelated_reads", n, rl, wl, d_q, d_qo, d_r, d_ro, f"{n} pairs {rl}x{wl}, reads independent of their windows")
        # 1 % of the reads carry one 'N': byte-compare routing (smith_waterman.cl:114 compares raw bytes)
        gen_workload(0)
        g = torch.Generator(device=dev); g.manual_seed(0xB200)
        sel = torch.randperm(n, device=dev, generator=g)[: n // 100].to(torch.int64)
        d_q[sel * rl + torch.randint(0, rl, (sel.numel(),), device=dev, generator=g)] = ord("N")
        torch.cuda.synchronize()
        timed("one_percent_N", n, rl, wl, d_q, d_qo, d_r, d_ro, f"{n} pairs {rl}x{wl}, related reads, 1 % of the reads contain one N")
        # the SURVEY.md 8d generator proper: every pair its own iid window (no reference); round 1's headline workload
        eng.synth_device(first_pair, n, rl, wl, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        timed("iid_windows", n, rl, wl, d_q, d_qo, d_r, d_ro, f"{n} pairs {rl}x{wl}, related reads, every window iid (SURVEY.md 8d generator, the round-1 workload)")
        # BASELINE.json configs[0] shape: 10 k reads against 1 kb windows
        c1 = 10_000
        eng.synth_device(0, c1, rl, 1000, 0, d_q.data_ptr(), d_qo.data_ptr(), d_r.data_ptr(), d_ro.data_ptr())
        timed("config0_shape", c1, rl, 1000, d_q, d_qo, d_r, d_ro, "BASELINE.json configs[0] shape: 10000 reads of 150 bp x 1 kb windows (one small launch)")
        eng.sync()
    del d_q, d_r, d_qo, d_ro, d_ws, d_out, d_ref
    torch.cuda.empty_cache()

    # ---- BASELINE.json configs[2]: 100 M reads, strong scaling (all ranks) ----
    strong = leg_strong(args, eng, dev, rank, world, barrier, sharding.reduce_scalars) if args.strong_pairs > 0 else None
    # ---- BASELINE.json configs[3]: 10 000 pairs 10 kb x 10 kb (rank 0) ----
    aux_long = leg_long(args, eng, dev) if rank == 0 and (args.long_pairs > 0 or args.long_wave > 0) else None
    # ---- FASTQ(BGZF) ingest on the GPU (rank 0) ----
    aux_bgzf = leg_bgzf(args, eng, lib, mp) if rank == 0 and args.bgzf_reads > 0 else None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    uniform_offsets = os.environ.get("SWB_UNIFORM_OFFSETS", "1") != "0"
    cells_step = float(n) * rl * wl
    gcups = world * cells_step * args.steps / (ms_total * 1e-3) / 1e9
    peaks = load_peaks()
    rate, rate_src = load_issue_rate()
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_gcups = sms * peaks["sm_max_mhz"] * 1e6 * rate / INT_ISSUE_PER_CELL / 1e9
    traffic, traffic_src = load_ncu_traffic("stream_kernel")
    pack_traffic, pack_traffic_src = load_ncu_traffic("pack2bit_kernel")
    k_ms = statistics.mean(short_ms)
    k_gcups = cells_step / (k_ms * 1e-3) / 1e9
    p_ms = statistics.mean(pack_ms)
    pack_bytes = 1.25 * n * (rl + wl)                      # 1 B read + 0.25 B written per base

    def e2e_obj(ms, same, h2d, api, extra=None):
        d = {"value": round(world * cells_step * e2e_steps / (ms * 1e-3) / 1e9, 2), "unit": "GCUPS",
             "reads_per_s": round(world * n * e2e_steps / (ms * 1e-3), 1), "steps": e2e_steps, "ms_per_step": round(ms / e2e_steps, 3),
             "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(4 * h_out.numel()), "equals_device_resident_results": same, "api": api}
        d.update(extra or {})
        return d
    off_bytes = 0 if uniform_offsets else 8 * (n + 1)
    # ---- CPU baseline: the same pairs, all host cores, a few full passes ----
    import oracle_lib as ol
    threads = os.cpu_count() or 1
    cpu_times = cpu_simd_time(cpu_inputs_host, threads, max(1, args.cpu_passes), warmup=1)
    cpu_val = n * rl * wl / statistics.mean(cpu_times) / 1e9
    exp = ol.batch(*cpu_inputs_host, threads=threads, simd=True)
    cpu_equal = bool(np.array_equal(gpu_first, np.stack([exp["score"], exp["end_i"], exp["end_j"]], axis=1)))
    t0 = time.perf_counter()
    ol.batch(cpu_inputs_host[0][: 2000 * rl], cpu_inputs_host[1][:2001], cpu_inputs_host[2][: 2000 * wl], cpu_inputs_host[3][:2001], threads=1, simd=False)
    scalar_gcups = 2000 * rl * wl / (time.perf_counter() - t0) / 1e9
    if aux_long:
        for v in aux_long.values():
            v["peak_gcups_at_4_instr_per_cell"] = round(sms * peaks["sm_max_mhz"] * 1e6 * rate / 4.0 / 1e9, 1)
            v["frac_of_that_peak"] = round(v["gcups"] / v["peak_gcups_at_4_instr_per_cell"], 4)

    cfg = workload_config(args, world, n)                  # identical in both arms
    cfg_detail = {"sharding": f"{world} x independent shard of the same counter-RNG stream (weak scaling)",
                  "inputs": "`value` and e2e_csr_windows: ASCII reads + every window materialised as ASCII (650 MB per step); `e2e`: the windows as ranges of the host reference buffer",
                  "l2_policy": f"inputs larger than L2 ({(n * (rl + wl)) >> 20} MiB ASCII per step vs 126 MB L2)",
                  "short_variant": args.variant, "routing": routing}
    line = {
        "metric": "GCUPS", "value": round(gcups, 2), "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "reads_per_s": round(world * n * args.steps / (ms_total * 1e-3), 1),
        "config": cfg, "config_detail": cfg_detail,
        "gpu_launches": launches,
        "gpu_launches_note": f"kernels of the {args.steps} timed device-resident steps; the e2e region launched {e2e_launches} more in {e2e_steps} steps",
        "clocks": clocks,
        "e2e": e2e_obj(rng_ms, rng_same, n * rl + off_bytes + rinfo["bytes_uploaded"] + 8 * n,
                       "swb_score_batch_ranges: reads, window coordinates and the buffer the windows are ranges of all start in pinned HOST memory "
                       "every step; the part of the buffer the windows touch crosses PCIe once per call, reads pipelined in chunks over 3 streams",
                       {"path": f"shared window buffer: {rinfo['window_bytes']} window bytes are ranges of {rinfo['bytes_uploaded']} uploaded bytes "
                                f"({rinfo['window_bytes'] / max(rinfo['bytes_uploaded'], 1):.1f}x overlap)",
                        "pinned_numa_node": numa_node, "stage_ms_sum_over_chunks": {k: round(v, 3) for k, v in rng_t.items() if k.endswith("_ms")}}),
        "e2e_csr_windows": e2e_obj(csr_ms, csr_same, n * (rl + wl) + 2 * off_bytes,
                                   "swb_score_batch: the same pairs with every window as its own ASCII copy (CSR), 650 MB per step over PCIe -- "
                                   "bound by the host-to-device copy, see profiles/h2d_ceiling_*",
                                   {"stage_ms_sum_over_chunks": {k: round(v, 3) for k, v in csr_t.items() if k.endswith("_ms")}}),
        "e2e_resident_reference": e2e_obj(ref_ms, ref_same, n * rl + off_bytes + 8 * n,
                                          "swb_score_batch_vs_reference: reads + window starts from pinned host memory, the reference uploaded once "
                                          "before the timed region"),
        "roofline": {"bound": "int_issue", "kernel": "sw_stream_kernel", "achieved": round(k_gcups, 1), "peak": round(peak_gcups, 1),
                     "unit": "GCUPS", "frac": round(k_gcups / peak_gcups, 4),
                     "traffic": None if traffic is None else int(traffic),       # dram read+write bytes of one launch (ncu --set full)
                     "traffic_source": None if traffic is None else f"profiles/{traffic_src}",
                     "algorithmic_bytes_per_launch": int(n * (rl / 4 + wl / 4 + 32 + 12)),
                     "kernel_ms": round(k_ms, 4), "kernel_share_of_step": round(k_ms / (ms_total / args.steps), 4),
                     "peak_is": f"{sms} SMs x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} MEASURED_PEAKS.json) x {rate:.2f} DPX s16x2 "
                                f"thread-instr/clk/SM (measured, {rate_src}) / {INT_ISSUE_PER_CELL} instr per cell"},
        "roofline_pack": {"bound": "hbm", "kernel": "pack2bit_kernel x2 + classify_kernel", "achieved": round(pack_bytes / (p_ms * 1e-3) / 1e9, 1),
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(pack_bytes / (p_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                          "kernel_ms": round(p_ms, 4), "traffic": None if pack_traffic is None else int(pack_traffic),
                          "traffic_source": None if pack_traffic is None else f"profiles/{pack_traffic_src} (pack2bit_kernel on the window array)",
                          "peak_is": f"{peaks['source']} copy bandwidth",
                          "pack2bit_kernel_alone": {"bytes": int(1.25 * n * wl), "ms": round(pack_alone_ms, 4),
                                                    "achieved": round(1.25 * n * wl / (pack_alone_ms * 1e-3) / 1e9, 1),
                                                    "frac": round(1.25 * n * wl / (pack_alone_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                                                    "what": "the window array of the step (500 MB in, 125 MB out), 10 back-to-back launches"}},
        "config2_strong": strong,
        "aux_long_pairs": aux_long,
        "aux_bgzf_ingest": aux_bgzf,
        "aux": aux,
        "cpu_baseline": {"value": round(cpu_val, 3), "unit": "GCUPS", "cores": threads, "kind": "port", "isa": ol.simd_isa(),
                         "scalar_oracle_gcups_1_core": round(scalar_gcups, 3), "equals_gpu_results_on_all_pairs": cpu_equal,
                         "sample": f"the {n} pairs rank 0 scores per step, {len(cpu_times)} full passes of {min(cpu_times):.3f}-{max(cpu_times):.3f} s on "
                                   f"{threads} threads (oracle/sw_simd.c, shared work cursor)"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
