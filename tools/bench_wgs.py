#!/usr/bin/env python
"""BASELINE.json configs[4], scaled: synthetic WGS lanes ({SAMPLE}_L{lane:03}_R{read}_001.fastq.gz, 150 bp reads)
streamed end to end through build/rustseq_mini --full-wgs --gpu (inflate + parse + H2D + pack + score + D2H).
Generates the files (reads cut from the driver's synthetic reference at the window each read is paired with,
1 % substitutions), runs the CLI, prints one JSON line with wall-clock reads/s and GCUPS."""
import argparse
import json
import os
import subprocess
import sys
import time
import zlib
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synth_reference(n):
    """Same bytes as load_reference() in rustseq_host.cpp when WGS_REFERENCE is unset."""
    k = np.arange((n + 31) // 32, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = splitmix64(np.uint64(0xB2F0) + k)
    sh = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    codes = ((x[:, None] >> sh) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]


def make_file(args):
    """One FASTQ file (or, with a 10th / 11th field, the slice [first_read, first_read + n_reads) of one: BGZF files are
    concatenations of independent blocks, so the slices of a file are made by different processes and joined)."""
    path, fi, n_reads, ref_len, rl, wl, level, blocked, noisy = args[:9]
    first_read = args[9] if len(args) > 9 else 0
    write_eof = args[10] if len(args) > 10 else True
    ref = synth_reference(ref_len)
    if blocked:                                    # BGZF: independent <= 64 KiB members with a 'BC' size field (bgzip / BCL Convert)
        sys.path.insert(0, ROOT)
        from mini_parallel_b200 import bgzf
    co = None if blocked else zlib.compressobj(level, zlib.DEFLATED, 31)
    step = 100_000
    with open(path, "wb") as f:
        for a in range(first_read, first_read + n_reads, step):
            m = min(step, first_read + n_reads - a)
            k = np.arange(a, a + m, dtype=np.uint64)
            with np.errstate(over="ignore"):
                g = (np.uint64(fi) << np.uint64(40)) + k
                ws = splitmix64(g ^ np.uint64(0xB202)) % np.uint64(ref_len - wl + 1)
                off = splitmix64(g ^ np.uint64(0xB203)) % np.uint64(wl - rl + 1)
                noise = splitmix64((g[:, None] * np.uint64(257) + np.arange(rl, dtype=np.uint64)[None, :]) ^ np.uint64(0xB204))
            idx = (ws + off)[:, None].astype(np.int64) + np.arange(rl, dtype=np.int64)[None, :]
            reads = ref[idx]
            sub = (noise % np.uint64(100)) == 0
            alt = np.frombuffer(b"ACGT", dtype=np.uint8)[((noise >> np.uint64(8)) & np.uint64(3)).astype(np.uint8)]
            reads = np.where(sub, alt, reads)
            rec = np.empty((m, 12 + rl + 3 + rl + 1), dtype=np.uint8)
            names = np.char.add("@", np.char.zfill(k.astype(str), 10)).astype("S11")
            rec[:, :11] = np.frombuffer(names.tobytes(), dtype=np.uint8).reshape(m, 11)
            rec[:, 11] = 10
            rec[:, 12:12 + rl] = reads
            rec[:, 12 + rl:12 + rl + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
            if noisy:                                  # four binned quality values (NovaSeq-like), 62 / 25 / 9 / 3 %
                qn = (noise >> np.uint64(16)) & np.uint64(255)
                rec[:, 12 + rl + 3:12 + 2 * rl + 3] = np.where(qn < 160, ord("F"), np.where(qn < 224, ord(":"), np.where(qn < 248, ord(","), ord("#"))))
            else:
                rec[:, 12 + rl + 3:12 + 2 * rl + 3] = ord("I")
            rec[:, -1] = 10
            if blocked:
                f.write(bgzf.compress(rec.tobytes(), level, 65280, eof=False))     # blocks may end anywhere inside a record
            else:
                f.write(co.compress(rec.tobytes()))
        if write_eof:
            f.write(bgzf.EOF_BLOCK if blocked else co.flush())
    return os.path.getsize(path), n_reads * rec.shape[1]


def io_ceiling(paths, gz_bytes):
    """What the host can deliver at all: threads that do nothing but pread() the compressed files in 112 MiB segments into
    their own buffers (no GPU, no parsing), 1 / 2 / 4 threads per file, segments dealt round-robin like the driver's readers."""
    import threading
    seg = 112 << 20
    out = {"what": "pread() of the compressed files only, 112 MiB segments, MB/s over all files", "files": len(paths), "gz_mb": round(gz_bytes / 1e6, 1),
           "host_cores": os.cpu_count()}

    def reader(path, r, nr, buf):
        fd = os.open(path, os.O_RDONLY)
        size = os.fstat(fd).st_size
        k = r
        while k * seg < size:
            got, want = 0, min(seg, size - k * seg)
            mv = memoryview(buf)
            while got < want:
                n = os.preadv(fd, [mv[got:want]], k * seg + got)
                if n <= 0:
                    break
                got += n
            k += nr
        os.close(fd)
    for nr in (1, 2, 4, 1):
        bufs = [bytearray(min(seg, os.path.getsize(p))) for p in paths for r in range(nr)]      # allocated (and touched) outside the timed region
        th = [threading.Thread(target=reader, args=(p, r, nr, bufs[i * nr + r])) for i, p in enumerate(paths) for r in range(nr)]
        t0 = time.time()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.time() - t0
        key = f"threads_per_file_{nr}" + ("_again" if f"threads_per_file_{nr}" in out else "")
        out[key] = {"seconds": round(dt, 3), "mb_per_s": round(gz_bytes / dt / 1e6, 1)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads-per-file", type=int, default=250_000)
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--chunk-reads", type=int, default=100_000)
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--ref-bases", type=int, default=16_000_000)
    ap.add_argument("--dir", default="/tmp/synwgs")
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--reuse", action="store_true", help="keep the files already in --dir (same parameters) instead of regenerating them")
    ap.add_argument("--bgzf", action="store_true", help="write blocked gzip (BGZF): the driver inflates and parses it on the GPU")
    ap.add_argument("--clone-files", action="store_true",
                    help="generate the first file only and copy it to the other 15 names (the driver pairs every read with a window by "
                         "(file index, read index), so the files still score differently): a full-size data set in a fraction of the time")
    ap.add_argument("--io-ceiling", action="store_true",
                    help="do not score: only pread() the files with 1, 2 and 4 threads per file (the host-side ceiling of the ingest)")
    ap.add_argument("--readers", type=int, default=0, help="SWB_READERS_PER_FILE for the run (0: the driver's default)")
    ap.add_argument("--quals", default="constant", choices=["constant", "noisy"], help="quality strings: all 'I', or four binned values at random")
    args = ap.parse_args()
    os.makedirs(args.dir, exist_ok=True)
    jobs = []
    fi = 0
    for lane in range(1, args.lanes + 1):
        for rd in (1, 2):
            jobs.append((os.path.join(args.dir, f"SYN_L{lane:03d}_R{rd}_001.fastq.gz"), fi, args.reads_per_file, args.ref_bases, 150, 500, args.level, args.bgzf, args.quals == "noisy"))
            fi += 1
    t0 = time.time()
    rec_len = 12 + 150 + 3 + 150 + 1
    all_jobs = jobs
    if args.clone_files and not (args.reuse and all(os.path.exists(j[0]) for j in jobs)):
        jobs = jobs[:1]
    if args.reuse and all(os.path.exists(j[0]) for j in jobs):     # files of an earlier run with the same parameters
        sizes = [(os.path.getsize(j[0]), args.reads_per_file * rec_len) for j in jobs]
    elif args.bgzf and (os.cpu_count() or 1) >= 2 * len(jobs):     # more cores than files: every file in slices, joined afterwards
        parts = max(1, (os.cpu_count() or 1) // len(jobs))
        per = (args.reads_per_file + parts - 1) // parts
        pjobs = []
        for j in jobs:
            for k in range(parts):
                a, b = k * per, min(args.reads_per_file, (k + 1) * per)
                if b > a:
                    pjobs.append((j[0] + f".part{k}",) + j[1:2] + (b - a,) + j[3:] + (a, False))
        with ProcessPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            list(ex.map(make_file, pjobs))
        sys.path.insert(0, ROOT)
        from mini_parallel_b200 import bgzf as _bgzf
        for j in jobs:
            with open(j[0], "wb") as out:
                for pj in pjobs:
                    if pj[0].startswith(j[0] + ".part"):
                        with open(pj[0], "rb") as src:
                            while True:
                                blk = src.read(64 << 20)
                                if not blk:
                                    break
                                out.write(blk)
                        os.unlink(pj[0])
                out.write(_bgzf.EOF_BLOCK)
        sizes = [(os.path.getsize(j[0]), args.reads_per_file * rec_len) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            sizes = list(ex.map(make_file, jobs))
    if len(all_jobs) > len(jobs):                                   # --clone-files
        import shutil
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=8) as ex:
            list(ex.map(lambda j: shutil.copyfile(jobs[0][0], j[0]), all_jobs[1:]))
        sizes = [sizes[0]] * len(all_jobs)
    jobs = all_jobs
    gen_s = time.time() - t0
    gz_bytes = sum(s[0] for s in sizes); text_bytes = sum(s[1] for s in sizes)
    if args.io_ceiling:
        print(json.dumps(io_ceiling([j[0] for j in jobs], gz_bytes)))
        return
    env = dict(os.environ, GPU_CHUNK_SIZE_READS=str(args.chunk_reads), WGS_DATA_DIR=args.dir, WGS_SAMPLE_ID="SYN", WGS_LANES=str(args.lanes),
               WGS_READS_PER_LANE="2", WGS_SYNTH_REFERENCE_BASES=str(args.ref_bases), SWB_NUM_DEVICES=str(args.devices), WGS_CHECKPOINT_DIR=args.dir)
    env.pop("SWB_GPU_ALIGN_MODE", None)
    if args.readers:
        env["SWB_READERS_PER_FILE"] = str(args.readers)
    cli = os.path.join(ROOT, "build", "rustseq_mini")
    subprocess.run([cli, "-1", "ACGT", "-2", "ACGT", "--gpu"], env=env, capture_output=True)      # warm the driver / context creation
    t0 = time.time()
    r = subprocess.run([cli, "--full-wgs", "--gpu"], env=env, capture_output=True, text=True)
    wall = time.time() - t0
    if os.environ.get("SWB_DEBUG") or os.environ.get("SWB_STAMPS"):
        print("\n".join(l for l in r.stderr.splitlines() if l.startswith("[main]") or l.startswith("[wgs]")), file=sys.stderr)
        print(f"[bench_wgs] subprocess wall {wall:.3f} s", file=sys.stderr)
    if os.environ.get("SWB_DEBUG"):
        print(r.stderr[-1500:], file=sys.stderr)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:], file=sys.stderr)
        raise SystemExit("rustseq_mini --full-wgs failed")
    scores = [int(l.split("Score=")[1].split(",")[0]) for l in r.stdout.splitlines() if "complete: Score=" in l]
    file_s = [float(l.split("Time:")[1].split("s")[0]) for l in r.stdout.splitlines() if "complete: Score=" in l]
    n_reads = args.reads_per_file * len(jobs)
    print(json.dumps({
        "workload": f"BASELINE.json configs[4] scaled: {len(jobs)} files x {args.reads_per_file} reads of 150 bp ({'BGZF blocked gzip' if args.bgzf else 'gzip'} -{args.level}, {args.quals} qualities), each read vs a 500 bp window "
                    f"of a {args.ref_bases} bp device-resident reference, GPU_CHUNK_SIZE_READS={args.chunk_reads}, {args.devices} GPU(s)",
        "wall_s": round(wall, 3), "reads_per_s": round(n_reads / wall, 1), "gcups_end_to_end": round(n_reads * 150 * 500 / wall / 1e9, 1),
        "gz_mb_per_s": round(gz_bytes / wall / 1e6, 1), "fastq_text_mb_per_s": round(text_bytes / wall / 1e6, 1),
        "slowest_file_s": max(file_s) if file_s else None,
        "pipeline_reads_per_s": round(n_reads / max(file_s), 1) if file_s else None,      # all files run concurrently: excludes process + CUDA context start-up
        "pipeline_gcups": round(n_reads * 150 * 500 / max(file_s) / 1e9, 1) if file_s else None,
        "startup_s": round(wall - max(file_s), 3) if file_s else None,                    # process start, CUDA contexts, reference upload: everything before / after the files
        "readers_per_file": args.readers or "default (cores / files, 1..4)",
        "host_cores": os.cpu_count(), "generate_s": round(gen_s, 1),
        "mean_score_per_read": round(sum(scores) / n_reads, 2), "files_done": len(scores)}))


if __name__ == "__main__":
    main()
