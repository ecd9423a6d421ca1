#!/usr/bin/env python
"""bench_host_fastq.py -- the HOST FASTQ path on one plain .gz file, no GPU: count_bases_in_fastq (= --test-wgs,
aligner.rs:535-544: inflate + line splitting) with the file's inflate on 0 (the parsing thread itself), 1 (one thread ahead,
hgz::AsyncGunzip) or n >= 3 threads (hgz::ParallelGunzip, csrc/host_pgunzip.h), and with zlib's gzread for scale.

    python tools/bench_host_fastq.py [--reads 1000000] [--level 1] [--quals noisy|const] [--threads 0,1,3,5,7] [file.fastq.gz]

Without a file a synthetic one is made with tools/bench_wgs.py's generator (150 bp reads, 316 bytes per record)."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file", nargs="?")
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--level", type=int, default=1)
    ap.add_argument("--quals", default="noisy", choices=["noisy", "const"])
    ap.add_argument("--threads", default="0,1,3,5,7")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    from mini_parallel_b200 import aligner
    os.environ.setdefault("GPU_CHUNK_SIZE_READS", "100000")
    path, made = args.file, None
    if not path:
        import bench_wgs
        made = tempfile.mkdtemp(prefix="swb_hostfq_")
        path = os.path.join(made, "SYN_L001_R1_001.fastq.gz")
        bench_wgs.make_file((path, 0, args.reads, 16_000_000, 150, 500, args.level, False, args.quals == "noisy"))
    out = {"file": os.path.basename(path), "gz_mb": round(os.path.getsize(path) / 1e6, 1), "host_cores": os.cpu_count(), "rows": []}
    def run(tag):
        best, bases = 1e30, 0
        for _ in range(args.reps):
            t0 = time.perf_counter(); bases = aligner.count_bases_in_fastq(path); best = min(best, time.perf_counter() - t0)
        out["rows"].append({"inflate": tag, "seconds": round(best, 3), "bases": bases, "m_bases_per_s": round(bases / best / 1e6, 1)})
    os.environ["SWB_HOST_INFLATE"] = "zlib"
    run("zlib gzread on the parsing thread")
    del os.environ["SWB_HOST_INFLATE"]
    for n in [int(x) for x in args.threads.split(",")]:
        os.environ["SWB_INFLATE_THREADS"] = str(n)
        run({0: "own decoder on the parsing thread"}.get(n, f"own decoder, {n} thread(s) beside the parsing thread" + (" (ParallelGunzip)" if n >= 3 else " (AsyncGunzip)")))
    print(json.dumps(out))
    if made:
        os.unlink(path); os.rmdir(made)


if __name__ == "__main__":
    main()
