"""GPU tests of the host mirror: the reference's entry points (gpu_align & callers, the WGS driver, the CLI)
driven end to end on a B200, results checked against the oracle."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from mini_parallel_b200 import aligner
from mini_parallel_b200.engine import to_csr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "build", "rustseq_mini")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _fastq(path, reads):
    with gzip.open(path, "wb") as f:
        for k, r in enumerate(reads):
            f.write(b"@r%d\n%s\n+\n%s\n" % (k, r, b"I" * len(r)))


@pytest.fixture
def device():
    devs = aligner.get_gpu_devices()
    assert devs and aligner.is_gpu_available()
    return devs[0]


def test_gpu_align_sw_and_compat_modes(device, monkeypatch):
    rng = np.random.default_rng(1)
    monkeypatch.delenv("SWB_GPU_ALIGN_MODE", raising=False)
    assert device.max_work_group_size == 1024 and device.memory_gb > 100
    for _ in range(20):
        a = bytes(ACGT[rng.integers(0, 4, int(rng.integers(1, 400)))])
        b = bytes(ACGT[rng.integers(0, 4, int(rng.integers(1, 900)))])
        assert aligner.gpu_align_ex(a, b, device) == ol.sw_linear(a, b)
        assert aligner.gpu_align(a, b, device) == ol.sw_linear(a, b)[0]
    assert aligner.gpu_align(b"", b"ACGT", device) == 0                               # aligner.rs:413-416
    monkeypatch.setenv("SWB_GPU_ALIGN_MODE", "ref_compat")
    for _ in range(20):
        a = bytes(ACGT[rng.integers(0, 4, int(rng.integers(1, 5000)))])
        b = bytes(ACGT[rng.integers(0, 4, int(rng.integers(1, 5000)))])
        assert aligner.gpu_align(a, b, device) == ol.ref_compat_align(a, b, 1024)
    chunk = bytes(ACGT[rng.integers(0, 4, 1500)])
    assert aligner.gpu_align_chunk_self(chunk, device) == 2                            # what --full-wgs sums today
    assert aligner.gpu_align_chunk_self(chunk[:999], device) == 0                      # aligner.rs:366-368
    monkeypatch.delenv("SWB_GPU_ALIGN_MODE")
    assert aligner.gpu_align_chunk_self(chunk, device) == 3000                         # true SW of a sequence with itself
    monkeypatch.setenv("SWB_MAX_CELLS", "1000")
    with pytest.raises(aligner.AlignerError, match="Sequence too large"):
        aligner.gpu_align(chunk, chunk, device)


def test_gpu_align_pair_files(tmp_path, device, monkeypatch):
    rng = np.random.default_rng(2)
    r1 = [bytes(ACGT[rng.integers(0, 4, 150)]) for _ in range(57)]
    r2 = [bytes(ACGT[rng.integers(0, 4, 150)]) for _ in range(57)]
    _fastq(tmp_path / "a.fastq.gz", r1)
    _fastq(tmp_path / "b.fastq.gz", r2)
    monkeypatch.setenv("GPU_CHUNK_SIZE_READS", "10")
    monkeypatch.delenv("SWB_GPU_ALIGN_MODE", raising=False)
    res = aligner.gpu_align_pair(tmp_path / "a.fastq.gz", tmp_path / "b.fastq.gz", device)
    assert res.score64 == sum(ol.sw_linear(a, b)[0] for a, b in zip(r1, r2))
    assert res.total_reads == 57 and res.total_bases == 2 * 57 * 150
    assert res.gpu_device.decode().startswith("NVIDIA")
    monkeypatch.setenv("SWB_GPU_ALIGN_MODE", "ref_compat")                              # aligner.rs:390-398: chunks x chunks
    res = aligner.gpu_align_pair(tmp_path / "a.fastq.gz", tmp_path / "b.fastq.gz", device)
    c1 = [b"".join(r1[k:k + 10]) for k in range(0, 57, 10)]
    c2 = [b"".join(r2[k:k + 10]) for k in range(0, 57, 10)]
    assert res.score64 == sum(ol.ref_compat_align(x, y, 1024) for x in c1 for y in c2)


def _make_lanes(tmp_path, lanes, reads_per_file, rng):
    files = {}
    for lane in range(1, lanes + 1):
        for rd in (1, 2):
            reads = [bytes(ACGT[rng.integers(0, 4, 150)]) for _ in range(reads_per_file)]
            name = f"SYN_L{lane:03d}_R{rd}_001.fastq.gz"
            _fastq(tmp_path / name, reads)
            files[name] = reads
    return files


def test_full_wgs_driver_both_modes(tmp_path, device, monkeypatch):
    monkeypatch.setenv("WGS_CHECKPOINT_DIR", str(tmp_path))
    rng = np.random.default_rng(3)
    files = _make_lanes(tmp_path, 2, 45, rng)
    for k, v in dict(WGS_DATA_DIR=str(tmp_path), WGS_SAMPLE_ID="SYN", WGS_LANES="2", WGS_READS_PER_LANE="2",
                     GPU_CHUNK_SIZE_READS="20", WGS_SYNTH_REFERENCE_BASES="200000", WGS_WINDOW_LEN="500").items():
        monkeypatch.setenv(k, v)
    # --- reference-compatible mode: per chunk concat + self compat-align => 2 per chunk of >= 1000 bases ---
    monkeypatch.setenv("SWB_GPU_ALIGN_MODE", "ref_compat")
    res = aligner.process_full_wgs_dataset(device)
    assert len(res) == 4
    for r in res:
        assert r.score64 == 2 * 2 and r.total_reads == 45 and r.total_bases == 45 * 150   # chunks of 20,20,5 reads; the 750-base tail is skipped (aligner.rs:366)
    # --- Smith-Waterman mode: every read against its window of the (synthetic) resident reference ---
    monkeypatch.delenv("SWB_GPU_ALIGN_MODE")
    res = aligner.process_full_wgs_dataset(device)
    assert len(res) == 4

    def splitmix(x):
        m = (1 << 64) - 1
        x = (x + 0x9E3779B97F4A7C15) & m
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
        return x ^ (x >> 31)
    n_ref = 200000
    ref = bytearray(n_ref)
    for k in range(0, n_ref, 32):
        x = splitmix(0xB2F0 + (k >> 5))
        for j in range(k, min(n_ref, k + 32)):
            ref[j] = b"ACGT"[x & 3]
            x >>= 2
    names = sorted(files)                                   # L001_R1, L001_R2, L002_R1, L002_R2 = the driver's order
    for fi, name in enumerate(names):
        exp = 0
        for k, read in enumerate(files[name]):
            start = splitmix(((fi << 40) + k) ^ 0xB202) % (n_ref - 500 + 1)
            exp += ol.sw_linear(read, bytes(ref[start:start + 500]))[0]
        assert res[fi].score64 == exp, name
        assert res[fi].total_reads == 45


def test_full_wgs_pipeline_pieces_and_caps(tmp_path, device, monkeypatch):
    """The threaded --full-wgs pipeline: chunks larger than one pipeline piece (16384 reads), a GPU_CHUNK_SIZE_BASES cap,
    ragged read lengths, one file per reader thread; per-file score totals against the SIMD oracle."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_wgs
    monkeypatch.setenv("WGS_CHECKPOINT_DIR", str(tmp_path))
    rng = np.random.default_rng(33)
    n_ref, n_reads = 300_000, 40_000
    ref = bench_wgs.synth_reference(n_ref)
    files = {}
    for fi, (lane, rd) in enumerate(((1, 1), (1, 2))):
        k = np.arange(n_reads, dtype=np.uint64)
        with np.errstate(over="ignore"):
            ws = bench_wgs.splitmix64(((np.uint64(fi) << np.uint64(40)) + k) ^ np.uint64(0xB202)) % np.uint64(n_ref - 500 + 1)
        lens = rng.integers(30, 161, n_reads)
        offs = rng.integers(0, 300, n_reads)
        reads = []
        for j in range(n_reads):
            a = int(ws[j]) + int(offs[j])
            r = ref[a:a + int(lens[j])].copy()
            m = rng.random(r.size) < 0.02
            r[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
            reads.append(r.tobytes())
        name = f"SYN_L{lane:03d}_R{rd}_001.fastq.gz"
        with gzip.open(tmp_path / name, "wb", compresslevel=1) as f:
            f.write(b"".join(b"@r%d\n%s\n+\n%s\n" % (j, r, b"I" * len(r)) for j, r in enumerate(reads)))
        files[name] = (reads, ws)
    for kk, v in dict(WGS_DATA_DIR=str(tmp_path), WGS_SAMPLE_ID="SYN", WGS_LANES="1", WGS_READS_PER_LANE="2",
                      WGS_SYNTH_REFERENCE_BASES=str(n_ref), WGS_WINDOW_LEN="500").items():
        monkeypatch.setenv(kk, v)
    monkeypatch.delenv("SWB_GPU_ALIGN_MODE", raising=False)
    exp = {}
    for name, (reads, ws) in files.items():
        q, qo = to_csr(reads)
        r, ro = to_csr([ref[int(s):int(s) + 500] for s in ws])
        exp[name] = int(ol.batch(q, qo, r, ro, threads=os.cpu_count() or 8, simd=True)["score"].astype(np.int64).sum())
    for chunk_reads, chunk_bases in (("35000", None), ("1000000", "1000000"), ("777", None)):
        monkeypatch.setenv("GPU_CHUNK_SIZE_READS", chunk_reads)
        if chunk_bases:
            monkeypatch.setenv("GPU_CHUNK_SIZE_BASES", chunk_bases)
        else:
            monkeypatch.delenv("GPU_CHUNK_SIZE_BASES", raising=False)
        res = aligner.process_full_wgs_dataset(device)
        assert len(res) == 2
        for fi, name in enumerate(sorted(files)):
            assert res[fi].score64 == exp[name], (name, chunk_reads, chunk_bases)
            assert res[fi].total_reads == n_reads and res[fi].total_bases == sum(len(x) for x in files[name][0])


def test_full_wgs_bgzf_files_take_the_gpu_ingest_path(tmp_path, device, monkeypatch, capfd):
    """The same reads as plain gzip (host zlib + line reader) and as BGZF (inflate + parse on the GPU) give the same
    per-file totals; a corrupt BGZF file falls back to the host reader (which then reports the zlib error)."""
    from mini_parallel_b200 import bgzf
    monkeypatch.setenv("WGS_CHECKPOINT_DIR", str(tmp_path))
    rng = np.random.default_rng(44)
    n_reads = 30_000
    texts = {}
    for lane, rd in ((1, 1), (1, 2)):
        reads = [bytes(ACGT[rng.integers(0, 4, int(rng.integers(20, 161)))]) for _ in range(n_reads)]
        reads[7] = reads[7][:5] + b"N" + reads[7][6:]
        texts[(lane, rd)] = b"".join(b"@r%d\n%s\n+\n%s\n" % (j, r, b"I" * len(r)) for j, r in enumerate(reads))
    for kk, v in dict(WGS_SAMPLE_ID="SYN", WGS_LANES="1", WGS_READS_PER_LANE="2", WGS_SYNTH_REFERENCE_BASES="300000",
                      WGS_WINDOW_LEN="500", GPU_CHUNK_SIZE_READS="7000").items():
        monkeypatch.setenv(kk, v)
    monkeypatch.delenv("SWB_GPU_ALIGN_MODE", raising=False)
    monkeypatch.delenv("GPU_CHUNK_SIZE_BASES", raising=False)
    results = {}
    for fmt in ("gzip", "bgzf"):
        d = tmp_path / fmt
        d.mkdir()
        for (lane, rd), text in texts.items():
            path = d / f"SYN_L{lane:03d}_R{rd}_001.fastq.gz"
            if fmt == "gzip":
                with gzip.open(path, "wb", compresslevel=1) as f:
                    f.write(text)
            else:
                path.write_bytes(bgzf.compress(text, 1, 40_000))
        monkeypatch.setenv("WGS_DATA_DIR", str(d))
        res = aligner.process_full_wgs_dataset(device)
        results[fmt] = [(r.score64, r.total_reads, r.total_bases) for r in res]
        out = capfd.readouterr().out
        assert ("inflate + FASTQ parsing on the GPU" in out) == (fmt == "bgzf")
        assert "Total lines read: %d" % (4 * n_reads) in out
    assert results["gzip"] == results["bgzf"] and results["gzip"][0][1] == n_reads
    # several readers per file (each pread()s its own segments, block headers walked in order), many small segments
    for readers, seg_kb in ((1, 128), (3, 128), (4, 200), (16, 129)):
        monkeypatch.setenv("SWB_READERS_PER_FILE", str(readers)); monkeypatch.setenv("SWB_BGZF_SEGMENT_KB", str(seg_kb))
        res = aligner.process_full_wgs_dataset(device)
        assert [(r.score64, r.total_reads, r.total_bases) for r in res] == results["gzip"], (readers, seg_kb)
        assert "inflate + FASTQ parsing on the GPU" in capfd.readouterr().out
    monkeypatch.delenv("SWB_BGZF_SEGMENT_KB")
    # a block with a damaged payload: the GPU path declines, the host reader takes over and fails like zlib does
    d = tmp_path / "bad"
    d.mkdir()
    for (lane, rd), text in texts.items():
        gz = bytearray(bgzf.compress(text, 1, 40_000))
        if rd == 2:
            blocks, _ = bgzf.walk(bytes(gz))
            for k in range(blocks[3][0] + 5, blocks[3][0] + 60):
                gz[k] ^= 0xA5
        (d / f"SYN_L{lane:03d}_R{rd}_001.fastq.gz").write_bytes(bytes(gz))
    monkeypatch.setenv("WGS_DATA_DIR", str(d))
    for readers in (3, 1):
        monkeypatch.setenv("SWB_READERS_PER_FILE", str(readers))
        with pytest.raises(aligner.AlignerError):
            aligner.process_full_wgs_dataset(device)
        assert "falling back to the host reader" in capfd.readouterr().out
    monkeypatch.delenv("SWB_READERS_PER_FILE")


def test_full_wgs_checkpoint_resume(tmp_path, device, monkeypatch, capfd):
    """A run with WGS_RUN_ID set leaves checkpoint_{run_id}.json; a second run skips the completed files and returns the
    same results; removing one entry makes exactly that file run again (aligner.rs:218-259, working here)."""
    import json
    rng = np.random.default_rng(55)
    _make_lanes(tmp_path, 2, 300, rng)
    for kk, v in dict(WGS_DATA_DIR=str(tmp_path), WGS_SAMPLE_ID="SYN", WGS_LANES="2", WGS_READS_PER_LANE="2", GPU_CHUNK_SIZE_READS="100",
                      WGS_SYNTH_REFERENCE_BASES="200000", WGS_RUN_ID="resume_test", WGS_CHECKPOINT_DIR=str(tmp_path)).items():
        monkeypatch.setenv(kk, v)
    monkeypatch.delenv("SWB_GPU_ALIGN_MODE", raising=False)
    first = [(r.score64, r.total_reads, r.total_bases) for r in aligner.process_full_wgs_dataset(device)]
    out = capfd.readouterr().out
    assert "No existing checkpoint found, starting fresh run" in out and "checkpoint_resume_test.json" in out
    ck = tmp_path / "checkpoint_resume_test.json"
    doc = json.loads(ck.read_text())
    assert doc["completed_files"] == 4 and sorted(f["file_index"] for f in doc["files"]) == [0, 1, 2, 3]
    assert [f["score64"] for f in sorted(doc["files"], key=lambda f: f["file_index"])] == [r[0] for r in first]
    second = [(r.score64, r.total_reads, r.total_bases) for r in aligner.process_full_wgs_dataset(device)]
    out = capfd.readouterr().out
    assert second == first and "Found existing checkpoint: 4 files completed" in out
    assert out.count("Skipping file") == 4 and "Processing file" not in out
    doc["files"] = [f for f in doc["files"] if f["file_index"] != 2]
    doc["completed_files"] = 3
    ck.write_text(json.dumps(doc, indent=2))
    third = [(r.score64, r.total_reads, r.total_bases) for r in aligner.process_full_wgs_dataset(device)]
    out = capfd.readouterr().out
    assert third == first and out.count("Skipping file") == 3 and "Processing file 3/4" in out
    assert json.loads(ck.read_text())["completed_files"] == 4
    # the run report (tools/benchmark.rs:17-42) with measured numbers
    rep = json.loads((tmp_path / "benchmark_results" / "run_resume_test_benchmark_results.json").read_text())
    assert rep["mode"] == "full_wgs" and rep["files_processed"] == 4 and rep["total_reads"] == 1200 and rep["total_bases"] == 1200 * 150
    assert rep["total_score64"] == sum(r[0] for r in first) and rep["throughput_reads_per_second"] > 0
    assert rep["system_info"]["gpu_name"].startswith("NVIDIA") and rep["gpu_memory_used_mb"] > 100 and rep["n_gpus"] >= 1


def test_cli_on_gpu(tmp_path):
    env = {k: v for k, v in os.environ.items() if k != "SWB_GPU_ALIGN_MODE"}
    r = subprocess.run([CLI, "-1", "TGTTACGG", "-2", "GGTTGACTA", "--gpu"], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "GPU acceleration enabled" in r.stdout and "Found GPU: NVIDIA" in r.stdout
    assert "GPU Alignment score: 8" in r.stdout and "End cell: (5, 6)" in r.stdout
    env["SWB_GPU_ALIGN_MODE"] = "ref_compat"
    r = subprocess.run([CLI, "--seq1", "TGTTACGG", "--seq2=GGTTGACTA", "-g"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "GPU Alignment score: 2" in r.stdout                  # what the reference prints today


def test_reads_vs_resident_reference(engine):
    rng = np.random.default_rng(5)
    ref = ACGT[rng.integers(0, 4, 50_000)]
    ref[1000:1010] = ord("N")
    engine.set_reference(ref)
    n = 3000
    starts = rng.integers(0, 50_000 - 600, n).astype(np.uint64)
    lens = rng.integers(300, 600, n).astype(np.uint32)
    reads = []
    for k in range(n):
        w = ref[int(starts[k]):int(starts[k]) + int(lens[k])]
        o = int(rng.integers(0, 150))
        r = w[o:o + 150].copy()
        m = rng.random(r.size) < 0.02
        r[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
        reads.append(r)
    q, qo = to_csr(reads)
    got = engine.score_batch_vs_reference(q, qo, starts, lens)
    wins = [ref[int(s):int(s) + int(l)] for s, l in zip(starts, lens)]
    r, ro = to_csr(wins)
    exp = ol.batch(q, qo, r, ro, threads=8)
    assert np.array_equal(got, exp)
    routing = engine.last_routing()
    assert routing["generic"] > 0 and routing["short"] > 2000                          # windows over the N stretch go generic
