"""CPU tests of the host mirror (csrc/rustseq_host.cpp): the FASTQ chunk reader, the chunk-size configuration,
the lane/read file naming, --test-wgs (which needs no GPU, main.rs:127-153) and the CLI's exit codes."""
import gzip

import numpy as np
import os
import subprocess

import pytest

from mini_parallel_b200 import aligner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "build", "rustseq_mini")


def _write_fastq(path, reads, gz=False, crlf=False, trailing_newline=True):
    nl = "\r\n" if crlf else "\n"
    text = nl.join(f"@r{k}{nl}{r}{nl}+{nl}{'I' * len(r)}" for k, r in enumerate(reads))
    if trailing_newline:
        text += nl
    data = text.encode()
    if gz:
        with gzip.open(path, "wb") as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


@pytest.fixture
def clean_env(monkeypatch):
    for k in ("GPU_CHUNK_SIZE_READS", "GPU_CHUNK_SIZE_BASES", "WGS_DATA_DIR", "WGS_SAMPLE_ID", "WGS_LANES", "WGS_READS_PER_LANE"):
        monkeypatch.delenv(k, raising=False)
    return monkeypatch


def test_chunk_size_reads_env_errors(clean_env):
    with pytest.raises(aligner.AlignerError, match="GPU_CHUNK_SIZE_READS not set in .env file"):           # aligner.rs:11
        aligner.get_chunk_size_reads()
    clean_env.setenv("GPU_CHUNK_SIZE_READS", "12x")
    with pytest.raises(aligner.AlignerError, match=r"Invalid GPU_CHUNK_SIZE_READS value '12x': invalid digit found in string"):
        aligner.get_chunk_size_reads()
    clean_env.setenv("GPU_CHUNK_SIZE_READS", "")
    with pytest.raises(aligner.AlignerError, match="cannot parse integer from empty string"):
        aligner.get_chunk_size_reads()
    clean_env.setenv("GPU_CHUNK_SIZE_READS", "99999999999999999999999")
    with pytest.raises(aligner.AlignerError, match="number too large to fit in target type"):
        aligner.get_chunk_size_reads()
    clean_env.setenv("GPU_CHUNK_SIZE_READS", "10000")
    assert aligner.get_chunk_size_reads() == 10000
    assert aligner.get_chunk_size_bases() == 0


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("crlf", [False, True])
def test_chunk_reader_chunks_and_tail(tmp_path, clean_env, gz, crlf):
    reads = ["ACGT" * (k % 7 + 1) for k in range(23)]
    p = tmp_path / ("x.fastq.gz" if gz else "x.fastq")
    _write_fastq(p, reads, gz=gz, crlf=crlf)
    chunks = []
    aligner.process_fastq_file_in_chunks(p, 5, lambda c: chunks.append(list(c)))
    assert [len(c) for c in chunks] == [5, 5, 5, 5, 3]                                  # aligner.rs:143-147, :168-170
    assert [r.decode() for c in chunks for r in c] == reads


def test_chunk_reader_no_trailing_newline_and_partial_record(tmp_path, clean_env):
    p = tmp_path / "y.fastq"
    with open(p, "wb") as f:
        f.write(b"@a\nACGT\n+\nIIII\n@b\nGGCC")                                        # line 6 is a sequence line without '\n'
    got = []
    aligner.process_fastq_file_in_chunks(p, 100, got.extend)
    assert got == [b"ACGT", b"GGCC"]


def test_chunk_reader_carriage_returns_follow_bufread_lines(tmp_path, clean_env):
    """BufRead::lines strips "\n" and "\r\n" (aligner.rs:133): a '\r' in the middle of a line is a byte of the line, and a last
    line WITHOUT '\n' keeps a trailing '\r'."""
    p = tmp_path / "cr.fastq"
    p.write_bytes(b"@a\r\nAC\rGT\r\n+\r\nIIIII\r\n@b\r\nACGT\r")           # record 2 is cut after its sequence line, which ends in '\r'
    got = []
    aligner.process_fastq_file_in_chunks(p, 10, lambda ch: got.extend(ch))
    assert got == [b"AC\rGT", b"ACGT\r"]
    p.write_bytes(b"@a\r\nACGT\r\n+\r\nIIII\r\n")
    got = []
    aligner.process_fastq_file_in_chunks(p, 10, lambda ch: got.extend(ch))
    assert got == [b"ACGT"]


def test_chunk_reader_bases_cap(tmp_path, clean_env):
    reads = ["A" * 150] * 10
    p = tmp_path / "z.fastq"
    _write_fastq(p, reads)
    clean_env.setenv("GPU_CHUNK_SIZE_BASES", "400")                                     # README.md:32: flush when either cap is hit
    sizes = []
    aligner.process_fastq_file_in_chunks(p, 1000, lambda c: sizes.append(len(c)))
    assert sizes == [3, 3, 3, 1]


def test_chunk_reader_invalid_utf8_tolerance(tmp_path, clean_env):
    p = tmp_path / "bad.fastq"
    with open(p, "wb") as f:
        f.write(b"@a\nACGT\n+\nIIII\n" + b"\xff\xfe\n" * 3 + b"@b\nTTTT\n+\nIIII\n")    # 3 undecodable lines are skipped, not counted
    got = []
    aligner.process_fastq_file_in_chunks(p, 10, got.extend)
    assert got == [b"ACGT", b"TTTT"]
    with open(p, "wb") as f:
        f.write(b"\xff\n" * 11)
    with pytest.raises(aligner.AlignerError, match=r"Too many read errors \(>10\), stopping at line 0"):   # aligner.rs:160-162
        aligner.process_fastq_file_in_chunks(p, 10, got.extend)


def test_processor_error_aborts(tmp_path, clean_env):
    p = tmp_path / "e.fastq"
    _write_fastq(p, ["ACGT"] * 10)

    def boom(chunk):
        raise ValueError("stop here")
    with pytest.raises(ValueError, match="stop here"):
        aligner.process_fastq_file_in_chunks(p, 2, boom)


def test_missing_file_is_an_error(tmp_path, clean_env):
    with pytest.raises(aligner.AlignerError, match="Failed to open file"):               # aligner.rs:124
        aligner.process_fastq_file_in_chunks(tmp_path / "nope.fastq", 10, lambda c: None)


def test_count_bases_and_test_wgs_without_gpu(tmp_path, clean_env):
    """--test-wgs counts bases of L001 R1/R2 and needs no GPU (main.rs:127-153)."""
    reads1 = ["ACGTN" * 30] * 40
    reads2 = ["TTGCA" * 30] * 25
    _write_fastq(tmp_path / "SYN_L001_R1_001.fastq.gz", reads1, gz=True)
    _write_fastq(tmp_path / "SYN_L001_R2_001.fastq.gz", reads2, gz=True)
    with pytest.raises(aligner.AlignerError, match="GPU_CHUNK_SIZE_READS not set"):     # aligner.rs:538
        aligner.count_bases_in_fastq(tmp_path / "SYN_L001_R1_001.fastq.gz")
    clean_env.setenv("GPU_CHUNK_SIZE_READS", "16")
    assert aligner.count_bases_in_fastq(tmp_path / "SYN_L001_R1_001.fastq.gz") == 150 * 40
    env = dict(os.environ, GPU_CHUNK_SIZE_READS="16", WGS_DATA_DIR=str(tmp_path), WGS_SAMPLE_ID="SYN", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([CLI, "--test-wgs"], env=env, capture_output=True, text=True)
    assert out.returncode == 0
    assert "Successfully counted 6000 bases in SYN_L001_R1_001.fastq.gz" in out.stdout
    assert "Successfully counted 3750 bases in SYN_L001_R2_001.fastq.gz" in out.stdout


def test_wgs_file_naming(clean_env):
    clean_env.setenv("WGS_DATA_DIR", "/data/wgs")
    clean_env.setenv("WGS_SAMPLE_ID", "NA12878")
    files = aligner.wgs_file_list()
    assert len(files) == 16                                                              # 8 lanes x R1/R2 (aligner.rs:190-204)
    assert files[0] == "/data/wgs/NA12878_L001_R1_001.fastq.gz"
    assert files[1] == "/data/wgs/NA12878_L001_R2_001.fastq.gz"
    assert files[-1] == "/data/wgs/NA12878_L008_R2_001.fastq.gz"
    clean_env.setenv("WGS_LANES", "2")
    clean_env.setenv("WGS_READS_PER_LANE", "1")
    assert aligner.wgs_file_list() == ["/data/wgs/NA12878_L001_R1_001.fastq.gz", "/data/wgs/NA12878_L002_R1_001.fastq.gz"]
    clean_env.setenv("WGS_LANES", "many")                                                # unparsable -> default 8 (aligner.rs:188-191)
    assert len(aligner.wgs_file_list()) == 8


def test_cli_gates_and_exit_codes_without_gpu(tmp_path):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([CLI, "-1", "ACGT", "-2", "ACGA", "--gpu"], env=env, capture_output=True, text=True)
    assert r.returncode == 1 and "error: gpu acceleration is required and no compatible gpu was found" in r.stderr   # main.rs:161
    r = subprocess.run([CLI, "-1", "ACGT", "-2", "ACGA"], env=env, capture_output=True, text=True)                     # no --gpu
    assert r.returncode == 1 and "no compatible gpu was found" in r.stderr
    r = subprocess.run([CLI, "--full-wgs", "--gpu"], env=env, capture_output=True, text=True)
    assert r.returncode == 1 and "error: gpu acceleration is required for full WGS processing" in r.stderr           # main.rs:77
    r = subprocess.run([CLI, "-1", "ACGT"], env=env, capture_output=True, text=True)
    assert r.returncode == 101 and "--seq2 is required when not in test mode" in r.stderr                             # main.rs:157 (panic)
    r = subprocess.run([CLI, "--bogus"], env=env, capture_output=True, text=True)
    assert r.returncode == 2
    r = subprocess.run([CLI, "--help"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "--full-wgs" in r.stdout and "--test-wgs" in r.stdout and "--chunk-size" in r.stdout


def test_dotenv_is_loaded_and_does_not_override(tmp_path):
    _write_fastq(tmp_path / "S_L001_R1_001.fastq.gz", ["ACGT"] * 3, gz=True)
    _write_fastq(tmp_path / "S_L001_R2_001.fastq.gz", ["ACGT"] * 2, gz=True)
    (tmp_path / ".env").write_text(f"GPU_CHUNK_SIZE_READS=7\nWGS_DATA_DIR={tmp_path}\nWGS_SAMPLE_ID=WRONG\n")
    env = {k: v for k, v in os.environ.items() if not k.startswith(("GPU_CHUNK", "WGS_"))}
    env.update(WGS_SAMPLE_ID="S", CUDA_VISIBLE_DEVICES="")                               # the process environment wins over .env
    r = subprocess.run([CLI, "-t"], env=env, cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and "Successfully counted 12 bases in S_L001_R1_001.fastq.gz" in r.stdout


def test_bgzf_files_are_ordinary_gzip_for_the_host_reader(tmp_path, clean_env):
    """Blocked gzip is multi-member gzip: the host reader (zlib) and the Python walker agree on it without a GPU."""
    import zlib
    from mini_parallel_b200 import bgzf
    rng = np.random.default_rng(9)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(rng.integers(1, 200)))) for _ in range(3000)]
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (k, r, b"I" * len(r)) for k, r in enumerate(reads))
    gz = bgzf.compress(text, 6, 5000)
    blocks, used = bgzf.walk(gz)
    assert used == len(gz) and blocks[-1][2] == 0 and sum(b[2] for b in blocks) == len(text)      # ends with the empty EOF block
    assert b"".join(zlib.decompress(gz[o:o + n], -15) for o, n, m in blocks) == text
    path = tmp_path / "SYN_L001_R1_001.fastq.gz"
    path.write_bytes(gz)
    os.environ["GPU_CHUNK_SIZE_READS"] = "700"
    assert aligner.count_bases_in_fastq(path) == sum(len(r) for r in reads)
    with pytest.raises(ValueError):
        bgzf.walk(gzip.compress(text))                                                            # plain gzip has no BC field


def test_checkpoint_round_trip_and_serde_layout(tmp_path, clean_env):
    """CheckpointState (aligner.rs:23-104): what save writes is what serde_json::to_string_pretty would write for the
    reference's structs (plus score64), and load reads it back -- also when it was written by the reference's serde."""
    import json
    files = []
    for k, (name, score, done) in enumerate((("/data/S_L001_R1_001.fastq.gz", 12345, True), ('/data/we"ird\\name.gz', -7, False))):
        fc = aligner.FileCheckpoint()
        fc.file_path = name.encode(); fc.file_index = k; fc.score = score; fc.score64 = score + (1 << 40) * k
        fc.processing_time_ms = 1234.0 + k; fc.total_bases = 150 * (k + 1); fc.total_reads = k + 1; fc.completed = int(done)
        files.append(fc)
    path = tmp_path / "checkpoint_wgs_1.json"
    assert aligner.checkpoint_load(path) is None                                   # aligner.rs:81
    aligner.checkpoint_save(path, "wgs_1", files, 16)
    text = path.read_text()
    doc = json.loads(text)
    assert list(doc) == ["run_id", "files", "total_files", "completed_files"]      # declaration order, aligner.rs:35-40
    assert doc["run_id"] == "wgs_1" and doc["total_files"] == 16 and doc["completed_files"] == 1
    assert list(doc["files"][0])[:7] == ["file_path", "file_index", "score", "processing_time_ms", "total_bases", "total_reads", "completed"]
    assert doc["files"][1]["file_path"] == '/data/we"ird\\name.gz' and doc["files"][1]["completed"] is False
    assert text.startswith('{\n  "run_id": "wgs_1",\n  "files": [\n    {\n      "file_path"')      # to_string_pretty indentation
    rid, back, tot = aligner.checkpoint_load(path)
    assert rid == "wgs_1" and tot == 16 and len(back) == 2
    for a, b in zip(files, back):
        for f, _ in aligner.FileCheckpoint._fields_:
            assert getattr(a, f) == getattr(b, f), f
    # a file in the reference's own format (no score64)
    path.write_text(json.dumps({"run_id": "x", "files": [{"file_path": "a", "file_index": 3, "score": 9, "processing_time_ms": 1.5,
                                                          "total_bases": 7, "total_reads": 2, "completed": True}], "total_files": 4, "completed_files": 1}, indent=2))
    rid, back, tot = aligner.checkpoint_load(path)
    assert (rid, tot, back[0].file_index, back[0].score, back[0].score64, back[0].completed) == ("x", 4, 3, 9, 9, 1)
    path.write_text("{ not json")
    with pytest.raises(aligner.AlignerError, match="Failed to parse checkpoint"):
        aligner.checkpoint_load(path)


def test_bgzf_readers_hand_segments_over_in_order(tmp_path):
    """The --full-wgs driver reads a BGZF file with several pread() threads (reader r takes segments r, r+R, ...): a segment
    takes its buffers after its predecessor did and is walked after it, because its first block starts where the last whole
    block of the predecessor ended.  Whatever the reader count, the segment size and the number of buffers, the consumer must
    see every block exactly once, in stream order (rsm_debug_bgzf_segments hashes payloads + sizes in the order it gets them)."""
    from mini_parallel_b200 import bgzf
    rng = np.random.default_rng(91)
    text = b"".join(b"@r%d\n" % k + bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(rng.integers(30, 200)))]) + b"\n+\n" +
                    b"I" * 40 + b"\n" for k in range(60_000))
    gz = bgzf.compress(text, 1, 30_000)                      # blocks of ~9 KB: hundreds per segment, many segments
    blocks, used = bgzf.walk(gz)
    assert used == len(gz)
    path = tmp_path / "a.fastq.gz"
    path.write_bytes(gz)
    h = 1469598103934665603
    for off, n, m in blocks:                                 # the same FNV-1a, straight over the file
        for byte in gz[off:off + n]:
            h = ((h ^ byte) * 1099511628211) & ((1 << 64) - 1)
        h = ((h ^ m) * 1099511628211) & ((1 << 64) - 1)
    seen = set()
    for readers, seg_kb, pool in ((1, 128, 2), (1, 4096, 3), (2, 128, 3), (3, 200, 4), (4, 131, 2), (8, 128, 9), (16, 257, 5), (2, 1 << 20, 3)):
        out = aligner.debug_bgzf_segments(path, readers, seg_kb << 10, pool)
        assert out["status"] == 0 and out["blocks"] == len(blocks) and out["text_bytes"] == len(text), (readers, seg_kb, pool, out)
        assert out["hash"] == h, (readers, seg_kb, pool)
        assert out["segments"] == max(1, -(-len(gz) // (seg_kb << 10)))
        seen.add(out["segments"])
    assert len(seen) >= 4
    # a file that stops being BGZF half way: status 2 (the driver then hands the file to the host reader), no hang
    bad = bytearray(gz)
    bad[blocks[len(blocks) // 2][0] - 18] ^= 0xFF                      # first magic byte of a block header in the middle
    (tmp_path / "b.fastq.gz").write_bytes(bytes(bad))
    for readers in (1, 3):
        assert aligner.debug_bgzf_segments(tmp_path / "b.fastq.gz", readers, 128 << 10, 3)["status"] == 2
    # a truncated file (cut inside a block) is not a whole BGZF stream either
    (tmp_path / "c.fastq.gz").write_bytes(gz[: blocks[-1][0] + 5])
    assert aligner.debug_bgzf_segments(tmp_path / "c.fastq.gz", 2, 128 << 10, 3)["status"] == 2
    # empty file: one empty final segment, nothing to walk
    (tmp_path / "d.fastq.gz").write_bytes(b"")
    out = aligner.debug_bgzf_segments(tmp_path / "d.fastq.gz", 2, 128 << 10, 3)
    assert out["status"] == 0 and out["blocks"] == 0 and out["segments"] == 1
