"""FASTQ.gz ingest on the GPU (swb_fastq_bgzf_score): BGZF blocks inflated one warp per block, indexed and scored in
place, against zlib + the oracle.  Segments, carries, ragged and N-containing reads, CRLF, unterminated last lines."""
import zlib

import numpy as np
import pytest

import oracle_lib as ol
from mini_parallel_b200 import bgzf
from mini_parallel_b200.engine import to_csr

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def splitmix64(x):
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


def _expected(reads, ref, file_index, first_read, w):
    starts = [splitmix64((((file_index << 40) + first_read + k)) ^ 0xB202) % (len(ref) - w + 1) for k in range(len(reads))]
    q, qo = to_csr(reads)
    r, ro = to_csr([ref[s:s + w] for s in starts])
    res = ol.batch(q, qo, r, ro, threads=8, simd=True)
    return int(res["score"].astype(np.int64).sum()), sum(len(x) for x in reads)


def _make(rng, ref, n, eol=b"\n", with_n=True, file_index=0, w=500):
    reads, recs = [], []
    for k in range(n):
        s = splitmix64(((file_index << 40) + k) ^ 0xB202) % (len(ref) - w + 1)
        ln = int(rng.integers(1, 161))
        o = int(rng.integers(0, w - ln + 1))
        r = ref[s + o:s + o + ln].copy()
        m = rng.random(ln) < 0.02
        r[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
        if with_n and k % 37 == 0:
            r[int(rng.integers(0, ln))] = ord("N")
        reads.append(r.tobytes())
        recs.append(b"@read%d some description" % k + eol + reads[-1] + eol + b"+" + eol + b"I" * ln + eol)
    return reads, b"".join(recs)


def _run_segments(engine, gz, seg_blocks, file_index=0, w=500, carry_cap=1 << 16):
    blocks, used = bgzf.walk(gz)
    assert used == len(gz)
    comp = np.frombuffer(gz, dtype=np.uint8)
    tot = {"score_sum": 0, "reads": 0, "bases": 0}
    carry = b""
    for a in range(0, len(blocks), seg_blocks):
        seg = blocks[a:a + seg_blocks]
        lo, hi = seg[0][0], seg[-1][0] + seg[-1][1]
        rel = [(o - lo, n, m) for o, n, m in seg]
        final = a + seg_blocks >= len(blocks)
        out = engine.fastq_bgzf_score(comp[lo:hi], rel, carry, final, file_index, tot["reads"], w, carry_cap)
        assert out["status"] == 0
        carry = out["carry"]
        for k in tot:
            tot[k] += out[k]
        lines = lines + out["lines"] if a else out["lines"]
    assert carry == b""
    tot["lines"] = lines
    return tot


@pytest.mark.parametrize("eol,with_n", [(b"\n", True), (b"\r\n", False)])
def test_bgzf_segments_match_the_oracle(engine, eol, with_n):
    rng = np.random.default_rng(77)
    ref = ACGT[rng.integers(0, 4, 120_000)]
    engine.set_reference(ref)
    reads, text = _make(rng, ref, 9000, eol, with_n, file_index=3)
    exp_score, exp_bases = _expected(reads, ref, 3, 0, 500)
    for level, block_size, seg_blocks in ((1, 65280, 1000), (6, 20000, 7), (9, 3001, 1), (0, 65280, 3)):
        gz = bgzf.compress(text, level, block_size)
        assert zlib.decompress(gz, 31) == text[: len(zlib.decompress(gz, 31))]          # zcat-compatible (first member at least)
        tot = _run_segments(engine, gz, seg_blocks, file_index=3)
        assert tot == {"score_sum": exp_score, "reads": len(reads), "bases": exp_bases, "lines": 4 * len(reads)}, (level, block_size, seg_blocks)


def test_unterminated_last_line_and_truncated_record(engine):
    rng = np.random.default_rng(78)
    ref = ACGT[rng.integers(0, 4, 50_000)]
    engine.set_reference(ref)
    reads, text = _make(rng, ref, 200, with_n=False)
    # (a) no newline after the last quality line; (b) the file ends inside the last sequence line: the reference's reader
    # still yields that line as a read (BufRead::lines, aligner.rs:133-141)
    for cut, n_reads in ((len(text) - 1, 200), (text.rfind(b"\n+\n") - 5, 200)):
        t = text[:cut]
        rd = list(reads)
        if n_reads == 200 and cut != len(text) - 1:
            rd[-1] = rd[-1][:-5]
        exp_score, exp_bases = _expected(rd, ref, 0, 0, 500)
        tot = _run_segments(engine, bgzf.compress(t, 1, 4000), 5)
        assert {k: tot[k] for k in ("score_sum", "reads", "bases")} == {"score_sum": exp_score, "reads": n_reads, "bases": exp_bases}
        assert tot["lines"] == t.count(b"\n") + 1


def test_bad_data_asks_for_the_host_path(engine):
    rng = np.random.default_rng(79)
    ref = ACGT[rng.integers(0, 4, 50_000)]
    engine.set_reference(ref)
    _, text = _make(rng, ref, 300, with_n=False)
    gz = bytearray(bgzf.compress(text, 6, 8000))
    blocks, _ = bgzf.walk(bytes(gz))
    o, n, m = blocks[2]
    for k in range(o + 10, o + 40):
        gz[k] ^= 0x5A                                                                   # corrupt one block's payload
    out = engine.fastq_bgzf_score(np.frombuffer(bytes(gz), dtype=np.uint8), blocks, b"", True, 0, 0, 500)
    assert out["status"] == 1 and out["reads"] == 0
    latin = text.replace(b"some description", "déscription!!!!".encode("latin1"), 1)   # a non-ASCII byte: lines() semantics -> host
    out = engine.fastq_bgzf_score(np.frombuffer(bgzf.compress(latin, 1), dtype=np.uint8), bgzf.walk(bgzf.compress(latin, 1))[0], b"", True, 0, 0, 500)
    assert out["status"] == 1
    reads, text = _make(rng, ref, 50, with_n=False)                                     # and the context still works
    tot = _run_segments(engine, bgzf.compress(text, 1), 100)
    assert tot["reads"] == 50


def test_crlf_at_chunk_and_tile_boundaries(engine):
    """Many short CRLF records of varying length: the "\\r\\n" that ends a sequence line falls on every offset of the index
    kernel's 16-byte chunks, 512-byte warps and 4096-byte tiles (where the '\\r' belongs to another thread, warp or CTA).
    One missed '\\r' would show up as an extra base and a different score."""
    rng = np.random.default_rng(77)
    ref = ACGT[rng.integers(0, 4, 60_000)]
    engine.set_reference(ref)
    reads, recs = [], []
    for k in range(30_000):
        ln = int(rng.integers(1, 41))
        s = splitmix64(((3 << 40) + k) ^ 0xB202) % (len(ref) - 500 + 1)
        r = ref[s + 7:s + 7 + ln].tobytes()
        reads.append(r)
        recs.append(b"@" + b"h" * int(rng.integers(1, 30)) + b"\r\n" + r + b"\r\n+\r\n" + b"I" * ln + b"\r\n")
    text = b"".join(recs)
    exp_score, exp_bases = _expected(reads, ref, 3, 0, 500)
    tot = _run_segments(engine, bgzf.compress(text, 1, 60000), 7, file_index=3)
    assert tot["reads"] == len(reads) and tot["bases"] == exp_bases and tot["score_sum"] == exp_score
    assert tot["lines"] == 4 * len(reads)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_fastq_shapes_agree_with_the_host_path(engine, seed):
    """Random record shapes (reads from 0 to 3000 bases, N and lower-case bases, long headers, CRLF or LF), random block
    sizes, levels and segmentations: the GPU ingest totals equal scoring the same reads through the host API."""
    rng = np.random.default_rng(1000 + seed)
    ref = ACGT[rng.integers(0, 4, 80_000)]
    engine.set_reference(ref)
    eol = b"\r\n" if seed % 2 else b"\n"
    alphabet = np.frombuffer(b"ACGT" * 8 + b"Nacgt", dtype=np.uint8)
    reads, recs = [], []
    for k in range(int(rng.integers(300, 2500))):
        ln = int(rng.choice([0, 1, 36, 100, 150, 151, 160, 161, 250, 3000], p=[.02, .03, .1, .2, .4, .05, .05, .05, .08, .02]))
        r = bytes(alphabet[rng.integers(0, alphabet.size, ln)])
        reads.append(r)
        recs.append(b"@" + bytes(rng.integers(48, 123, int(rng.integers(1, 120)), dtype=np.uint8)) + eol + r + eol + b"+" + eol + b"#" * ln + eol)
    text = b"".join(recs)
    w = 400
    starts = np.array([splitmix64(((7 << 40) + k) ^ 0xB202) % (len(ref) - w + 1) for k in range(len(reads))], dtype=np.uint64)
    q, qo = to_csr(reads)
    host = engine.score_batch_vs_reference(q, qo, starts, np.full(len(reads), w, dtype=np.uint32))
    exp = {"score_sum": int(host["score"].astype(np.int64).sum()), "reads": len(reads), "bases": sum(len(r) for r in reads)}
    for _ in range(3):
        gz = bgzf.compress(text, int(rng.integers(0, 10)), int(rng.integers(200, 65281)))
        tot = _run_segments(engine, gz, int(rng.integers(1, 40)), file_index=7, w=w)
        assert {k: tot[k] for k in exp} == exp
        assert tot["lines"] == 4 * len(reads)


def test_stray_carriage_returns_are_left_to_the_host_reader(engine):
    """a '\\r' that is not the first half of "\\r\\n" -- in the middle of a sequence line, or ending a last line
    that has no '\\n' -- is a byte of its line for BufRead::lines (a mismatching base), not a terminator.  The in-place
    masking of the GPU path only knows terminators, so it must decline such text (status 1), at every position of the index
    kernel's 16-byte chunks; plain CRLF text, also cut between '\\r' and '\\n' by a segment boundary, stays on the GPU."""
    rng = np.random.default_rng(80)
    ref = ACGT[rng.integers(0, 4, 50_000)]
    engine.set_reference(ref)
    reads, text = _make(rng, ref, 120, eol=b"\r\n", with_n=False)
    exp_score, exp_bases = _expected(reads, ref, 0, 0, 500)
    tot = _run_segments(engine, bgzf.compress(text, 1, 777), 3)                       # small blocks: segment ends fall everywhere, also after a '\r'
    assert {k: tot[k] for k in ("score_sum", "reads", "bases")} == {"score_sum": exp_score, "reads": 120, "bases": exp_bases}
    seq_line = text.index(b"\r\n") + 2                                                # first byte of the first sequence line
    for off in range(0, 40):                                                          # a stray '\r' at 40 consecutive positions
        bad = bytearray(text)
        bad[seq_line + 200 + off] = 0x0D
        if bytes(bad[seq_line + 200 + off:seq_line + 202 + off]) == b"\r\n":
            continue
        gz = bgzf.compress(bytes(bad), 1)
        out = engine.fastq_bgzf_score(np.frombuffer(gz, dtype=np.uint8), bgzf.walk(gz)[0], b"", True, 0, 0, 500)
        assert out["status"] == 1 and out["reads"] == 0, off
    for pad in range(0, 17):                                                          # the text ENDS in '\r' (no '\n'), at every chunk offset
        t = text[: text.rfind(b"\r\n+\r\n")] + b"A" * pad + b"\r"
        gz = bgzf.compress(t, 1)
        out = engine.fastq_bgzf_score(np.frombuffer(gz, dtype=np.uint8), bgzf.walk(gz)[0], b"", True, 0, 0, 500)
        assert out["status"] == 1, pad
    tot = _run_segments(engine, bgzf.compress(text, 1), 100)                          # and the context still works
    assert tot["reads"] == 120


def test_files_of_short_reads_move_to_the_128_row_kernel_and_back(engine):
    """A file's segments carry the longest read of the previous one as the next one's bound: files of <= 128 bp reads (2 x 100,
    2 x 125 runs) are scored on 128 rows from their second segment on.  The bound is a hint: a later segment with longer reads
    is still exact (those reads are routed past the kernel), and the segment after it is back on 160 rows."""
    rng = np.random.default_rng(81)
    ref = ACGT[rng.integers(0, 4, 60_000)]
    engine.set_reference(ref)

    def make(lengths, file_index, first):
        reads, recs = [], []
        for k, ln in enumerate(lengths):
            s = splitmix64(((file_index << 40) + first + k) ^ 0xB202) % (len(ref) - 500 + 1)
            o = int(rng.integers(0, 500 - ln + 1))
            r = ref[s + o:s + o + ln].copy()
            m = rng.random(ln) < 0.02
            r[m] = ACGT[rng.integers(0, 4, int(m.sum()))]
            reads.append(r.tobytes())
            recs.append(b"@r%d\n" % (first + k) + reads[-1] + b"\n+\n" + b"I" * ln + b"\n")
        return reads, b"".join(recs)
    fi, first, total = 9, 0, 0
    plan = [[100] * 900, [100] * 900, list(rng.integers(60, 129, 900)), [100] * 450 + [150] * 450, [150] * 900, [125] * 900, [125] * 900]
    for seg, lengths in enumerate(plan):
        reads, text = make([int(x) for x in lengths], fi, first)
        gz = bgzf.compress(text, 1, 30000)
        blocks, used = bgzf.walk(gz)
        out = engine.fastq_bgzf_score(np.frombuffer(gz, dtype=np.uint8), blocks, b"", True, fi, first, 500)
        assert out["status"] == 0 and out["reads"] == len(reads), seg
        q, qo = to_csr(reads)
        starts = [splitmix64(((fi << 40) + first + k) ^ 0xB202) % (len(ref) - 500 + 1) for k in range(len(reads))]
        r, ro = to_csr([ref[s:s + 500] for s in starts])
        exp = ol.batch(q, qo, r, ro, threads=8, simd=True)
        assert out["score_sum"] == int(exp["score"].astype(np.int64).sum()), seg
        first += len(reads); total += len(reads)           # (segment 3 runs with the bound 128 of segment 2 and 150 bp reads: the long-pair kernel)
