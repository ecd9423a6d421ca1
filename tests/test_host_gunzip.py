"""The host gzip reader of the FASTQ path (csrc/host_gunzip.h, hgz::GunzipStream) against zlib: same bytes on every kind of
deflate stream, the same behaviour at the edges as the gzread() it replaced (multi-member files, trailing garbage,
non-gzip files passed through, truncated files end early, corrupt data is an error), through the test hook
rsm_debug_gunzip.  No GPU."""
import gzip
import os
import zlib

import numpy as np
import pytest

from mini_parallel_b200 import aligner

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def gz(raw, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(raw) + c.flush()


def fastq(rng, n, quals="const"):
    out = []
    for k in range(n):
        ln = int(rng.integers(30, 260))
        q = b"I" * ln if quals == "const" else bytes(np.frombuffer(b"#,:F", dtype=np.uint8)[rng.integers(0, 4, ln)])
        out.append(b"@read%d lane:1\n" % k + ACGT[rng.integers(0, 4, ln)].tobytes() + b"\n+\n" + q + b"\n")
    return b"".join(out)


def both(path, cap):
    """(our bytes, our failed) and (zlib's bytes, zlib's failed) for read() calls of `cap` bytes"""
    a, na, fa = aligner.debug_gunzip(path, cap, use_zlib=False)
    b, nb, fb = aligner.debug_gunzip(path, cap, use_zlib=True)
    assert na == len(a) and nb == len(b)
    return (a, fa), (b, fb)


def same_as_zlib(tmp_path, name, blob, raw=None, caps=(1 << 20, 4097, 1)):
    p = tmp_path / name
    p.write_bytes(blob)
    for cap in caps:
        if cap == 1 and len(blob) > 200_000:
            continue
        (a, fa), (b, fb) = both(p, cap)
        assert fa == fb, (name, cap, "failed flags differ", fa, fb)
        assert a == b, (name, cap, len(a), len(b))
        if raw is not None:
            assert a == raw and not fa, (name, cap)


def test_every_level_and_strategy(tmp_path):
    rng = np.random.default_rng(5)
    texts = {
        "fastq": fastq(rng, 3000), "fastq_noisy": fastq(rng, 3000, "noisy"),
        "random": rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes(),       # incompressible: stored blocks
        "zeros": bytes(400_000), "two": bytes(rng.integers(0, 2, 200_000, dtype=np.uint8)),
        "text": (b"the quick brown fox jumps over the lazy dog. " * 9000)[:333_333],
        "empty": b"", "one": b"A", "short": b"ACGT\n",
    }
    for tn, raw in texts.items():
        for level in (0, 1, 2, 4, 6, 9):
            same_as_zlib(tmp_path, f"{tn}_{level}.gz", gz(raw, level), raw)
        for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            same_as_zlib(tmp_path, f"{tn}_s{strat}.gz", gz(raw, 6, strat), raw)
        same_as_zlib(tmp_path, f"{tn}_w9.gz", gz(raw, 9, wbits=16 + 9, memlevel=1), raw)     # small window, tiny blocks


def test_long_codes_and_far_matches(tmp_path):
    """Skewed symbol statistics give 12..15-bit codes (the sub-tables); a 32 KiB period gives maximum-distance matches."""
    rng = np.random.default_rng(6)
    p = np.array([2.0 ** -(k // 6) for k in range(256)]); p /= p.sum()
    skew = rng.choice(256, 600_000, p=p).astype(np.uint8).tobytes()
    same_as_zlib(tmp_path, "skew.gz", gz(skew, 9), skew)
    same_as_zlib(tmp_path, "skew_h.gz", gz(skew, 6, zlib.Z_HUFFMAN_ONLY), skew)
    block = rng.integers(0, 256, 32768, dtype=np.uint8).tobytes()
    far = block * 12 + block[:777]
    same_as_zlib(tmp_path, "far.gz", gz(far, 9), far)
    near = bytes(rng.integers(0, 4, 7, dtype=np.uint8)) * 50_000                         # distances 1..7: the byte-wise copy
    same_as_zlib(tmp_path, "near.gz", gz(near, 9), near)


def test_multi_member_header_fields_and_trailing_garbage(tmp_path):
    rng = np.random.default_rng(7)
    a, b, c = fastq(rng, 500), fastq(rng, 700, "noisy"), b""
    import io
    buf = io.BytesIO()
    with gzip.GzipFile(filename="lane1.fastq", mode="wb", fileobj=buf, compresslevel=1, mtime=12345) as f:      # FNAME set
        f.write(a)
    named = buf.getvalue()
    extra = bytearray(gz(b, 6))
    extra[3] |= 4 | 16                                                   # FEXTRA + FCOMMENT, spliced in after the 10-byte header
    extra = bytes(extra[:10]) + b"\x05\x00hello" + b"a comment\x00" + bytes(extra[10:])
    multi = named + extra + gz(c, 9) + gz(a[:1000], 1)
    same_as_zlib(tmp_path, "multi.gz", multi, a + b + c + a[:1000])
    same_as_zlib(tmp_path, "garbage.gz", multi + b"\x00\x00trailing bytes that are not a member", a + b + c + a[:1000])
    same_as_zlib(tmp_path, "bgzf_like.gz", b"".join(gz(a[k:k + 60000], 1) for k in range(0, len(a), 60000)), a)
    # a ".gz" that is not gzip at all is passed through, as gzread does
    same_as_zlib(tmp_path, "plain.gz", a[:50_000], a[:50_000])
    same_as_zlib(tmp_path, "x.gz", b"x", b"x")


def test_truncated_files_end_early_and_corrupt_ones_fail(tmp_path):
    rng = np.random.default_rng(8)
    raw = fastq(rng, 4000)
    blob = gz(raw, 6)
    for cut in (len(blob) - 1, len(blob) - 4, len(blob) - 8, len(blob) - 9, len(blob) // 2, 100, 11, 10, 5, 2, 1):
        p = tmp_path / f"cut{cut}.gz"
        p.write_bytes(blob[:cut])
        (a, fa), (b, fb) = both(p, 1 << 16)
        assert not fa and not fb, cut
        if cut < 18:                                                     # inside the header (one byte alone is not even gzip: passed through)
            assert a == b, cut
            continue
        assert raw.startswith(a) and raw.startswith(b), cut
        if cut >= 100:                                                   # everything zlib could decode, to the byte
            assert len(a) >= len(b) - 300 and len(a) <= len(raw), (cut, len(a), len(b))
    # a wrong CRC or length in the trailer
    for off in (-8, -5, -4, -1):
        bad = bytearray(blob); bad[off] ^= 0x40
        p = tmp_path / f"trailer{off}.gz"
        p.write_bytes(bytes(bad))
        (a, fa), (b, fb) = both(p, 1 << 16)
        assert fa and fb and a == raw and raw.startswith(b)              # ours: the data first, the error on the following read
                                                                         # (gzread drops what its last buffer held when the check fails)
    # flipped bits inside the deflate data: never a crash or a hang; an error whenever zlib reports one, and never different bytes
    # than zlib without an error
    n_err = 0
    for trial in range(300):
        bad = bytearray(blob)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(10, len(blob) - 8))] ^= 1 << int(rng.integers(0, 8))
        p = tmp_path / "flip.gz"
        p.write_bytes(bytes(bad))
        (a, fa), (b, fb) = both(p, 1 << 15)
        assert fa == fb, trial
        n = min(len(a), len(b))
        if not fa:
            assert a == b, trial
        else:
            n_err += 1
            # both deliver what they decoded before the damage; they may stop at different symbol boundaries
            m = min(n, max(0, n - 70_000))
            assert a[:m] == b[:m], trial
    assert n_err > 250
    # damaged headers
    for k, v in ((2, 7), (3, 0xE0)):
        bad = bytearray(blob); bad[k] = v
        p = tmp_path / f"hdr{k}.gz"
        p.write_bytes(bytes(bad))
        (a, fa), (b, fb) = both(p, 4096)
        assert fa and fb and a == b == b""


def test_large_file_crosses_every_buffer_boundary(tmp_path):
    rng = np.random.default_rng(9)
    raw = fastq(rng, 60_000, "noisy") + fastq(rng, 60_000)                # ~40 MB: many refills of the 1 MiB input and 2 MiB output buffers
    for level in (1, 6):
        same_as_zlib(tmp_path, f"big{level}.gz", gz(raw, level), raw, caps=(4 << 20, 999_983))


def test_fastq_reader_uses_it_and_agrees_with_the_zlib_path(tmp_path, monkeypatch):
    rng = np.random.default_rng(10)
    monkeypatch.setenv("GPU_CHUNK_SIZE_READS", "700")
    reads = [ACGT[rng.integers(0, 4, int(rng.integers(1, 200)))].tobytes() for _ in range(5000)]
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (k, r, b"I" * len(r)) for k, r in enumerate(reads))
    p = tmp_path / "r.fastq.gz"
    p.write_bytes(gz(text, 1) + gz(text[:0], 1))
    got = {}
    for how in ("zlib", "own"):
        if how == "zlib":
            monkeypatch.setenv("SWB_HOST_INFLATE", "zlib")
        else:
            monkeypatch.delenv("SWB_HOST_INFLATE", raising=False)
        out = []
        aligner.process_fastq_file_in_chunks(p, 700, lambda ch: out.extend(ch))
        got[how] = out
        assert aligner.count_bases_in_fastq(p) == sum(len(r) for r in reads)
    assert got["zlib"] == got["own"] == reads
    bad = bytearray(gz(text, 6)); bad[len(bad) // 2] ^= 0xFF
    (tmp_path / "bad.fastq.gz").write_bytes(bytes(bad))
    for how in ("zlib", None):
        if how:
            monkeypatch.setenv("SWB_HOST_INFLATE", how)
        else:
            monkeypatch.delenv("SWB_HOST_INFLATE", raising=False)
        with pytest.raises(aligner.AlignerError, match="gzip stream error"):
            aligner.process_fastq_file_in_chunks(tmp_path / "bad.fastq.gz", 700, lambda ch: None)


def test_inflate_on_its_own_thread_reads_the_same(tmp_path, monkeypatch):
    """SWB_ASYNC_INFLATE: the gzip stream decoded by a second thread while the reader's thread parses (used where the box has
    two cores per file): the same reads, the same error behaviour, no thread left behind when a file is abandoned early."""
    rng = np.random.default_rng(12)
    reads = [ACGT[rng.integers(0, 4, int(rng.integers(1, 200)))].tobytes() for _ in range(120_000)]      # ~25 MB of text: many 4 MiB buffers
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (k, r, b"I" * len(r)) for k, r in enumerate(reads))
    p = tmp_path / "r.fastq.gz"
    p.write_bytes(gz(text, 1))
    got = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SWB_ASYNC_INFLATE", mode)
        out = []
        aligner.process_fastq_file_in_chunks(p, 5000, lambda ch: out.extend(ch))
        got[mode] = out
    assert got["0"] == got["1"] == reads
    monkeypatch.setenv("SWB_ASYNC_INFLATE", "1")
    bad = bytearray(gz(text, 6)); bad[len(bad) // 2] ^= 0xFF
    (tmp_path / "bad.fastq.gz").write_bytes(bytes(bad))
    seen = []
    with pytest.raises(aligner.AlignerError, match="gzip stream error"):
        aligner.process_fastq_file_in_chunks(tmp_path / "bad.fastq.gz", 5000, lambda ch: seen.extend(ch))
    assert 50_000 < len(seen) < len(reads) and seen[:50_000] == reads[:50_000]        # what preceded the damage was delivered (a flipped
                                                                                       # bit decodes to wrong text for a while before a check trips)
    (tmp_path / "cut.fastq.gz").write_bytes(gz(text, 1)[:200_000])                     # truncated: ends early, no error
    seen = []
    aligner.process_fastq_file_in_chunks(tmp_path / "cut.fastq.gz", 5000, lambda ch: seen.extend(ch))
    assert 0 < len(seen) < len(reads) and seen[:-1] == reads[: len(seen) - 1]
    # a callback that gives up after the first chunk: the reader is closed with its producer blocked on a full queue
    class Stop(Exception):
        pass

    def once(ch):
        raise Stop()
    for _ in range(5):
        with pytest.raises(Stop):
            aligner.process_fastq_file_in_chunks(p, 100, once)


def test_flush_points_and_empty_stored_blocks(tmp_path):
    """Streams written with Z_SYNC_FLUSH / Z_FULL_FLUSH (pigz, network writers) carry empty stored blocks and byte-aligned
    restarts in the middle of the data."""
    rng = np.random.default_rng(13)
    raw = fastq(rng, 6000, "noisy")
    for level in (1, 6):
        c = zlib.compressobj(level, zlib.DEFLATED, 31)
        parts, pos = [], 0
        while pos < len(raw):
            n = int(rng.integers(1, 70_000))
            parts.append(c.compress(raw[pos:pos + n]))
            parts.append(c.flush(zlib.Z_SYNC_FLUSH if rng.integers(0, 2) else zlib.Z_FULL_FLUSH))
            if rng.integers(0, 4) == 0:
                parts.append(c.flush(zlib.Z_SYNC_FLUSH))                 # two flushes in a row: consecutive empty stored blocks
            pos += n
        parts.append(c.flush())
        same_as_zlib(tmp_path, f"flush{level}.gz", b"".join(parts), raw, caps=(1 << 20, 333))
