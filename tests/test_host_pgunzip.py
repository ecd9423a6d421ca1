"""The PARALLEL host gzip reader (csrc/host_pgunzip.h, hgz::ParallelGunzip: chunks of one plain .gz file decoded side by side
into 16-bit symbols with markers for the unknown 32 KiB in front of each chunk, accepted only where a chunk starts on the bit
the known text ends on, everything else decoded serially) against the serial reader it must be indistinguishable from
(hgz::GunzipStream, itself held against zlib in test_host_gunzip.py): the same bytes, the same error flag and the same bytes
in front of an error, on every kind of deflate stream, with chunk sizes down to a few hundred bytes so that chunk borders
fall on stored, fixed, final and empty blocks, inside headers and trailers and between members.  No GPU."""
import gzip
import io
import zlib

import numpy as np
import pytest

from mini_parallel_b200 import aligner
from test_host_gunzip import ACGT, fastq, gz

pytestmark = pytest.mark.timeout(900)       # threads and condition variables: a lost wake-up must fail a test, not hang the suite


OUT_CAP = 64 << 20          # every text of this file fits (the hooks' default is 64 x the compressed size: too little for a file of zeros)


def serial(path, cap=1 << 20):
    a, n, f = aligner.debug_gunzip(path, cap, use_zlib=False, out_cap=OUT_CAP)
    assert n == len(a)
    return a, f


def parallel(path, threads, chunk, cap=1 << 20):
    a, n, f, info = aligner.debug_pgunzip(path, threads, chunk, cap, out_cap=OUT_CAP)
    assert n == len(a)
    return a, f, info


def same_as_serial(tmp_path, name, blob, raw=None, chunks=(512, 4096, 65536), threads=(2, 5), caps=(1 << 20, 4097), want_parallel=True):
    p = tmp_path / name
    p.write_bytes(blob)
    exp, ef = serial(p)
    if raw is not None:
        assert exp == raw and not ef, name
    infos = []
    for ch in chunks:
        for t in threads:
            for cap in caps:
                got, gf, info = parallel(p, t, ch, cap)
                assert gf == ef, (name, ch, t, cap, "failed flags differ")
                assert got == exp, (name, ch, t, cap, len(got), len(exp))
                if want_parallel and len(blob) >= 3 * ch:
                    assert info["parallel"], (name, ch)
                infos.append(info)
    return infos


def test_every_level_and_strategy(tmp_path):
    rng = np.random.default_rng(21)
    texts = {
        "fastq": fastq(rng, 3000), "fastq_noisy": fastq(rng, 3000, "noisy"),
        "random": rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes(),       # incompressible: stored blocks, nothing for the block finder
        "zeros": bytes(3_000_000), "two": bytes(rng.integers(0, 2, 400_000, dtype=np.uint8)),
        "text": (b"the quick brown fox jumps over the lazy dog. " * 20000)[:777_777],
    }
    accepted = 0
    for tn, raw in texts.items():
        for level in (0, 1, 6, 9):
            accepted += sum(i["accepted"] for i in same_as_serial(tmp_path, f"{tn}_{level}.gz", gz(raw, level), raw))
        for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            same_as_serial(tmp_path, f"{tn}_s{strat}.gz", gz(raw, 6, strat), raw, chunks=(4096,), threads=(3,))
        same_as_serial(tmp_path, f"{tn}_w9.gz", gz(raw, 9, wbits=16 + 9, memlevel=1), raw, chunks=(600, 8192))   # small window, tiny blocks
    assert accepted > 300          # the chain did close: most of the text came from the decoder threads


def test_the_workers_do_the_work_on_fastq(tmp_path):
    """On FASTQ text at the usual levels nearly every chunk is accepted as the workers decoded it; serial stretches are the
    exception (the first block, a chunk border on a stored block)."""
    rng = np.random.default_rng(22)
    raw = fastq(rng, 40_000, "noisy")
    for level in (1, 6):
        p = tmp_path / f"fq{level}.gz"
        blob = gz(raw, level)
        p.write_bytes(blob)
        got, failed, info = parallel(p, 4, 65536)
        assert got == raw and not failed and info["parallel"]
        n_chunks = (len(blob) + 65535) // 65536
        assert info["accepted"] >= n_chunks - 2, (info, n_chunks)
        assert info["serial_stretches"] <= 3, info


def test_multi_member_header_fields_and_trailing_garbage(tmp_path):
    rng = np.random.default_rng(23)
    a, b, c = fastq(rng, 6000), fastq(rng, 8000, "noisy"), b""
    buf = io.BytesIO()
    with gzip.GzipFile(filename="lane1.fastq", mode="wb", fileobj=buf, compresslevel=1, mtime=12345) as f:      # FNAME set
        f.write(a)
    named = buf.getvalue()
    extra = bytearray(gz(b, 6))
    extra[3] |= 4 | 16                                                   # FEXTRA + FCOMMENT
    extra = bytes(extra[:10]) + b"\x05\x00hello" + b"a comment\x00" + bytes(extra[10:])
    multi = named + extra + gz(c, 9) + gz(a[:100_000], 1)
    raw = a + b + c + a[:100_000]
    same_as_serial(tmp_path, "multi.gz", multi, raw)
    same_as_serial(tmp_path, "garbage.gz", multi + b"\x00\x00trailing bytes that are not a member", raw)
    same_as_serial(tmp_path, "garbage2.gz", multi + b"\x1f", raw, chunks=(4096,))
    # a header longer than several chunks (FNAME + FCOMMENT): the first chunks hold no deflate data at all
    long_hdr = bytearray(gz(b, 6)); long_hdr[3] |= 8 | 16
    long_hdr = bytes(long_hdr[:10]) + b"n" * 3000 + b"\x00" + b"c" * 5000 + b"\x00" + bytes(long_hdr[10:])
    same_as_serial(tmp_path, "long_header.gz", long_hdr, b, chunks=(512, 2048))
    same_as_serial(tmp_path, "long_header_twice.gz", named + long_hdr, a + b, chunks=(512,))
    # BGZF-like: every block is a member of its own, all blocks final -- nothing for the block finder, all of it serial
    same_as_serial(tmp_path, "bgzf_like.gz", b"".join(gz(a[k:k + 60000], 1) for k in range(0, len(a), 60000)), a, chunks=(4096, 65536))
    # not gzip at all / tiny: the serial reader's business
    for name, blob in (("plain.gz", a[:50_000]), ("x.gz", b"x"), ("empty.gz", b""), ("hdr_only.gz", gz(b"", 6)[:10])):
        p = tmp_path / name
        p.write_bytes(blob)
        exp, ef = serial(p)
        got, gf, info = parallel(p, 4, 512)
        assert (got, gf) == (exp, ef), name


def test_bgzf_files_run_as_whole_members(tmp_path):
    """A file that begins with a BGZF member is read as BGZF: the decoder threads inflate runs of whole members (sizes from the
    headers, CRC-32 and length of every member checked by the thread, no markers), the chain takes a run when it starts at
    the header it stands in front of; anything else -- a member that fails its check, a plain member in between, a
    truncated tail -- goes member by member through the serial decoder, and the runs pick up again behind it."""
    from mini_parallel_b200 import bgzf
    rng = np.random.default_rng(29)
    raw = fastq(rng, 9000, "noisy")
    whole = bgzf.compress(raw, 1, 65280, eof=True)
    infos = same_as_serial(tmp_path, "b.gz", whole, raw, chunks=(512, 30_000, 200_000), threads=(2, 5))
    assert all(i["parallel"] and i["accepted"] >= 3 and i["serial_stretches"] == 0 for i in infos), infos
    # two bgzip files joined (an empty EOF member in the middle), small members, another level
    joined = bgzf.compress(raw[:700_000], 1, 65280, eof=True) + bgzf.compress(raw[700_000:], 6, 9_000, eof=True)
    infos = same_as_serial(tmp_path, "bb.gz", joined, raw, chunks=(4096, 100_000))
    assert all(i["serial_stretches"] == 0 for i in infos), infos
    # a plain member between BGZF members, a BGZF tail behind a plain file's worth of data
    mixed = bgzf.compress(raw[:500_000], 1, 65280, eof=False) + gz(raw[500_000:900_000], 6) + bgzf.compress(raw[900_000:], 1, 65280, eof=True)
    infos = same_as_serial(tmp_path, "mixed.gz", mixed, raw, chunks=(4096, 50_000))
    assert all(i["accepted"] >= 2 and i["serial_stretches"] >= 1 for i in infos), infos
    # truncated inside a member, inside a header, inside a trailer; garbage behind the last member
    for cut in (len(whole) - 1, len(whole) - 20, len(whole) - 30, len(whole) // 2, len(whole) // 2 + 7, 100_000):
        same_as_serial(tmp_path, f"bcut{cut}.gz", whole[:cut], chunks=(4096, 60_000), threads=(4,), caps=(1 << 16,))
    same_as_serial(tmp_path, "bgarbage.gz", whole + b"not a member at all", raw, chunks=(4096, 60_000), threads=(4,), caps=(1 << 16,))
    # a member with a damaged CRC-32, a damaged ISIZE, damaged deflate data, a damaged BSIZE: the same bytes and the same verdict
    blocks, _ = bgzf.walk(whole)
    off, ln, _ = blocks[len(blocks) // 2]                    # payload offset (= member start + 18) and length; the trailer follows it
    for name, at, x in (("crc", off + ln + 1, 0x10), ("isize", off + ln + 5, 0x01), ("data", off + ln // 2, 0x04), ("bsize", off - 2, 0x20), ("magic", off - 17, 0x01)):
        bad = bytearray(whole); bad[at] ^= x
        same_as_serial(tmp_path, f"bbad_{name}.gz", bytes(bad), chunks=(4096, 60_000), threads=(4,), caps=(1 << 16,))


def test_truncated_files_end_early_and_corrupt_ones_fail(tmp_path):
    rng = np.random.default_rng(24)
    raw = fastq(rng, 12_000, "noisy")
    blob = gz(raw, 6)
    for cut in (len(blob) - 1, len(blob) - 4, len(blob) - 5, len(blob) - 8, len(blob) - 9, len(blob) // 2, len(blob) // 3, 70_000, 20_000):
        same_as_serial(tmp_path, f"cut{cut}.gz", blob[:cut], chunks=(4096, 65536), threads=(4,), caps=(1 << 16,))
    for off in (-8, -5, -4, -1):                                          # a wrong CRC or length in the trailer
        bad = bytearray(blob); bad[off] ^= 0x40
        infos = same_as_serial(tmp_path, f"trailer{off}.gz", bytes(bad), chunks=(4096,), threads=(4,), caps=(1 << 16,))
        p = tmp_path / f"trailer{off}.gz"
        got, failed, _ = parallel(p, 4, 4096)
        assert failed and got == raw
    # two members, the first one's trailer damaged: the second is never delivered
    two = bytearray(blob + gz(raw[:50_000], 1)); two[len(blob) - 6] ^= 1
    same_as_serial(tmp_path, "two_bad.gz", bytes(two), chunks=(4096,), threads=(4,), caps=(1 << 16,))
    # damaged headers of a later member
    for k, v in ((2, 7), (3, 0xE0)):
        two = bytearray(blob + gz(raw[:50_000], 1)); two[len(blob) + k] = v
        same_as_serial(tmp_path, f"hdr{k}.gz", bytes(two), chunks=(4096,), threads=(4,), caps=(1 << 16,))


def test_matches_reaching_in_front_of_the_member(tmp_path):
    """A distance beyond the start of the member is an error the decoder threads cannot see (to them every chunk has 32 KiB of
    unknown text in front of it): the chain checks the markers against the text the member really has -- none at all for its
    first chunk.  Streams made with a preset dictionary reach behind their own start."""
    import struct
    rng = np.random.default_rng(28)
    zdict = fastq(rng, 150, "noisy")[-32768:]
    for at in (0, 3000, 20_000, 40_000):                       # text of the member in front of the first match into the dictionary
        lead = bytes(rng.integers(0, 256, at, dtype=np.uint8))
        raw = lead + zdict[1000:30_000] + fastq(rng, 3000, "noisy")
        c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, zdict)
        body = c.compress(raw) + c.flush()
        blob = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + body + struct.pack("<II", zlib.crc32(raw), len(raw))
        infos = same_as_serial(tmp_path, f"zdict{at}.gz", blob, chunks=(512, 4096, 100_000), threads=(3,), caps=(1 << 16,))
        exp, ef = serial(tmp_path / f"zdict{at}.gz")
        if at < 32768:
            assert ef and len(exp) < at + 2000                  # "invalid distance too far back", right where the lead ends
        else:
            assert not ef and exp == raw                        # the dictionary is out of reach: an ordinary stream


def test_flipped_bits_anywhere(tmp_path):
    """Damage inside the deflate data: the same bytes in front of it and the same verdict as the serial reader -- never other
    bytes, a crash or a hang -- whichever chunk it falls into and however the block finder reads the damaged bits."""
    rng = np.random.default_rng(25)
    raw = fastq(rng, 5000, "noisy") + bytes(rng.integers(0, 256, 70_000, dtype=np.uint8)) + fastq(rng, 3000)
    for level, trials in ((6, 120), (1, 60)):
        blob = gz(raw, level)
        n_err = 0
        for trial in range(trials):
            bad = bytearray(blob)
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(10, len(blob) - 8))] ^= 1 << int(rng.integers(0, 8))
            p = tmp_path / "flip.gz"
            p.write_bytes(bytes(bad))
            exp, ef = serial(p, 1 << 15)
            got, gf, info = parallel(p, 3, int(rng.choice([700, 4096, 30_000])), 1 << 15)
            assert gf == ef, (level, trial)
            assert got == exp, (level, trial, len(got), len(exp))
            n_err += ef
        assert n_err > trials * 0.7


def test_large_file_default_chunks(tmp_path):
    rng = np.random.default_rng(26)
    raw = fastq(rng, 60_000, "noisy") + fastq(rng, 60_000)                # ~40 MB of text, > 3 MiB compressed: the default 1 MiB chunks
    for level in (1, 6):
        blob = gz(raw, level)
        assert len(blob) > 3 << 20
        infos = same_as_serial(tmp_path, f"big{level}.gz", blob, raw, chunks=(0,), threads=(4,), caps=(4 << 20, 999_983), want_parallel=False)
        assert all(i["parallel"] and i["accepted"] >= len(blob) // (1 << 20) - 1 for i in infos), infos


def test_fastq_reader_uses_it_and_reads_the_same(tmp_path, monkeypatch):
    """SWB_INFLATE_THREADS >= 3 puts FastqReader on the parallel reader: the same reads and base counts, the same error."""
    rng = np.random.default_rng(27)
    monkeypatch.setenv("GPU_CHUNK_SIZE_READS", "5000")
    reads = [ACGT[rng.integers(0, 4, int(rng.integers(1, 200)))].tobytes() for _ in range(150_000)]
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (k, r, bytes(np.frombuffer(b"#,:F", dtype=np.uint8)[rng.integers(0, 4, len(r))])) for k, r in enumerate(reads))
    p = tmp_path / "r.fastq.gz"
    p.write_bytes(gz(text, 1))
    assert p.stat().st_size > 3 << 20
    got = {}
    for n in ("0", "4"):
        monkeypatch.setenv("SWB_INFLATE_THREADS", n)
        out = []
        aligner.process_fastq_file_in_chunks(p, 5000, lambda ch: out.extend(ch))
        got[n] = out
        assert aligner.count_bases_in_fastq(p) == sum(len(r) for r in reads)
    assert got["0"] == got["4"] == reads
    bad = bytearray(gz(text, 6)); bad[len(bad) // 2] ^= 0xFF
    (tmp_path / "bad.fastq.gz").write_bytes(bytes(bad))
    with pytest.raises(aligner.AlignerError, match="gzip stream error"):
        aligner.process_fastq_file_in_chunks(tmp_path / "bad.fastq.gz", 5000, lambda ch: None)
    # abandoned early: the callback refuses the first chunk, the decoder threads are stopped and joined
    class Stop(Exception):
        pass
    def refuse(ch):
        raise Stop()
    for _ in range(3):
        with pytest.raises(Stop):
            aligner.process_fastq_file_in_chunks(p, 5000, refuse)
