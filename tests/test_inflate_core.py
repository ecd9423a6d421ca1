"""The DEFLATE decoder the GPU runs (csrc/swb_inflate.cuh), compiled for the host with one lane and compared with
zlib on stored / fixed / dynamic blocks, every compression level, FASTQ-like and random data, and malformed input."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "inflate_host.cpp")
LIB = os.path.join(ROOT, "build", "libinflate_host.so")


@pytest.fixture(scope="module")
def inflate():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdr = os.path.join(ROOT, "mini_parallel_b200", "csrc", "swb_inflate.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", LIB], check=True)
    lib = ctypes.CDLL(LIB)
    lib.swi_inflate_host.restype = ctypes.c_int
    lib.swi_inflate_host.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]

    def run(payload, cap):
        src = np.frombuffer(payload, dtype=np.uint8)
        out = np.zeros(cap + 8, dtype=np.uint8)
        n = ctypes.c_uint32()
        rc = lib.swi_inflate_host(src.ctypes.data if src.size else None, src.size, out.ctypes.data, cap, ctypes.byref(n))
        return rc, out[: n.value].tobytes()
    return run


def _raw(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def _fastq(rng, n):
    recs = []
    for k in range(n):
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 150))
        recs.append(b"@r%09d\n%s\n+\n%s\n" % (k, seq, b"I" * 150))
    return b"".join(recs)


def test_matches_zlib_on_every_block_type_and_level(inflate):
    rng = np.random.default_rng(1)
    samples = [b"", b"A", b"ACGT" * 4000, _fastq(rng, 200), bytes(rng.integers(0, 256, 60000, dtype=np.uint8)),
               bytes(rng.integers(0, 4, 65000, dtype=np.uint8) + 65), b"I" * 65280, _fastq(rng, 200)[:65280]]
    for data in samples:
        for level in (0, 1, 2, 4, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                rc, got = inflate(_raw(data, level, strategy), len(data))
                assert rc == 0 and got == data, (len(data), level, strategy, rc)


def test_many_random_members(inflate):
    rng = np.random.default_rng(2)
    for _ in range(300):
        n = int(rng.integers(1, 65536))
        kind = rng.integers(0, 3)
        if kind == 0:
            data = bytes(rng.integers(0, int(rng.integers(1, 256)), n, dtype=np.uint8))
        elif kind == 1:
            data = _fastq(rng, n // 316 + 1)[:n]
        else:
            motif = bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
            data = (motif * (n // len(motif) + 1))[:n]                       # short distances: overlapping copies
        rc, got = inflate(_raw(data, int(rng.integers(1, 10))), n)
        assert rc == 0 and got == data


def test_malformed_input_ends_with_an_error_not_a_hang(inflate):
    rng = np.random.default_rng(3)
    data = _fastq(rng, 100)
    good = _raw(data, 6)
    assert inflate(good, len(data) - 1)[0] != 0                              # output too small
    assert inflate(good[: len(good) // 2], len(data))[0] != 0                # truncated
    for _ in range(200):                                                     # random corruption: any status, but it returns
        bad = bytearray(good)
        for _k in range(int(rng.integers(1, 8))):
            bad[int(rng.integers(0, len(bad)))] ^= int(rng.integers(1, 256))
        rc, got = inflate(bytes(bad), len(data))
        assert rc != 0 or len(got) <= len(data)
    for _ in range(100):                                                     # pure noise
        inflate(bytes(rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8)), 4096)
