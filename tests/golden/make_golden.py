"""Generates tests/golden/sw_vectors.json.  Run in the BUILD container (needs /root/reference for the
reference-kernel columns):   python tests/golden/make_golden.py

Columns per vector
  score, end_i, end_j, last_row_max, ref_compat_1024, ref_compat_256 : oracle/sw_oracle.c
  ref_detailed   : the reference's own smith_waterman_detailed (cl:74-151) executed through oracle/_ref
                   (null when len2 > 256, which that kernel cannot hold, cl:93-94)
  ref_gpu_align  : the reference's own smith_waterman_align (cl:11-71) under gpu_align's geometry with a
                   256-wide work-group (oracle/_ref; local_scores[256] caps the emulated work-group size)
The first six vectors are SURVEY.md 8c's table (row 1 = README.md:7-11 read as ATCGT / ATTGG).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

rng = np.random.default_rng(0xB200)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def rnd(n, alphabet=ACGT):
    return bytes(alphabet[rng.integers(0, alphabet.size, n)])


def mutate(s, kind, pos):
    s = bytearray(s)
    if kind == "sub":
        s[pos] = ord("A") if s[pos] != ord("A") else ord("C")
    elif kind == "ins":
        s.insert(pos, ord("G"))
    elif kind == "del":
        del s[pos]
    return bytes(s)


vectors = [
    ("survey-readme", b"ATCGT", b"ATTGG"), ("survey-identical", b"ACGT", b"ACGT"), ("survey-zero", b"AAAA", b"TTTT"),
    ("survey-gap", b"ACGTACGT", b"ACGACGT"), ("survey-gattaca", b"GATTACA", b"GCATGCU"),
    ("survey-wiki", b"TGTTACGG", b"GGTTGACTA"),
    ("empty-both", b"", b""), ("empty-read", b"", b"ACGT"), ("empty-window", b"ACGT", b""),
    ("single-match", b"A", b"A"), ("single-mismatch", b"A", b"C"),
    ("all-N", b"N" * 40, b"N" * 60), ("mixed-case", b"ACGTacgtACGT", b"acgtACGTacgt"),
    ("homopolymer-ties", b"A" * 33, b"A" * 70), ("homopolymer-short-window", b"A" * 64, b"A" * 9),
    ("two-letter-repeat", b"AC" * 40, b"CA" * 90),
]
w = rnd(500)
r150 = w[173:323]
vectors += [("identical-150", r150, r150), ("embedded-exact", r150, w), ("embedded-1sub", mutate(r150, "sub", 70), w),
            ("embedded-1ins", mutate(r150, "ins", 70)[:150], w), ("embedded-1del", mutate(r150, "del", 70), w)]
for n in (1, 31, 32, 33, 149, 150, 151, 159, 160, 161, 191, 192, 193):
    win = rnd(230)
    vectors.append((f"len-{n}", (win + rnd(200))[20:20 + n], win))
for m in (499, 500, 501, 1000):
    win = rnd(m)
    vectors.append((f"window-{m}", mutate(win[m // 3:m // 3 + 150], "sub", 10), win))
vectors += [("unrelated-150x500", rnd(150), rnd(500)), ("read-longer-than-window", rnd(150), rnd(40)),
            ("acgtn-mix", rnd(120, np.frombuffer(b"ACGTN", dtype=np.uint8)), rnd(256, np.frombuffer(b"ACGTN", dtype=np.uint8)))]
long_seq = rnd(16500)
vectors.append(("identical-16500-forces-32bit", long_seq, long_seq))
vectors.append(("long-2000x3000", rnd(2000), rnd(3000)))

have_ref = ol.ref_cl() is not None
out = []
for name, a, b in vectors:
    s, i, j = ol.sw_linear(a, b)
    rec = {"name": name, "seq1": a.decode("latin1"), "seq2": b.decode("latin1"), "score": s, "end_i": i, "end_j": j,
           "last_row_max": ol.last_row_max(a, b), "ref_compat_1024": ol.ref_compat_align(a, b, 1024),
           "ref_compat_256": ol.ref_compat_align(a, b, 256), "ref_detailed": None, "ref_gpu_align": None}
    if have_ref and len(a) <= 4096 and len(b) <= 4096:
        rec["ref_gpu_align"] = ol.ref_gpu_align(a, b, 256)
        if 0 < len(b) <= 256 and len(a) > 0:
            rec["ref_detailed"] = ol.ref_detailed(a, b, 256)
    out.append(rec)

with open(os.path.join(HERE, "sw_vectors.json"), "w") as f:
    json.dump({"generated_with_reference": have_ref, "vectors": out}, f, indent=0)
print(f"wrote {len(out)} vectors (reference kernels available: {have_ref})")
