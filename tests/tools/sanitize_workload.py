"""One small, oracle-checked pass over EVERY kernel of libswb200.so, sized to finish under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck slow a kernel down 10-200x).  tools/sanitize.sh runs it under each tool and
keeps the logs in profiles/.  Test infrastructure: imports the oracle as the checker, like tests/."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_lib as ol                          # noqa: E402
import mini_parallel_b200 as mp                  # noqa: E402
from mini_parallel_b200 import bgzf, synth       # noqa: E402
from mini_parallel_b200.engine import to_csr     # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
SCALE = float(os.environ.get("SWB_SANITIZE_SCALE", "1"))


def n_of(n):
    return max(4, int(n * SCALE))


def rand(rng, n, alphabet=b"ACGT"):
    al = np.frombuffer(alphabet, dtype=np.uint8)
    return al[rng.integers(0, al.size, n)]


def pairs(rng, n, rl, wl, related=True, alphabet=b"ACGT", mut=0.03):
    reads, wins = [], []
    for _ in range(n):
        n1, n2 = int(rng.integers(rl[0], rl[1] + 1)), int(rng.integers(wl[0], wl[1] + 1))
        w = rand(rng, n2, alphabet)
        if related and n2 >= n1 > 0:
            o = int(rng.integers(0, n2 - n1 + 1))
            r = w[o:o + n1].copy()
            m = rng.random(n1) < mut
            r[m] = rand(rng, int(m.sum()), alphabet)
        else:
            r = rand(rng, n1, alphabet)
        reads.append(r); wins.append(w)
    return reads, wins


def parity(eng, reads, wins, what):
    q, qo = to_csr(reads); r, ro = to_csr(wins)
    got = eng.score_batch_csr(q, qo, r, ro)
    exp = ol.batch(q, qo, r, ro, threads=os.cpu_count() or 1, simd=False)
    bad = np.nonzero(got != exp)[0]
    assert bad.size == 0, f"{what}: {bad.size} pairs differ, first {bad[0]}: {got[bad[0]]} vs {exp[bad[0]]}"
    print(f"  ok {what}: {len(reads)} pairs, routing {eng.last_routing_ex()}", flush=True)
    guards(eng, what)
    return q, qo, r, ro, got


def guards(eng, what):
    """SWB_GUARD=1: no kernel or copy of the stage just run wrote outside its arenas."""
    bad, n, report = eng.guard_check()
    if bad >= 0:
        assert bad == 0, f"{what}: {bad} damaged guard zones round {n} arenas\n{report}"
        print(f"     guard zones of {n} arenas intact", flush=True)


def splitmix64(x):
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


def main():
    t0 = time.time()
    rng = np.random.default_rng(2026)
    eng = mp.Engine(0)
    # sw_stream_kernel (reads <= 160): uniform, ragged, ties, window limits; pack2bit + classify + chunk_prepare on the way
    parity(eng, *pairs(rng, n_of(3000), (150, 150), (500, 500)), "stream uniform 150x500")
    parity(eng, *pairs(rng, n_of(3000), (1, 160), (1, 700)), "stream ragged")
    parity(eng, *pairs(rng, n_of(800), (1, 160), (1, 600), alphabet=b"A"), "stream ties")
    parity(eng, *pairs(rng, n_of(64), (100, 160), (3900, 4096)), "stream widest windows")
    parity(eng, *pairs(rng, n_of(2000), (1, 128), (1, 700)), "stream 128-row instantiation")
    # the 256- and 320-row instantiations
    parity(eng, *pairs(rng, n_of(800), (161, 256), (1, 900)), "mid 161..256 (256 rows)")
    parity(eng, *pairs(rng, n_of(800), (161, 320), (1, 900)), "mid 161..320")
    # sw_long_kernel (32-bit bands), its byte variant, the generic kernel
    parity(eng, *pairs(rng, n_of(40), (321, 1500), (200, 3000)), "long bands")
    parity(eng, *pairs(rng, 2, (3000, 3000), (5000, 5000)), "long 3000x5000")
    parity(eng, *pairs(rng, n_of(80), (1, 700), (1, 900), alphabet=b"ACGTNacgt"), "bytes")
    reads, wins = pairs(rng, n_of(400), (0, 400), (0, 900))
    reads[0] = reads[0][:0]; wins[1] = wins[1][:0]
    r2, w2 = pairs(rng, n_of(40), (1, 300), (1, 600), alphabet=b"ACGTN")
    q, qo, r, ro, got = parity(eng, reads + r2, wins + w2, "mixed batch with empties")
    a, b = b"TGTTACGGNNACGT" * 30, b"GGTTGACTANNACG" * 45
    assert eng.last_row_max(a, b) == ol.last_row_max(a, b)
    assert eng.ref_compat_align(a, b) == ol.ref_compat_align(a, b)
    assert eng.score_pair(a, b) == ol.sw_linear(a, b)
    print("  ok generic kernel, last-row max, ref_compat_kernel", flush=True)
    guards(eng, "  ok generic kernel, last-row max, ref_compat_kernel")
    # host path cut into many small chunks over the three lanes
    eng.set_chunking(64 << 10, 64)
    parity(eng, *pairs(rng, n_of(3000), (100, 160), (300, 600)), "chunked host path")
    eng.set_chunking(32 << 20)
    # traceback: diagonal rule + matrix kernel
    m = min(len(reads), n_of(300))
    al, ops = eng.traceback_batch(q[:int(qo[m])], qo[:m + 1], r[:int(ro[m])], ro[:m + 1], got[:m])
    for k in range(m):
        a1 = q[int(qo[k]):int(qo[k + 1])].tobytes(); b1 = r[int(ro[k]):int(ro[k + 1])].tobytes()
        exp = ol.traceback(a1, b1, int(got[k]["end_i"]), int(got[k]["end_j"]))
        assert (int(al[k]["start_i"]), int(al[k]["start_j"]), eng.cigar_of(al[k], ops)) == exp, k
    lr, lw = pairs(rng, 3, (900, 1500), (1500, 2500), mut=0.08)
    lq, lqo = to_csr(lr); lwb, lwo = to_csr(lw)
    lres = eng.score_batch_csr(lq, lqo, lwb, lwo)
    lal, lops = eng.traceback_batch(lq, lqo, lwb, lwo, lres)
    for k in range(3):
        exp = ol.traceback(lr[k].tobytes(), lw[k].tobytes(), int(lres[k]["end_i"]), int(lres[k]["end_j"]))
        assert (int(lal[k]["start_i"]), int(lal[k]["start_j"]), eng.cigar_of(lal[k], lops)) == exp, k
    print("  ok traceback (diag + matrix kernels)", flush=True)
    guards(eng, "  ok traceback (diag + matrix kernels)")
    # windows as ranges of one buffer / of the resident reference; the device generators
    ref = synth.synth_reference(120_000)
    n = n_of(2000)
    q2, qo2, r2b, ro2, ws = synth.make_pairs_ref(ref, 0, n, 150, 500, 0)
    exp2 = ol.batch(q2, qo2, r2b, ro2, threads=os.cpu_count() or 1, simd=True)
    wl = np.full(n, 500, dtype=np.uint32)
    assert np.array_equal(eng.score_batch_ranges(q2, qo2, ref, ws, wl), exp2)
    eng.set_reference(ref)
    assert np.array_equal(eng.score_batch_vs_reference(q2, qo2, ws, wl), exp2)
    print("  ok ranges + resident reference", flush=True)
    guards(eng, "  ok ranges + resident reference")
    # FASTQ ingest: inflate_bgzf_kernel, fq_* kernels, in three segments with carries
    nfq = n_of(1500)
    rd = [q2[k * 150:(k + 1) * 150].tobytes() for k in range(min(nfq, n))]
    nfq = len(rd)
    text = b"".join(b"@r%d\n" % k + x + b"\n+\n" + b"I" * 150 + b"\n" for k, x in enumerate(rd))
    gz = bgzf.compress(text, 1, 20000)
    blocks, used = bgzf.walk(gz)
    comp = np.frombuffer(gz, dtype=np.uint8)
    tot_score, tot_reads, carry = 0, 0, b""
    step = max(1, len(blocks) // 3)
    for a0 in range(0, len(blocks), step):
        seg = blocks[a0:a0 + step]
        lo, hi = seg[0][0], seg[-1][0] + seg[-1][1]
        out = eng.fastq_bgzf_score(comp[lo:hi], [(o - lo, x, y) for o, x, y in seg], carry, a0 + step >= len(blocks), 5, tot_reads, 500)
        assert out["status"] == 0
        tot_score += out["score_sum"]; tot_reads += out["reads"]; carry = out["carry"]
    starts = [splitmix64(((5 << 40) + k) ^ 0xB202) % (ref.size - 500 + 1) for k in range(nfq)]
    wr = np.concatenate([ref[s:s + 500] for s in starts]); wro = np.arange(nfq + 1, dtype=np.uint64) * 500
    expfq = ol.batch(q2[:nfq * 150], qo2[:nfq + 1], wr, wro, threads=os.cpu_count() or 1, simd=True)
    assert tot_reads == nfq and tot_score == int(expfq["score"].astype(np.int64).sum())
    print("  ok BGZF ingest (inflate + index + extract + score)", flush=True)
    guards(eng, "  ok BGZF ingest (inflate + index + extract + score)")
    eng.close()
    print(f"sanitize workload done in {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
