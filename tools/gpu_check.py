"""Ad-hoc GPU bring-up check (not a test): random pairs vs the oracle, then a timing of config 2."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mini_parallel_b200 as mp
import oracle_lib as ol

rng = np.random.default_rng(1)
eng = mp.Engine(0)

def rand_pairs(n, rl, wl, related=True, alphabet=b"ACGT"):
    reads, wins = [], []
    al = np.frombuffer(alphabet, dtype=np.uint8)
    for k in range(n):
        n1 = int(rng.integers(rl[0], rl[1] + 1)); n2 = int(rng.integers(wl[0], wl[1] + 1))
        w = al[rng.integers(0, al.size, n2)]
        if related and n2 >= n1 and n1 > 0:
            o = int(rng.integers(0, n2 - n1 + 1)); r = w[o:o + n1].copy()
            mut = rng.random(n1) < 0.03
            r[mut] = al[rng.integers(0, al.size, int(mut.sum()))]
        else:
            r = al[rng.integers(0, al.size, n1)]
        reads.append(r); wins.append(w)
    return reads, wins

def check(name, reads, wins):
    q, qo = mp.engine.to_csr(reads); r, ro = mp.engine.to_csr(wins)
    got = eng.score_batch_csr(q, qo, r, ro)
    exp = ol.batch(q, qo, r, ro, threads=8, simd=False)
    bad = np.nonzero((got["score"] != exp["score"]) | (got["end_i"] != exp["end_i"]) | (got["end_j"] != exp["end_j"]))[0]
    print(name, "pairs", len(reads), "routing", eng.last_routing(), "mismatches", bad.size, flush=True)
    for k in bad[:5]:
        print("   pair", k, "got", got[k], "exp", exp[k], "n", qo[k+1]-qo[k], "m", ro[k+1]-ro[k])
    return bad.size == 0

ok = True
for v in (0, 1, 2, 3):
    eng.set_short_variant(v)
    ok &= check(f"v{v} uniform150x500 related", *rand_pairs(2001, (150, 150), (500, 500)))
    ok &= check(f"v{v} uniform150x500 unrelated", *rand_pairs(2000, (150, 150), (500, 500), related=False))
    ok &= check(f"v{v} ragged short", *rand_pairs(3000, (1, 160), (1, 700)))
    ok &= check(f"v{v} homopolymer", *rand_pairs(500, (1, 160), (1, 300), alphabet=b"A"))
    ok &= check(f"v{v} two-letter", *rand_pairs(1000, (100, 160), (100, 400), related=False, alphabet=b"AC"))
ok &= check("generic: N bytes", *rand_pairs(500, (1, 200), (1, 600), alphabet=b"ACGTN"))
ok &= check("generic: long reads", *rand_pairs(40, (161, 900), (200, 3000)))
ok &= check("generic: lowercase", *rand_pairs(200, (10, 150), (10, 500), alphabet=b"ACGTacgt"))
ok &= check("generic: 2k x 5k", *rand_pairs(4, (2000, 2000), (5000, 5000)))
print("ALL OK" if ok else "FAILURES", flush=True)

# compat values
for a, b in [(b"ATCGT", b"ATTGG"), (b"TGTTACGG", b"GGTTGACTA"), (b"AAAA", b"TTTT")]:
    print(a, b, eng.score_pair(a, b), eng.last_row_max(a, b), eng.ref_compat_align(a, b), "oracle", ol.sw_linear(a, b), ol.last_row_max(a, b), ol.ref_compat_align(a, b))

# timing, device-resident synthetic config 2
n, rl, wl = 1_000_000, 150, 500
dq = eng.malloc_device(n * rl); dqo = eng.malloc_device((n + 1) * 8)
dr = eng.malloc_device(n * wl); dro = eng.malloc_device((n + 1) * 8); dout = eng.malloc_device(n * 12)
for dist in (0, 1):
    eng.synth_device(0, n, rl, wl, dist, dq, dqo, dr, dro); eng.sync()
    for v in (0, 1, 2, 3):
        eng.set_short_variant(v)
        for rep in range(3):
            eng.score_batch_device(dq, dqo, n * rl, dr, dro, n * wl, n, rl, wl, dout); eng.sync()
        t = eng.last_timings()
        gcups = n * rl * wl / (t["short_ms"] * 1e-3) / 1e9
        print(json.dumps({"dist": dist, "variant": v, **t, "routing": eng.last_routing(), "short_gcups": round(gcups, 1),
                          "frac_of_9310": round(gcups / 9310, 3)}), flush=True)
    out = np.zeros(n, dtype=mp.RESULT_DTYPE); eng.d2h(out, dout, n * 12)
    print("dist", dist, "score mean", out["score"].mean(), "min", out["score"].min(), "max", out["score"].max())
    # verify a sample of the big run against the oracle
    m = 20000
    q = np.zeros(m * rl, dtype=np.uint8); r = np.zeros(m * wl, dtype=np.uint8)
    eng.d2h(q, dq, m * rl); eng.d2h(r, dr, m * wl)
    qo = np.arange(m + 1, dtype=np.uint64) * rl; ro = np.arange(m + 1, dtype=np.uint64) * wl
    exp = ol.batch(q, qo, r, ro, threads=16, simd=True)
    print("  big-run sample parity:", bool(np.array_equal(exp, out[:m])))
