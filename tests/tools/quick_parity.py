import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import mini_parallel_b200 as mp
from mini_parallel_b200.engine import to_csr
import oracle_lib as ol
rng = np.random.default_rng(1)
A = np.frombuffer(b"ACGT", dtype=np.uint8)
def pairs(n, rl, wl, alpha=A):
    R = [alpha[rng.integers(0, alpha.size, int(rng.integers(rl[0], rl[1]+1)))] for _ in range(n)]
    W = [alpha[rng.integers(0, alpha.size, int(rng.integers(wl[0], wl[1]+1)))] for _ in range(n)]
    return R, W
eng = mp.Engine(0)
for name, (R, W) in {"short": pairs(3000, (1,160), (1,700)), "short150": pairs(2001, (150,150), (500,500)),
                     "long": pairs(40, (161,900), (1,1500)), "generic": pairs(200, (1,300), (1,600), np.frombuffer(b"ACGTN", dtype=np.uint8))}.items():
    q, qo = to_csr(R); r, ro = to_csr(W)
    for cb in ((1<<14, 1), (64<<20, 16384)):
        eng.set_chunking(*cb)
        got = eng.score_batch_csr(q, qo, r, ro)
        exp = ol.batch(q, qo, r, ro, threads=8, simd=False)
        assert np.array_equal(got, exp), name
    print(name, "ok", eng.last_routing(), flush=True)
eng.close()
