#!/usr/bin/env python
"""Soak of the oracle's pin against the reference's own kernel source (oracle/_ref/libref_cl.so = smith_waterman.cl compiled
unmodified): the cases of tests/test_ref_emulator.py::test_global_max_from_reference_kernel_on_prefixes with other seeds and
many more of them, plus last-row and live-kernel comparisons.  Needs /root/reference only for building oracle/_ref; no GPU.

    python tests/tools/soak_ref_pin.py [--small 12000] [--large 3000] [--last-row 5000] [--live 2000] [--seed 900000]"""
import argparse
import multiprocessing as mpc
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol                                    # noqa: E402
from test_ref_emulator import _prefix_case                 # noqa: E402


def _last_row_case(seed):
    rng = np.random.default_rng(seed)
    al = np.frombuffer([b"ACGT", b"ACGTN", b"ACGTacgtNn", b"AC"][seed % 4], dtype=np.uint8)
    a = al[rng.integers(0, al.size, int(rng.integers(1, 200)))]
    b = al[rng.integers(0, al.size, int(rng.integers(1, 257)))]
    got, exp = ol.ref_detailed(a, b, 256), ol.last_row_max(a, b)
    return None if got == exp else f"seed {seed}: smith_waterman_detailed {got} != sw_last_row_max {exp}"


def _live_case(seed):
    rng = np.random.default_rng(seed)
    n1, n2 = int(rng.integers(1, 4000)), int(rng.integers(1, 4000))
    a = rng.integers(0, 4, n1).astype(np.uint8) + 65
    b = rng.integers(0, 4, n2).astype(np.uint8) + (65 if rng.random() < 0.8 else 97)
    wg = int(rng.choice([32, 64, 128, 256]))
    got, exp = ol.ref_gpu_align(a, b, wg), ol.ref_compat_align(a, b, wg)
    return None if got == exp else f"seed {seed}: smith_waterman_align under gpu_align's geometry {got} != ref_compat_align {exp}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", type=int, default=12000)
    ap.add_argument("--large", type=int, default=3000)
    ap.add_argument("--last-row", type=int, default=5000)
    ap.add_argument("--live", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=900000)
    args = ap.parse_args()
    if ol.ref_cl() is None:
        raise SystemExit("oracle/_ref/libref_cl.so is not built (make -C oracle in a container that has /root/reference)")
    alphabets = [b"ACGT", b"ACGTN", b"ACGTacgt", b"AC", b"A"]
    t0 = time.time()
    with mpc.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        legs = [("every row prefix, reads <= 40 x windows <= 128", _prefix_case, [(args.seed + k, alphabets[k % 5], 40, 128, True) for k in range(args.small)]),
                ("deciding prefixes, reads <= 160 x windows <= 256", _prefix_case, [(args.seed + 10_000_000 + k, alphabets[k % 3], 160, 256, False) for k in range(args.large)]),
                ("smith_waterman_detailed == sw_last_row_max", _last_row_case, [args.seed + 20_000_000 + k for k in range(args.last_row)]),
                ("smith_waterman_align (gpu_align geometry) == ref_compat_align", _live_case, [args.seed + 30_000_000 + k for k in range(args.live)])]
        bad = 0
        for name, fn, cases in legs:
            errs = [e for e in pool.map(fn, cases, chunksize=16) if e]
            print(f"{name}: {len(cases)} cases, {len(errs)} differences, {time.time() - t0:.0f} s", flush=True)
            for e in errs[:5]:
                print("   ", e)
            bad += len(errs)
    print("ok: the restatement and the reference kernels agree on every case" if not bad else f"FAILED: {bad} differences")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
