"""The restatement against the reference's OWN kernel source, compiled unmodified by oracle/Makefile into
oracle/_ref/libref_cl.so (built in the container that has /root/reference; the .so travels to the GPU box).
Skipped when the .so is absent."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.skipif(ol.ref_cl() is None, reason="oracle/_ref/libref_cl.so not built")


def test_detailed_kernel_equals_last_row_max():
    rng = np.random.default_rng(11)
    al = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for _ in range(150):
        a = al[rng.integers(0, 5, int(rng.integers(1, 180)))]
        b = al[rng.integers(0, 5, int(rng.integers(1, 257)))]
        assert ol.ref_detailed(a, b, 256) == ol.last_row_max(a, b)


def test_global_max_from_reference_kernel_on_prefixes():
    """sw_linear's score pinned through the reference kernel alone: max over row prefixes of the dead kernel."""
    rng = np.random.default_rng(12)
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(12):
        a = al[rng.integers(0, 4, int(rng.integers(1, 40)))]
        b = al[rng.integers(0, 4, int(rng.integers(1, 120)))]
        best = max(ol.ref_detailed(a[:k], b, 256) for k in range(1, a.size + 1))
        assert best == ol.sw_linear(a, b)[0]


@pytest.mark.parametrize("wg", [32, 64, 256])
def test_live_kernel_equals_ref_compat(wg):
    rng = np.random.default_rng(13)
    for _ in range(40):
        n1, n2 = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        a = rng.integers(0, 4, n1).astype(np.uint8) + 65
        b = rng.integers(0, 4, n2).astype(np.uint8) + (65 if rng.random() < 0.8 else 97)
        assert ol.ref_gpu_align(a, b, wg) == ol.ref_compat_align(a, b, wg)


def test_live_kernel_explicit_ndrange_multi_iteration():
    """Fewer groups than the host would launch => work-items walk several strided positions (cl:39-53)."""
    import ctypes
    rng = np.random.default_rng(14)
    a = rng.integers(0, 2, 4000).astype(np.uint8) + 65
    b = rng.integers(0, 2, 4000).astype(np.uint8) + 65
    out = ctypes.c_int32()
    assert ol.ref_cl().refcl_run_align(a.ctypes.data, b.ctypes.data, 4000, 16, 5, ctypes.byref(out)) == 0
    # restate that NDRange by hand
    chunk = (4000 + 4) // 5
    best = 0
    for g in range(5):
        for lid in range(16):
            cur = 0
            for i in range(g * chunk + lid, min((g + 1) * chunk, 4000), 16):
                cur = max(cur + (2 if a[i] == b[i] else -1), 0)
                best = max(best, cur)
    assert out.value == best
