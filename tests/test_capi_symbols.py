"""The C-ABI library loads without a GPU and exports every symbol include/swb200.h declares."""
import os
import re

import mini_parallel_b200 as mp
from mini_parallel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:swb|rsm)_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = mp.load_library()
    names = _declared("swb200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/swb200.h but not exported by libswb200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in mini_parallel_b200/_lib.py"


def test_struct_layouts_match_the_header():
    """The numpy views of swb_result / swb_alignment have the C structs' layout (include/swb200.h)."""
    src = open(os.path.join(ROOT, "include", "swb200.h")).read()
    assert "typedef struct { int32_t score; int32_t end_i; int32_t end_j; } swb_result;" in re.sub(r"\s+", " ", src)
    assert "typedef struct { int32_t start_i, start_j; uint32_t cigar_len; uint32_t status; uint64_t cigar_off; } swb_alignment;" in re.sub(r"\s+", " ", src)
    assert _lib.RESULT_DTYPE.itemsize == 12 and _lib.RESULT_DTYPE.names == ("score", "end_i", "end_j")
    a = _lib.ALIGNMENT_DTYPE
    assert a.itemsize == 24 and [a.fields[n][1] for n in a.names] == [0, 4, 8, 12, 16]


def test_no_device_is_an_error_not_a_fallback():
    """Like main.rs:76-79 / :160-163: without a GPU the engine refuses to run."""
    if mp.device_count() > 0:
        return
    try:
        mp.Engine(0)
    except mp.SwbError as e:
        assert "no compatible gpu" in str(e)
    else:
        raise AssertionError("Engine() succeeded without a CUDA device")


def test_product_does_not_reference_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "mini_parallel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "oracle_lib" not in text and "sw_oracle" not in text, f
