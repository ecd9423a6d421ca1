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
