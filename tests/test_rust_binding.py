"""The Rust binding (integration/rust/src/swb200.rs) cannot be compiled here (no Rust toolchain in the image), so its `extern "C"`
block is held against include/swb200.h textually: every function it declares exists in the header with the same number of
parameters, every parameter has the same shape (pointer or not, constness of the pointee, integer width and signedness), and
the return types agree; the #[repr(C)] structs have the header's fields in the header's order.  A drifted binding is undefined
behaviour at the first call -- this is the check `cargo build` would not even make."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SCALARS = {"int": "i32", "unsigned": "u32", "unsigned int": "u32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64",
             "size_t": "usize", "double": "f64", "float": "f32", "char": "i8", "uint8_t": "u8", "void": "void"}
RUST_SCALARS = {"c_int": "i32", "c_uint": "u32", "i32": "i32", "u32": "u32", "i64": "i64", "u64": "u64", "usize": "usize", "f64": "f64", "f32": "f32",
                "c_char": "i8", "u8": "u8", "c_void": "void"}
STRUCTS = {"swb_ctx": "SwbCtx", "swb_multi": "SwbMulti", "swb_result": "SwbResult", "swb_alignment": "SwbAlignment", "swb_bgzf_block": "SwbBgzfBlock",
           "swb_params": "void"}       # the binding passes params as *const c_void (always null: the reference's constants)


def c_type(t):
    """C parameter type -> (pointer depth, const pointee, base)"""
    t = re.sub(r"\s+", " ", t.strip())
    depth = t.count("*")
    const = "const" in t.split("*")[0].split() if depth else False
    base = t.replace("*", " ").replace("const", " ").replace("struct", " ").split()
    base = " ".join(base)
    base = STRUCTS.get(base, C_SCALARS.get(base))
    assert base is not None, f"unknown C type {t!r}"
    return depth, const, base


def rust_type(t):
    t = t.strip()
    depth, const = 0, False
    while t.startswith("*"):
        m = re.match(r"\*(const|mut)\s+", t)
        assert m, t
        depth += 1
        const = m.group(1) == "const"        # constness of the innermost pointee is what the C side declares
        t = t[m.end():]
    base = RUST_SCALARS.get(t, t)
    return depth, const, base


def c_functions():
    src = open(os.path.join(ROOT, "include", "swb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(swb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        ps = []
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                last = p.split()[-1]
                named = not p.rstrip().endswith("*") and last not in C_SCALARS and last.replace("*", "") not in STRUCTS
                ps.append(c_type(re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", p) if named else p))      # drop the parameter's name
        out[name] = (c_type(ret), ps)
    return out


def rust_functions():
    src = open(os.path.join(ROOT, "integration", "rust", "src", "swb200.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', src, flags=re.S).group(1)
    block = re.sub(r"//[^\n]*", "", block)
    out = {}
    for m in re.finditer(r"pub fn (swb_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        name, params, ret = m.group(1), m.group(2).strip(), (m.group(3) or "void").strip()
        ps = [rust_type(p.split(":", 1)[1]) for p in params.split(",") if p.strip()]
        out[name] = (rust_type(ret) if ret != "void" else (0, False, "void"), ps)
    return out


def test_extern_block_matches_the_header():
    c, r = c_functions(), rust_functions()
    assert len(r) >= 25
    for name, (rret, rps) in r.items():
        assert name in c, f"{name} is bound in swb200.rs but not declared in include/swb200.h"
        cret, cps = c[name]
        assert len(cps) == len(rps), f"{name}: {len(cps)} parameters in the header, {len(rps)} in the binding"
        assert cret[0] == rret[0] and cret[2] == rret[2], f"{name}: return type {cret} vs {rret}"
        for k, (cp, rp) in enumerate(zip(cps, rps)):
            assert cp[0] == rp[0], f"{name} parameter {k}: pointer depth {cp} vs {rp}"
            assert cp[2] == rp[2], f"{name} parameter {k}: type {cp} vs {rp}"
            if cp[0]:
                assert cp[1] == rp[1], f"{name} parameter {k}: constness {cp} vs {rp}"


def test_repr_c_structs_match_the_header():
    h = re.sub(r"\s+", " ", open(os.path.join(ROOT, "include", "swb200.h")).read())
    rs = open(os.path.join(ROOT, "integration", "rust", "src", "swb200.rs")).read()
    def fields(name):
        body = re.search(r"#\[repr\(C\)\][^{]*pub struct " + name + r" \{(.*?)\}", rs, flags=re.S).group(1)
        return [(f, RUST_SCALARS[t]) for f, t in re.findall(r"pub (\w+): (\w+),", body)]
    assert fields("SwbResult") == [("score", "i32"), ("end_i", "i32"), ("end_j", "i32")]
    assert "typedef struct { int32_t score; int32_t end_i; int32_t end_j; } swb_result;" in h
    assert fields("SwbAlignment") == [("start_i", "i32"), ("start_j", "i32"), ("cigar_len", "u32"), ("status", "u32"), ("cigar_off", "u64")]
    assert "typedef struct { int32_t start_i, start_j; uint32_t cigar_len; uint32_t status; uint64_t cigar_off; } swb_alignment;" in h
    assert fields("SwbBgzfBlock") == [("in_off", "u64"), ("in_len", "u32"), ("out_len", "u32")]
    assert re.search(r"typedef struct \{ uint64_t in_off; uint32_t in_len; uint32_t out_len; \} swb_bgzf_block;", h)


def _ctypes_shape(t):
    """ctypes type -> ("ptr",) or ("int", size in bytes) / ("float", size)"""
    import ctypes
    if t is None:
        return ("void",)
    if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type) and issubclass(t, (ctypes._Pointer, ctypes._CFuncPtr)):
        return ("ptr",)
    if t in (ctypes.c_double, ctypes.c_float):
        return ("float", ctypes.sizeof(t))
    return ("int", ctypes.sizeof(t))


_WIDTH = {"i32": ("int", 4), "u32": ("int", 4), "i64": ("int", 8), "u64": ("int", 8), "usize": ("int", 8), "f64": ("float", 8), "f32": ("float", 4),
          "i8": ("int", 1), "u8": ("int", 1), "void": ("void",)}


def _c_shape(ct):
    depth, _, base = ct
    return ("ptr",) if depth else _WIDTH[base]


def test_ctypes_signatures_match_the_header():
    """The Python view (mini_parallel_b200/_lib.py SIGNATURES) against the same header: arity, pointer-ness and scalar widths."""
    from mini_parallel_b200 import _lib
    c = c_functions()
    for name, (restype, argtypes) in _lib.SIGNATURES.items():
        assert name in c, name
        cret, cps = c[name]
        assert len(cps) == len(argtypes), f"{name}: {len(cps)} parameters in the header, {len(argtypes)} in _lib.py"
        assert _c_shape(cret) == _ctypes_shape(restype), f"{name}: return type"
        for k, (cp, a) in enumerate(zip(cps, argtypes)):
            assert _c_shape(cp) == _ctypes_shape(a), f"{name} parameter {k}: {cp} in the header, {a} in _lib.py"
