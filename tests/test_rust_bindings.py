"""rust/valida-b200-sys/src/lib.rs against include/valida_b200.h (CPU; no Rust toolchain needed).

The crate is written by hand and cannot be compiled in this container, so the two texts are parsed here and compared
declaration by declaration: every exported function once, same order of parameters, every parameter and return type the
Rust spelling of the C one, every #define with the same value, every struct field for field — and every symbol the crate
declares is exported by the built library."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "valida_b200.h")
SYS = os.path.join(ROOT, "rust", "valida-b200-sys", "src", "lib.rs")

SCALARS = {"int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "uint8_t": "u8", "float": "f32", "double": "f64",
           "char": "c_char", "void": "c_void"}


def strip_c_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def c_type_to_rust(ctype, is_array):
    """`const vgpu_dmat* const` (+ array suffix) -> `*const *const vgpu_dmat`."""
    toks = re.findall(r"[A-Za-z_0-9]+|\*", ctype)
    # pointer levels, innermost first: (pointee_is_const) for each '*'
    base, base_const, levels = None, False, []
    pending_const = False
    for t in toks:
        if t == "const":
            if base is None:
                base_const = True
            else:
                pending_const = True          # qualifies the pointer to its left (`* const`) — the pointer itself, irrelevant for the next level's pointee... see below
            continue
        if t == "*":
            levels.append(pending_const)
            pending_const = False
            continue
        if t in ("struct",):
            continue
        if base is None:
            base = t
        else:
            raise AssertionError("unexpected token %r in %r" % (t, ctype))
    # `T* const` after the last '*' marks that pointer const: it is the pointee of an array parameter's implied pointer
    last_ptr_const = pending_const
    rust = SCALARS.get(base, base)
    # level i points at: base (i = 0) or the previous pointer; constness of the pointee
    pointee_const = base_const
    for i, _ in enumerate(levels):
        rust = ("*const " if pointee_const else "*mut ") + rust
        # is THIS pointer const-qualified (`* const`)?  recorded on the next '*' or at the end
        pointee_const = levels[i + 1] if i + 1 < len(levels) else last_ptr_const
    if is_array:
        rust = ("*const " if pointee_const else "*mut ") + rust
    return rust


def parse_header():
    text = strip_c_comments(open(HEADER).read())
    text = re.sub(r"#[^\n]*", lambda m: m.group(0) if m.group(0).startswith("#define") else "", text)
    defines = {}
    for m in re.finditer(r"#define\s+(VGPU_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", text):
        defines[m.group(1)] = int(m.group(2))
    text_nodef = re.sub(r"#define[^\n]*", "", text)
    funcs = []
    for m in re.finditer(r"\b((?:const\s+)?[A-Za-z_0-9]+\s*\**)\s*\b(vgpu_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text_nodef, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3)
        plist = []
        for p in [q.strip() for q in params.split(",") if q.strip()]:
            if p == "void":
                continue
            am = re.match(r"(.*?)([A-Za-z_0-9]+)\s*(\[[^\]]*\])?$", p, flags=re.S)
            ctype, pname, arr = am.group(1).strip(), am.group(2), am.group(3)
            plist.append((pname, c_type_to_rust(ctype, bool(arr))))
        rret = None if ret == "void" else c_type_to_rust(ret, False)
        funcs.append((name, plist, rret))
    return defines, funcs, text_nodef


def parse_rust():
    text = re.sub(r"//[^\n]*", "", open(SYS).read())
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (VGPU_[A-Z0-9_]+): \w+ = (-?\d+);", text)}
    ext = re.search(r'extern "C" \{(.*)\}\s*$', text, flags=re.S).group(1)
    funcs = []
    for m in re.finditer(r"pub fn (vgpu_[a-z0-9_]+)\((.*?)\)\s*(?:->\s*([^;]+?))?;", ext, flags=re.S):
        name, params, ret = m.group(1), m.group(2), m.group(3)
        plist = []
        for p in [q.strip() for q in params.split(",") if q.strip()]:
            pname, ptype = [s.strip() for s in p.split(":", 1)]
            plist.append((pname, ptype))
        funcs.append((name, plist, ret.strip() if ret else None))
    return consts, funcs, text


def test_every_export_is_declared_once_with_the_same_signature():
    _, cfuncs, _ = parse_header()
    _, rfuncs, _ = parse_rust()
    assert len(cfuncs) >= 50, len(cfuncs)                       # the parser saw the whole header
    cnames, rnames = [f[0] for f in cfuncs], [f[0] for f in rfuncs]
    assert len(set(cnames)) == len(cnames) and len(set(rnames)) == len(rnames)
    assert cnames == rnames, (sorted(set(cnames) ^ set(rnames)), "or the order differs")
    for (name, cp, cr), (_, rp, rr) in zip(cfuncs, rfuncs):
        assert len(cp) == len(rp), (name, cp, rp)
        for (cn, ct), (rn, rt) in zip(cp, rp):
            assert ct == rt, (name, cn, ct, rn, rt)
            assert cn == rn or (cn, rn) == ("in", "input"), (name, cn, rn)      # `in` is a Rust keyword
        assert cr == rr, (name, cr, rr)


def test_constants_agree():
    cdef, _, _ = parse_header()
    rconst, _, _ = parse_rust()
    assert set(cdef) == set(rconst), set(cdef) ^ set(rconst)
    for k, v in cdef.items():
        assert rconst[k] == v, (k, v, rconst[k])


def test_struct_layouts_agree():
    _, _, ctext = parse_header()
    _, _, rtext = parse_rust()

    def rust_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, rtext, flags=re.S).group(1)
        return [(m.group(1), m.group(2).strip()) for m in re.finditer(r"pub (\w+): ([^,\n]+),", body)]

    assert rust_fields("vgpu_matrix") == [("data", "*const u32"), ("height", "u64"), ("width", "u64")]
    assert re.search(r"typedef struct vgpu_matrix \{\s*const uint32_t\* data;\s*uint64_t height;\s*uint64_t width;\s*\} vgpu_matrix;", ctext)
    assert rust_fields("vgpu_pair_term") == [("is_preprocessed", "u32"), ("column", "u32"), ("weight", "u32")]
    assert re.search(r"struct \{ uint32_t is_preprocessed, column, weight; \} terms\[VGPU_MAX_TERMS\];", ctext)
    assert rust_fields("vgpu_pair_col") == [("constant", "u32"), ("n_terms", "u32"), ("terms", "[vgpu_pair_term; VGPU_MAX_TERMS]")]
    assert rust_fields("vgpu_interaction") == [("n_fields", "u32"), ("fields", "[vgpu_pair_col; VGPU_MAX_FIELDS]"), ("count", "vgpu_pair_col"),
                                               ("bus", "u32"), ("is_send", "u32")]
    assert re.search(r"uint32_t n_fields;\s*vgpu_pair_col fields\[VGPU_MAX_FIELDS\];\s*vgpu_pair_col count;\s*uint32_t bus;[^;]*?uint32_t is_send;", ctext, flags=re.S)
    assert rust_fields("vgpu_chip_desc") == [("chip_id", "u32"), ("width", "u32"), ("preprocessed_width", "u32"), ("n_interactions", "u32"),
                                             ("interactions", "[vgpu_interaction; VGPU_MAX_INTERACTIONS]")]
    assert re.search(r"uint32_t chip_id;[^;]*?uint32_t width, preprocessed_width;\s*uint32_t n_interactions;\s*vgpu_interaction interactions\[VGPU_MAX_INTERACTIONS\];",
                     ctext, flags=re.S)
    # sizes as the C compiler lays them out == what #[repr(C)] gives the Rust structs (every member is a u32 or an array of them,
    # so there is no padding; vgpu_matrix is pointer + 2 x u64)
    import subprocess
    import tempfile

    pair_col = 4 * (2 + 3 * 4)
    interaction = 4 + 14 * pair_col + pair_col + 8
    chip_desc = 16 + 5 * interaction
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "valida_b200.h"\nint main(void) { printf("%zu %zu %zu %zu\\n", sizeof(vgpu_matrix), '
                             'sizeof(vgpu_pair_col), sizeof(vgpu_interaction), sizeof(vgpu_chip_desc)); return 0; }\n')
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [24, pair_col, interaction, chip_desc], sizes


def test_the_built_library_exports_every_symbol_the_crate_declares():
    lib = os.path.join(ROOT, "valida_b200", "libvalida_b200.so")
    if not os.path.exists(lib):
        import pytest

        pytest.skip("libvalida_b200.so not built")
    _, rfuncs, _ = parse_rust()
    out = os.popen("nm -D --defined-only %s" % lib).read()
    exported = set(re.findall(r" T (vgpu_[a-z0-9_]+)", out))
    missing = [f[0] for f in rfuncs if f[0] not in exported]
    assert not missing, missing


def test_safe_crate_only_calls_declared_functions():
    _, rfuncs, _ = parse_rust()
    declared = {f[0] for f in rfuncs}
    safe = open(os.path.join(ROOT, "rust", "valida-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(vgpu_[a-z0-9_]+)\(", safe))
    assert used and used <= declared, used - declared
