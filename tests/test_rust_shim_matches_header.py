"""The Rust shim (instant-distance_b200/rust/src/lib.rs) cannot be compiled in this image (no rustc/cargo), so the one thing that
can be checked mechanically is checked here: every `extern "C"` declaration and the #[repr(C)] parameter struct in the shim agree
with include/instant_distance_b200.h — same functions, same number and kind of arguments, same field order and widths."""
import os
import re

from tests.conftest import ROOT

RUST = os.path.join(ROOT, "instant-distance_b200", "rust", "src", "lib.rs")
HEADER = os.path.join(ROOT, "include", "instant_distance_b200.h")

# canonical spelling of a type on either side of the FFI
C_TYPES = {
    "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "float": "f32", "idb_status": "i32", "void": "void", "char": "c_char",
    "idb_params": "IdbParams", "idb_index": "IdbIndex",
}


def _c_type(t):
    t = t.replace("const", " ").strip()
    stars = t.count("*")
    base = t.replace("*", " ").split()[0]
    return C_TYPES.get(base, base) + "*" * stars  # (types the shim does not use keep their C spelling)


def _rust_type(t):
    t = t.strip()
    stars = 0
    while t.startswith("*const ") or t.startswith("*mut "):
        t = t.split(" ", 1)[1].strip()
        stars += 1
    return t.replace("std::ffi::", "") + "*" * stars


def _c_functions():
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"IDB_API\s+([\w\s\*]+?)\s*\b(idb_\w+)\s*\(([^)]*)\)\s*;", text):
        params = [] if args.strip() in ("", "void") else [a.strip() for a in args.split(",")]
        types = [_c_type(re.sub(r"\b\w+$", "", p)) for p in params]  # drop the parameter name
        out[name] = (_c_type(ret), types)
    return out


def _rust_functions():
    text = open(RUST).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    out = {}
    for name, args, ret in re.findall(r"fn (idb_\w+)\(([^)]*)\)(?:\s*->\s*([^;]+))?;", block):
        types = [_rust_type(a.split(":", 1)[1]) for a in args.split(",") if a.strip()]
        out[name] = (_rust_type(ret) if ret else "void", types)
    return out


def test_every_ffi_declaration_matches_the_header():
    c, r = _c_functions(), _rust_functions()
    assert {"idb_params_default", "idb_build_f32", "idb_search_batch_f32", "idb_index_free", "idb_last_error"} <= set(r)
    for name, (ret, types) in r.items():
        assert name in c, f"{name}: declared in the Rust shim but not in the header"
        c_ret, c_types = c[name]
        assert ret == c_ret, f"{name}: return type {ret} vs {c_ret}"
        assert types == c_types, f"{name}: arguments {types} vs {c_types}"


def test_params_struct_field_order_and_widths():
    header = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct idb_params \{(.*?)\} idb_params;", header, flags=re.S).group(1)
    c_fields = []
    for decl in [d.strip() for d in body.split(";") if d.strip()]:
        fp = re.match(r"void \(\*(\w+)\)\(", decl)
        if fp:
            c_fields.append((fp.group(1), "fnptr"))
            continue
        t, name = decl.rsplit(None, 1)
        name = name.lstrip("*")
        c_fields.append((name, "ptr" if "*" in decl else C_TYPES[t.split()[0]]))
    rust = open(RUST).read()
    rbody = re.search(r"struct IdbParams \{(.*?)\n\}", rust, flags=re.S).group(1)
    r_fields = []
    for line in rbody.strip().splitlines():
        name, t = line.strip().rstrip(",").split(":", 1)
        t = t.strip()
        kind = "fnptr" if t.startswith("Option<extern") else ("ptr" if t.startswith("*") else t)
        r_fields.append((name.strip(), kind))
    assert [k for _, k in r_fields] == [k for _, k in c_fields]
    assert [n.lower() for n, _ in r_fields] == [n.lower() for n, _ in c_fields]


def test_shim_source_is_well_formed_and_exposes_the_reference_surface():
    """No rustc here: at least the delimiters balance and every public item of the reference's surface (SURVEY §8b:
    Builder / Hnsw / HnswMap / Search / Point / PointId / Item / MapItem / Heuristic) is defined."""
    src = open(RUST).read()
    text = re.sub(r"//.*", "", src)
    text = re.sub(r'"(\\.|[^"\\])*"', '""', text)
    text = re.sub(r"'(\\.|[^'\\])'", "''", text)
    stack, pairs = [], {")": "(", "]": "[", "}": "{"}
    for ch in text:
        if ch in "([{":
            stack.append(ch)
        elif ch in ")]}":
            assert stack and stack.pop() == pairs[ch], "unbalanced delimiters in the Rust shim"
    assert not stack
    for item in ("struct Builder", "struct Hnsw", "struct HnswMap", "struct Search", "trait Point", "struct PointId", "struct Item",
                 "struct MapItem", "struct Heuristic"):
        assert re.search(r"\bpub " + item + r"\b", src), item
    for method in ("fn ef_construction", "fn ef_search", "fn select_heuristic", "fn ml", "fn seed", "fn build_hnsw", "fn build<", "fn search<",
                   "fn iter", "fn builder"):
        assert "pub " + method in src, method
