#!/usr/bin/env python
"""Generates the `extern "C"` block of rust/src/ffi.rs from include/feotensor_b200.h (bindgen-style, mechanical).

    python tools/gen_rust_ffi.py            # rewrite the block between the BEGIN/END GENERATED markers in rust/src/ffi.rs
    python tools/gen_rust_ffi.py --check    # exit 1 if ffi.rs is out of date (tests/test_abi_and_host.py runs this)

There is no Rust toolchain in this image, so the crate cannot be compiled here; generating the declarations from the
header (and testing that they are current) is what keeps the two from drifting apart.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "feotensor_b200.h")
FFI = os.path.join(ROOT, "rust", "src", "ffi.rs")
BEGIN, END = "    // BEGIN GENERATED (tools/gen_rust_ffi.py)\n", "    // END GENERATED\n"

BASE = {"int": "c_int", "size_t": "usize", "uint64_t": "u64", "double": "f64", "char": "c_char", "unsigned char": "u8",
        "void": "c_void", "feo_ctx": "feo_ctx", "feo_comm": "feo_comm", "feo_nccl": "feo_nccl"}
RUST_KEYWORDS = {"in": "input", "type": "ty", "fn": "func", "ref": "reference", "mod": "modulus", "C": "c", "A": "a", "B": "b", "M": "m", "K": "k", "N": "n"}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def c_type_to_rust(ctype):
    """`const void *const *` -> `*const *const c_void`, following C's right-to-left pointer reading."""
    t = " ".join(ctype.split())
    parts = [p.strip() for p in t.split("*")]          # base [quals], then one entry per '*' holding that pointer's own qualifiers
    base_tokens = parts[0].split()
    base_const = "const" in base_tokens
    base = BASE[" ".join(tok for tok in base_tokens if tok != "const")]
    levels = []                                         # innermost first: constness of what each '*' points AT
    pointee_const = base_const
    for q in parts[1:]:
        levels.append("*const " if pointee_const else "*mut ")
        pointee_const = "const" in q.split()
    return "".join(reversed(levels)) + base


def declarations():
    text = strip_comments(open(HEADER).read())
    out = []
    for m in re.finditer(r"FEO_API\s+([^;(]*?)\b(feo_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        rargs = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
                ctype, pname = mm.group(1).strip(), mm.group(2)
                rargs.append(f"{RUST_KEYWORDS.get(pname, pname)}: {c_type_to_rust(ctype)}")
        rret = "" if ret == "void" else f" -> {c_type_to_rust(ret)}"
        out.append((name, f"    pub fn {name}({', '.join(rargs)}){rret};\n"))
    return out


def generated_block():
    return BEGIN + "".join(line for _, line in declarations()) + END


def main():
    src = open(FFI).read()
    if BEGIN not in src or END not in src:
        raise SystemExit(f"{FFI}: GENERATED markers not found")
    head, rest = src.split(BEGIN, 1)
    _, tail = rest.split(END, 1)
    new = head + generated_block() + tail
    if "--check" in sys.argv:
        if new != src:
            raise SystemExit("rust/src/ffi.rs is out of date: run python tools/gen_rust_ffi.py")
        print(f"ffi.rs is current: {len(declarations())} functions")
        return
    open(FFI, "w").write(new)
    print(f"wrote {len(declarations())} declarations to {FFI}")


if __name__ == "__main__":
    main()
