#!/usr/bin/env python
"""Generates ffi/spgpu-sys/src/lib.rs from include/spgpu.h: one `extern "C"` declaration per
function the header declares, the opaque handle types, spg_fq and the status codes.

  python tools/gen_sys_crate.py            # rewrite ffi/spgpu-sys/src/lib.rs
  python tools/gen_sys_crate.py --check    # exit 1 if the committed file is stale

tests/test_ffi_crate.py runs the --check and compares the symbol sets of the header, the crate
and libspgpu.so."""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spgpu.h")
OUT = os.path.join(ROOT, "ffi", "spgpu-sys", "src", "lib.rs")

SCALARS = {"int": "c_int", "size_t": "usize", "uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "char": "c_char",
           "void": "c_void", "double": "f64"}
RUST_KEYWORDS = {"in", "ref", "type", "box", "move", "fn", "mod", "use", "loop", "match", "self"}


def strip_comments(src: str) -> str:
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def parse_header(path: str = HEADER):
    """-> (handles, functions); a function is (name, return C type, [(C type, name)])."""
    raw = open(path).read()
    src = strip_comments(raw)
    handles = re.findall(r"typedef struct (spg_\w+) \1;", src)
    funcs = []
    for m in re.finditer(r"^([A-Za-z_][\w \*]*?)\b(spg_\w+)\(([^;{]*?)\);", src, flags=re.M | re.S):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        plist = []
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                arr = re.match(r"(.*?)(\w+)\[(\d*)\]$", p)
                if arr:  # T name[N] decays to T *name
                    plist.append((arr.group(1).strip() + " *", arr.group(2)))
                    continue
                pm = re.match(r"(.*?)(\w+)$", p)
                plist.append((pm.group(1).strip(), pm.group(2)))
        funcs.append((name, ret, plist))
    return handles, funcs


def rust_type(c: str, handles) -> str:
    """C declarator (without the name) -> Rust FFI type."""
    c = " ".join(c.replace("*", " * ").split())
    toks = c.split()
    # base type: leading [const] T
    const_base = False
    if toks and toks[0] == "const":
        const_base = True
        toks = toks[1:]
    base, toks = toks[0], toks[1:]
    if base in SCALARS:
        ty = SCALARS[base]
    elif base == "spg_fq" or base in handles:
        ty = base
    else:
        raise ValueError(f"unknown C type {base!r} in {c!r}")
    pointee_const = const_base
    i = 0
    while i < len(toks):
        assert toks[i] == "*", c
        ty = ("*const " if pointee_const else "*mut ") + ty
        pointee_const = False
        i += 1
        if i < len(toks) and toks[i] == "const":
            pointee_const = True
            i += 1
    return ty


def render() -> str:
    handles, funcs = parse_header()
    src = strip_comments(open(HEADER).read())
    codes = re.findall(r"(SPG_\w+) = (-?\d+)", src)
    out = []
    out.append("//! Raw bindings to `libspgpu.so`, the B200 prover backend for spartan-parallel's data-parallel\n"
               "//! R1CS proving path. GENERATED from `include/spgpu.h` by `tools/gen_sys_crate.py` -- do not edit;\n"
               "//! the header documents every function (reference file:line each one replaces).\n"
               "#![allow(non_camel_case_types, non_snake_case)]\n"
               "#![no_std]\n\n"
               "use core::ffi::{c_char, c_int, c_void};\n")
    out.append("/// The reference's `Scalar`: four little-endian u64 limbs of a * 2^256 mod q, fully reduced\n"
               "/// (`src/scalar/ristretto255.rs:193-199`). Layout-compatible with `Scalar(pub(crate) [u64; 4])`.\n"
               "#[repr(C)]\n#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]\npub struct spg_fq {\n    pub l: [u64; 4],\n}\n")
    for h in handles:
        out.append(f"#[repr(C)]\npub struct {h} {{\n    _opaque: [u8; 0],\n}}")
    out.append("")
    for name, val in codes:
        out.append(f"pub const {name}: c_int = {val};")
    out.append("\nextern \"C\" {")
    for name, ret, plist in funcs:
        args = []
        for cty, pname in plist:
            if pname in RUST_KEYWORDS:
                pname += "_"
            args.append(f"{pname}: {rust_type(cty, handles)}")
        r = "" if ret == "void" else f" -> {rust_type(ret, handles)}"
        line = f"    pub fn {name}({', '.join(args)}){r};"
        if len(line) > 110:
            line = f"    pub fn {name}(\n" + "".join(f"        {a},\n" for a in args) + f"    ){r};"
        out.append(line)
    out.append("}")
    return "\n".join(out) + "\n"


def main():
    text = render()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            print(f"{OUT} is stale: run python tools/gen_sys_crate.py", file=sys.stderr)
            sys.exit(1)
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print(f"wrote {OUT}: {text.count('pub fn ')} functions")


if __name__ == "__main__":
    main()
