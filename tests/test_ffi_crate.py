"""The boundary as files (ffi/): the `spgpu-sys` crate declares exactly the functions
include/spgpu.h declares and libspgpu.so exports, and ffi/gpu-feature.patch applies to the
reference crate. No Rust toolchain exists in this image, so these are the mechanical checks."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_sys_crate_is_generated_from_the_header():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_sys_crate.py"), "--check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_symbol_sets_agree():
    import gen_sys_crate as g

    handles, funcs = g.parse_header()
    header = {f[0] for f in funcs}
    rs = open(os.path.join(ROOT, "ffi", "spgpu-sys", "src", "lib.rs")).read()
    crate = set(re.findall(r"pub fn (spg_\w+)\(", rs))
    assert header == crate and len(header) > 90
    for h in handles:
        assert f"pub struct {h} " in rs
    import spartan_parallel_b200 as sp
    from spartan_parallel_b200 import _lib

    L = _lib.lib()
    assert set(sp.declared_symbols()) == header
    assert [s for s in header if not hasattr(L, s)] == []


def test_type_mapping():
    import gen_sys_crate as g

    h = ["spg_ctx", "spg_vec"]
    assert g.rust_type("const spg_fq *", h) == "*const spg_fq"
    assert g.rust_type("spg_vec **", h) == "*mut *mut spg_vec"
    assert g.rust_type("spg_vec *const *", h) == "*const *mut spg_vec"
    assert g.rust_type("const size_t *", h) == "*const usize"
    assert g.rust_type("uint8_t *", h) == "*mut u8"
    assert g.rust_type("const char *", h) == "*const c_char"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree is only present in the build container")
def test_patch_applies_to_the_reference(tmp_path):
    dst = tmp_path / "ref"
    shutil.copytree("/root/reference", dst, ignore=shutil.ignore_patterns(".git", "target"))
    patch = os.path.join(ROOT, "ffi", "gpu-feature.patch")
    subprocess.run(["git", "init", "-q"], cwd=dst, check=True)
    out = subprocess.run(["git", "apply", "--check", "--verbose", patch], cwd=dst, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    subprocess.run(["git", "apply", patch], cwd=dst, check=True)
    assert (dst / "src" / "gpu.rs").exists()
    assert "sc1_round_eval" in (dst / "src" / "sumcheck.rs").read_text()
    # the generator reproduces the committed patch from the reference
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_gpu_patch.py")], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    assert subprocess.run(["git", "diff", "--quiet", "--", "ffi/gpu-feature.patch"], cwd=ROOT).returncode == 0
