"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol
include/spgpu.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import os
import subprocess

import pytest

import spartan_parallel_b200 as sp
from spartan_parallel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", ROOT, "-j8"])
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    decl = sp.declared_symbols()
    assert len(decl) >= 50
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, f"libspgpu.so does not export {missing}"


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "spgpu.h"\nint main(void){ spg_fq x; (void)x; return SPG_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sp.SpgError) as e:
        sp.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spartan_parallel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace("oracle/.", ""), f
