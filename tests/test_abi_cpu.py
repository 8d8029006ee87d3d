"""CPU-only: the C-ABI library loads, exports every symbol include/sandstorm_b200.h declares,
and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    if not os.path.exists(os.path.join(ROOT, "sandstorm_b200", "libsandstorm_b200.so")):
        g.build()
    import sandstorm_b200._lib as L

    return L


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sandstorm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    cdll = lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(cdll, name), f"{name} declared in the header but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(lib.SIGNATURES) == names


def test_no_cpu_fallback_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = ctypes.c_void_p()
    assert lib.load().ss_create(0, ctypes.byref(handle)) == lib.SS_ERR_CUDA
    from sandstorm_b200.context import Context

    with pytest.raises(lib.SandstormError):
        Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sandstorm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/sandstorm_b200.h must compile as C99 (what cgo / bindgen / a Rust build.rs
    feed to their C front end) as well as C++, with no CUDA or torch types in the signatures."""
    import subprocess

    src = tmp_path / "hdr.c"
    src.write_text('#include "sandstorm_b200.h"\nint main(void) { ss_ctx *c = 0; (void)c; return (int)SS_OK; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
    text = open(os.path.join(inc, "sandstorm_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert not re.search(r"cudaStream_t|cudaError_t|#include\s*<cuda|at::|torch::", code), "CUDA / torch types leak into the C ABI"
