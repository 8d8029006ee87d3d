"""The compiled-language host mirror (include/sandstorm_b200.hpp: Matrix / MatrixMerkleTree over the C ABI, the shape of the
Rust binding in INTEGRATION.md) builds against the header and links against libsandstorm_b200.so.  Without a GPU the example
must fail loudly (no CPU fallback); on a B200 it must run: fused LDE == interpolate + evaluate, commit, open."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    lib_dir = os.path.join(ROOT, "sandstorm_b200")
    if not os.path.exists(os.path.join(lib_dir, "libsandstorm_b200.so")):
        import __graft_entry__ as g

        g.build()
    exe = str(tmp_path / "commit_lde")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "commit_lde.cpp"),
                           "-L", lib_dir, "-lsandstorm_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe])
    return exe


def test_example_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (covered by the gpu test)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_example_runs_on_the_gpu(tmp_path):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "lde == interpolate+evaluate: yes" in r.stdout and "opened 9 elements, 33 path nodes" in r.stdout
