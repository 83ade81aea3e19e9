"""The C ABI driven from a plain C host (tests/abi_smoke.c: gcc + -ljrb200, no Python / torch in the process): what the Julia
extension does through ccall.  CPU leg: it compiles, links against every entry point it uses and fails loudly without a GPU;
GPU leg: SolVi3D 16^3 through jr_malloc / jr_memcpy_h2d / jr_stokes3d_solve_VA / jr_memcpy_d2h meets the reference test's
criterion (test/test_stokes_solvi3D.jl: norm_Rx[end] < 1e-8), and the iteration session runs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "justrelax_jl_b200")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-std=gnu11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_smoke.c"),
                           "-o", exe, "-L", LIBDIR, "-ljrb200", "-lm", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def test_abi_smoke_links_and_fails_loudly_without_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        return
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr, r.stderr


@pytest.mark.gpu
def test_abi_smoke_solvi3d_from_c(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stderr
    assert "norm_Rx[end]" in r.stdout
