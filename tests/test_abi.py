"""The C-ABI library loads and exports every symbol include/jrb200.h declares (no compute calls: no GPU here),
and the oracle's field enumeration is identical to the product's."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "jrb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from justrelax_jl_b200 import _abi

    L = _abi.lib()
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libjrb200.so does not export {n}"
    assert L.jr_abi_version() == 3


def test_field_enum_matches_oracle(oracle):
    from justrelax_jl_b200 import _abi

    assert _abi.field_names() == oracle.field_names()


def test_struct_sizes_match_header():
    # the ctypes mirrors must have the C layout (catch drift between _abi.py and jrb200.h)
    from justrelax_jl_b200 import _abi

    n = len(_abi.field_names())
    assert C.sizeof(_abi.make_fields_struct(n)) == 16 + 8 * n
    assert C.sizeof(_abi.StokesOpts) == 5 * 8 + 3 * 8 + 8 + 16 + 12 + 3 * 24 + 4 + 4 * 8 + 8 + 16


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        return
    from justrelax_jl_b200 import _abi

    h = C.c_void_p()
    st = _abi.lib().jr_context_create(0, None, C.byref(h))
    assert st == _abi.JR_ERR_CUDA
    assert b"no CPU fallback" in _abi.lib().jr_last_error()


def test_legacy_single_phase_variant_fails_loudly():
    """2D-V3 (Stokes2D.jl:345-557, solve! with a single MaterialParams) is deliberately outside the backend"""
    import pytest

    from justrelax_jl_b200 import rheology as R, stokes as jst

    class _S:
        ni = (8, 8)

    rheo = R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=1.0),
                               CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0), R.ConstantElasticity(G=1.0, ν=0.45))))
    with pytest.raises(NotImplementedError, match="legacy single-phase"):
        jst.solve_(_S(), None, None, None, None, rheo, {}, 0.1, None)
