"""Shared helpers for the parity tests (host dict of ABI slots <-> B200 StokesArrays)."""
import numpy as np


def bc_flags(flow_bcs):
    return dict(free_slip=flow_bcs.flags("free_slip"), no_slip=flow_bcs.flags("no_slip"), periodic=flow_bcs.flags("periodic"))


def device_stokes(ni, host: dict):
    """StokesArrays on the B200 filled from a host slot dict; returns (stokes, extra device arrays)."""
    from justrelax_jl_b200 import B200Backend, PTArray, StokesArrays

    st = StokesArrays(B200Backend, *ni, vertex_normals=False)
    sl = st.slots()
    extra = {}
    for k, a in host.items():
        if k in sl and sl[k] is not None:
            sl[k].copy_(PTArray(B200Backend)(a))
        else:
            extra[k] = PTArray(B200Backend)(a)
    return st, extra


def max_rel_diff(a, b):
    """per-field max |a-b| / max|b| (field-scale relative difference, the north-star metric)."""
    a, b = np.asarray(a), np.asarray(b)
    fin = np.isfinite(b)
    if not np.array_equal(np.isfinite(a), fin):
        return np.inf
    if not fin.all() and not np.array_equal(a[~fin], b[~fin], equal_nan=True):
        return np.inf
    if not fin.any():
        return 0.0
    scale = np.max(np.abs(b[fin]))
    d = np.max(np.abs(a[fin] - b[fin]))
    return 0.0 if d == 0 else d / (scale if scale > 0 else 1.0)


def compare_slots(dev_slots: dict, host_slots: dict, names, tol, label=""):
    from justrelax_jl_b200 import to_host

    worst = {}
    for nm in names:
        got = to_host(dev_slots[nm])
        r = max_rel_diff(got, host_slots[nm])
        worst[nm] = r
    bad = {k: v for k, v in worst.items() if not (v <= tol)}
    assert not bad, f"{label} fields over tol {tol}: {bad}"
    return worst
