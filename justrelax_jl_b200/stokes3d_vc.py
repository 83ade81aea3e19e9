"""Host mirror of the 3D multiphase visco-elasto-plastic Stokes solve (variant 3D-VC) and of the stand-alone 3D kernels the
reference calls either side of it — only marshalling, no numerics, no CPU fallback.

Reference: solve!(stokes, pt_stokes, grid|di, flow_bcs, ρg, phase_ratios, rheology, args, dt, igg; kwargs)
src/stokes/Stokes3D.jl:447-668, dispatched like src/ext/CUDA/3D.jl:375-377; compute_viscosity! (:231-263),
compute_ρg! (:287-299), tensor_invariant! (:266-272), shear2center! (:319-327).
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Optional

from . import _abi
from .stokes import (_Hist, _common_checks, _dT_ghosted, _grid_of, _vc_opts, build_fields, context, vc_inputs, vc_slots)
from .types import IGG, data_ptr


def solve3d_VC_(stokes, pt_stokes, grid, flow_bcs, ρg, phase_ratios, rheology, args, dt, igg: Optional[IGG] = None, *, kwargs=None):
    """returns (iter, err_evo1, err_evo2, norm_Rx, norm_Ry, norm_Rz, norm_∇V→norm_divV, time, av_time)  Stokes3D.jl:657-667"""
    kw = dict(iterMax=10e3, nout=500, b_width=(4, 4, 4), verbose=True, viscosity_relaxation=1e-2, λ_relaxation=0.2,
              viscosity_cutoff=(-math.inf, math.inf), iterMin=0)
    kw.update(kwargs or {})
    _common_checks(stokes, flow_bcs, ρg)
    igg = igg or IGG()
    grid = _grid_of(stokes, grid, igg)
    opts = _vc_opts(stokes, pt_stokes, grid, flow_bcs, dt, igg, kw)
    vc = vc_inputs(rheology, phase_ratios, ndim=3)
    slots = vc_slots(stokes, ρg, args)
    opts.dT_ghosted = int(_dT_ghosted(stokes, slots))
    fs = build_fields(slots, stokes.ni)
    hist = _Hist(int(opts.iterMax // max(opts.nout, 1)) + 3, _abi.StokesResult)
    st = _abi.lib().jr_stokes3d_solve_VC(context(), C.byref(fs), C.byref(opts), C.byref(vc), C.byref(hist.res))
    if st == _abi.JR_ERR_NAN:
        raise RuntimeError("NaN(s)")  # Stokes3D.jl:631
    _abi.check(st)
    out = hist.named(3)
    if kw.get("verbose") and igg.me == 0:
        for c in range(len(out.err_evo1)):
            print("iter = %d, abs_err = %1.3e [norm_Rx=%1.3e, norm_Ry=%1.3e, norm_Rz=%1.3e, norm_∇V=%1.3e]" % (
                out.err_evo2[c], out.err_evo1[c], out.norm_Rx[c], out.norm_Ry[c], out.norm_Rz[c], out.norm_divV[c]))
    return out


def iterate3d_VC_(stokes, pt_stokes, grid, flow_bcs, ρg, phase_ratios, rheology, args, dt, niter: int, igg: Optional[IGG] = None, *,
                  finish=False, kwargs=None):
    """pre-loop initialisation + exactly `niter` iterations of variant 3D-VC (+ the exit kernels when finish)"""
    kw = dict(iterMax=niter, iterMin=0, viscosity_relaxation=1e-2, λ_relaxation=0.2, nout=max(niter, 1), viscosity_cutoff=(-math.inf, math.inf))
    kw.update(kwargs or {})
    igg = igg or IGG()
    grid = _grid_of(stokes, grid, igg)
    opts = _vc_opts(stokes, pt_stokes, grid, flow_bcs, dt, igg, kw)
    vc = vc_inputs(rheology, phase_ratios, ndim=3)
    slots = vc_slots(stokes, ρg, args)
    opts.dT_ghosted = int(_dT_ghosted(stokes, slots))
    fs = build_fields(slots, stokes.ni)
    res = _abi.StokesResult()
    _abi.check(_abi.lib().jr_stokes3d_iterate_VC(context(), C.byref(fs), C.byref(opts), C.byref(vc), int(niter), int(finish), C.byref(res)))
    return SimpleNamespace(iter=int(res.iter), time=float(res.time_s), kernel_launches=int(res.kernel_launches))


def compute_viscosity3d_(stokes, phase_ratios, args, rheology, cutoff=(-math.inf, math.inf), *, relaxation=1.0):
    vc = vc_inputs(rheology, phase_ratios, ndim=3)
    o = _abi.StokesOpts()
    o.visc_cutoff_lo, o.visc_cutoff_hi = float(cutoff[0]), float(cutoff[1])
    fs = build_fields(vc_slots(stokes, (stokes.P, stokes.P, stokes.P), args), stokes.ni)
    _abi.check(_abi.lib().jr_compute_viscosity3d(context(), C.byref(fs), C.byref(o), C.byref(vc), float(relaxation)))


def compute_rhog3d_(ρg, phase_ratios, rheology, args, stokes):
    vc = vc_inputs(rheology, phase_ratios, ndim=3)
    fs = build_fields(vc_slots(stokes, ρg, args), stokes.ni)
    _abi.check(_abi.lib().jr_compute_rhog3d(context(), C.byref(fs), C.byref(vc)))


def tensor_invariant3d_(T, ni):
    """tensor_invariant!(A::SymmetricTensor) 3D — II at the centres from the staggered tensor"""
    _abi.check(_abi.lib().jr_tensor_invariant3d(context(), data_ptr(T.II), data_ptr(T.xx), data_ptr(T.yy), data_ptr(T.zz), data_ptr(T.yz),
                                                 data_ptr(T.xz), data_ptr(T.xy), _abi.i32x(list(ni))))


def shear2center3d_(T, ni):
    """shear2center!(A::SymmetricTensor) 3D"""
    _abi.check(_abi.lib().jr_shear2center3d(context(), data_ptr(T.yz_c), data_ptr(T.xz_c), data_ptr(T.xy_c), data_ptr(T.yz), data_ptr(T.xz),
                                             data_ptr(T.xy), _abi.i32x(list(ni))))
