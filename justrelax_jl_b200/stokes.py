"""Host side of the Stokes PT solvers — the Python twin of the methods a Julia
`JustRelaxB200Ext` adds to JustRelax{2,3}D (src/ext/CUDA/3D.jl:375-377, 195-218, 358-372):

    solve_(stokes, pt_stokes, grid|di, flow_bcs, ρg, K, G, dt, igg; kwargs=dict(iterMax, nout, ...))   3D-VA
    flow_bcs_(stokes, bcs), compute_maxloc_(B, A, window), velocity2displacement_(stokes, dt) ...

Python has no `!`: a trailing underscore marks the mutating functions (`solve!` → `solve_`).
Like the reference, the solver options travel in ONE keyword literally named `kwargs`
(quirk Q1, src/stokes/Stokes3D.jl:18-23).  Every function here only marshals arguments into the
C ABI of libjrb200 (include/jrb200.h); there is no numerical code and no CPU fallback on this side.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Optional

import numpy as np

from . import _abi
from . import rheology as _rheology
from .types import (require_uniform, B200BackendTrait, CPUBackendTrait, Geometry, IGG, StokesArrays, VelocityBoundaryConditions,
                    DisplacementBoundaryConditions, AbstractFlowBoundaryConditions, PhaseRatios, backend, data_ptr, is_device_array,
                    legacy_uniform_grid)

_ctx_cache = {}


def context(device: Optional[int] = None):
    """One libjrb200 context per CUDA device, running on torch's current stream of that device."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("justrelax_jl_b200 needs a CUDA device (B200); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev not in _ctx_cache:
        h = C.c_void_p()
        # torch's default stream has handle 0 = the legacy default stream; pass it as cudaStreamLegacy (0x1) so the library's
        # launches are ordered with torch's fills / copies of the same arrays (handle 0 would ask for an own stream)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _abi.check(_abi.lib().jr_context_create(dev, C.c_void_p(stream if stream else 1), C.byref(h)))
        _ctx_cache[dev] = h
    return _ctx_cache[dev]


def set_flags(flags: int, device: Optional[int] = None):
    _abi.check(_abi.lib().jr_context_set_flags(context(device), int(flags)))


def build_fields(slots: dict, ni):
    """Pack name→array into a jr_fields struct; returns (struct, keepalive)."""
    names = _abi.field_names()
    Fields = _abi.make_fields_struct(len(names))
    fs = Fields()
    fs.ndim = len(ni)
    for d in range(3):
        fs.n[d] = int(ni[d]) if d < len(ni) else 1
    for i, nm in enumerate(names):
        a = slots.get(nm)
        fs.f[i] = None if a is None else data_ptr(a)
    unknown = set(slots) - set(names)
    if unknown:
        raise KeyError(f"unknown field slots {sorted(unknown)}")
    return fs


def build_opts(pt_stokes, _di, dt, flow_bcs, n_g, *, iterMax, nout, OptsType=_abi.StokesOpts, viscosity_relaxation=1e-2,
               λ_relaxation=0.2, viscosity_cutoff=(-math.inf, math.inf), iterMin=100, strain_rate_ni_only=0, strain_increment=False):
    o = OptsType()
    o.r, o.theta_dtau, o.eta_dtau = pt_stokes.r, pt_stokes.θ_dτ, pt_stokes.ηdτ
    o.eps_rel, o.eps_abs = pt_stokes.ϵ_rel, pt_stokes.ϵ_abs
    for d in range(3):
        o._di[d] = float(_di[d]) if d < len(_di) else 0.0
        o.n_g[d] = int(n_g[d]) if d < len(n_g) else 1
    o.dt = float(dt)
    o.iterMax, o.nout = int(iterMax), int(nout)
    for name in ("free_slip", "no_slip", "periodic"):
        fl = flow_bcs.flags(name)
        arr = getattr(o, name)
        for q in range(6):
            arr[q] = fl[q]
    o.viscosity_relaxation, o.lambda_relaxation = float(viscosity_relaxation), float(λ_relaxation)
    o.visc_cutoff_lo, o.visc_cutoff_hi = float(viscosity_cutoff[0]), float(viscosity_cutoff[1])
    o.iterMin = int(iterMin)
    o.strain_rate_ni_only = int(strain_rate_ni_only)
    o.strain_increment = int(bool(strain_increment))
    o.displacement_bcs = int(isinstance(flow_bcs, DisplacementBoundaryConditions))
    return o


class _Hist:
    """history vectors of the returned NamedTuple (Stokes3D.jl:64-69)"""

    def __init__(self, cap: int, ResultType):
        self.cap = cap
        self.err_evo1 = np.zeros(cap)
        self.err_evo2 = np.zeros(cap, dtype=np.int64)
        self.norm_Rx, self.norm_Ry, self.norm_Rz, self.norm_divV = (np.zeros(cap) for _ in range(4))
        r = ResultType()
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        r.err_evo1, r.err_evo2 = dp(self.err_evo1), self.err_evo2.ctypes.data_as(C.POINTER(C.c_int64))
        r.norm_Rx, r.norm_Ry, r.norm_Rz, r.norm_divV = dp(self.norm_Rx), dp(self.norm_Ry), dp(self.norm_Rz), dp(self.norm_divV)
        self.res = r

    def named(self, ndim: int, with_time=True):
        n = int(self.res.nhist)
        out = dict(iter=int(self.res.iter), err_evo1=self.err_evo1[:n].copy(), err_evo2=self.err_evo2[:n].copy(),
                   norm_Rx=self.norm_Rx[:n].copy(), norm_Ry=self.norm_Ry[:n].copy())
        if ndim == 3:
            out["norm_Rz"] = self.norm_Rz[:n].copy()
        out["norm_divV"] = self.norm_divV[:n].copy()
        if with_time and hasattr(self.res, "time_s"):
            it = max(int(self.res.iter) - 1, 1)
            out["time"] = float(self.res.time_s)
            out["av_time"] = float(self.res.time_s) / it
            out["kernel_launches"] = int(self.res.kernel_launches)
        return SimpleNamespace(**out)


def _grid_of(stokes, grid_or_di, igg):
    if isinstance(grid_or_di, Geometry):
        require_uniform(grid_or_di, "solve!")
        return grid_or_di
    di = grid_or_di.center if hasattr(grid_or_di, "center") else grid_or_di
    if any(np.ndim(x) > 0 for x in di):
        raise NotImplementedError("solve!: vector grid spacings (non-uniform grids) are outside the B200 backend's subset")
    return legacy_uniform_grid(stokes.ni, tuple(di), igg)


def va_slots(stokes: StokesArrays, ρg, K, G) -> dict:
    d = stokes.slots()
    d["rhogx"], d["rhogy"] = ρg[0], ρg[1]
    if len(ρg) > 2:
        d["rhogz"] = ρg[2]
    d["K"], d["G"] = K, G
    return d


def solve_(stokes: StokesArrays, pt_stokes, grid, flow_bcs, ρg, *rest, kwargs=None):
    """solve!(stokes, pt_stokes, grid|di, flow_bcs, ρg, …) — every Stokes variant of the hot path, selected like the reference's
    method table by dimension and argument types:

      3D-VA  solve_(stokes, pt, grid, bcs, ρg, K, G, dt, igg)                          Stokes3D.jl:18-41
      2D-V2  solve_(stokes, pt, di,   bcs, ρg, G, K, dt, igg)                          Stokes2D.jl:181-196   (G before K in 2D)
      2D-VC  solve_(stokes, pt, di,   bcs, ρg, phase_ratios, rheology, args, dt, igg)  Stokes2D.jl:577-599
      3D-VC  solve_(stokes, pt, grid, bcs, ρg, phase_ratios, rheology, args, dt, igg)  Stokes3D.jl:447-466
    """
    if rest and isinstance(rest[0], PhaseRatios):
        if len(stokes.ni) == 2:
            return _solve2d_VC(stokes, pt_stokes, grid, flow_bcs, ρg, *rest, kwargs=kwargs)
        from .stokes3d_vc import solve3d_VC_

        return solve3d_VC_(stokes, pt_stokes, grid, flow_bcs, ρg, *rest, kwargs=kwargs)
    from .rheology import MaterialParams

    if rest and isinstance(rest[0], MaterialParams):
        # legacy single-phase VEP variant 2D-V3 (Stokes2D.jl:345-557): not built, deliberately — fail loudly
        raise NotImplementedError("solve!(stokes, pt, grid, bcs, ρg, rheology::MaterialParams, args, dt, igg) — the legacy single-phase "
                                  "variant is outside the B200 backend; pass PhaseRatios with one phase and a 1-tuple rheology (2D-VC)")
    if len(stokes.ni) == 2:
        return _solve2d_V2(stokes, pt_stokes, grid, flow_bcs, ρg, *rest, kwargs=kwargs)
    return _solve3d_VA(stokes, pt_stokes, grid, flow_bcs, ρg, *rest, kwargs=kwargs)


def _solve3d_VA(stokes: StokesArrays, pt_stokes, grid, flow_bcs, ρg, K, G, dt, igg: Optional[IGG] = None, *, kwargs=None):
    """3D visco-elastic Stokes solve with K, G arrays (variant 3D-VA).

    Reference: solve!(stokes, pt_stokes, grid::Geometry{3}|di, flow_bcs, ρg, K, G, dt, igg; kwargs)
    src/stokes/Stokes3D.jl:18-41 (entry), :25-186 (loop) — dispatched like src/ext/CUDA/3D.jl:375-377.
    Returns (iter, err_evo1, err_evo2, norm_Rx, norm_Ry, norm_Rz, norm_∇V→norm_divV, time, av_time).
    """
    kw = dict(iterMax=10e3, nout=500, b_width=(4, 4, 4), verbose=True, viscosity_relaxation=1e-2)
    kw.update(kwargs or {})
    tr = backend(stokes)
    if isinstance(tr, CPUBackendTrait):
        raise RuntimeError("solve_: StokesArrays live on the host (CPUBackend). This package only provides the B200 "
                           "backend; the CPU solver is JustRelax.jl's own. No CPU fallback.")
    if not isinstance(flow_bcs, AbstractFlowBoundaryConditions):
        raise TypeError(f"Unknown boundary conditions type: {type(flow_bcs)}")  # types/displacement.jl:68-70
    if isinstance(flow_bcs, DisplacementBoundaryConditions):
        raise NotImplementedError("DisplacementBoundaryConditions are outside the supported subset (SURVEY §8f-3)")
    for a in (*ρg, K, G):
        if not is_device_array(a):
            raise ValueError("ρg, K, G must be B200 arrays (use PTArray(B200Backend)(x))")
    igg = igg or IGG()
    grid = _grid_of(stokes, grid, igg)
    _di = grid._di.center
    opts = build_opts(pt_stokes, _di, dt, flow_bcs, igg.n_g(stokes.ni), iterMax=kw["iterMax"], nout=kw["nout"],
                      viscosity_relaxation=kw["viscosity_relaxation"])
    fs = build_fields(va_slots(stokes, ρg, K, G), stokes.ni)
    hist = _Hist(int(opts.iterMax // max(opts.nout, 1)) + 3, _abi.StokesResult)
    st = _abi.lib().jr_stokes3d_solve_VA(context(), C.byref(fs), C.byref(opts), C.byref(hist.res))
    if st == _abi.JR_ERR_NAN:
        raise RuntimeError("NaN(s)")  # Stokes3D.jl:162
    _abi.check(st)
    out = hist.named(3)
    if kw.get("verbose") and igg.me == 0:
        for c in range(len(out.err_evo1)):
            print("iter = %d, abs_err = %1.3e [norm_Rx=%1.3e, norm_Ry=%1.3e, norm_Rz=%1.3e, norm_∇V=%1.3e]" % (
                out.err_evo2[c], out.err_evo1[c], out.norm_Rx[c], out.norm_Ry[c], out.norm_Rz[c], out.norm_divV[c]))
    return out


def iterate_(stokes: StokesArrays, pt_stokes, grid, flow_bcs, ρg, K, G, dt, niter: int, igg: Optional[IGG] = None):
    """Run exactly `niter` PT iterations of variant 3D-VA (benchmark / fixed-iteration parity)."""
    igg = igg or IGG()
    grid = _grid_of(stokes, grid, igg)
    opts = build_opts(pt_stokes, grid._di.center, dt, flow_bcs, igg.n_g(stokes.ni), iterMax=niter, nout=max(niter, 1))
    fs = build_fields(va_slots(stokes, ρg, K, G), stokes.ni)
    res = _abi.StokesResult()
    _abi.check(_abi.lib().jr_stokes3d_iterate_VA(context(), C.byref(fs), C.byref(opts), int(niter), C.byref(res)))
    return SimpleNamespace(iter=int(res.iter), time=float(res.time_s), kernel_launches=int(res.kernel_launches))


class IterationSession:
    """3D-VA PT iterations with the state resident in the library's TMA box layout between calls (jr_stokes3d_VA_begin /
    _step / _end): the body of the reference's `while` loop (Stokes3D.jl:76-122) for a host that keeps its own convergence
    logic.  `step(n)` returns the device time of exactly those n iterations.

        with IterationSession(stokes, pt, grid, bcs, ρg, K, G, dt, igg) as it:
            it.step(10); r = it.step(200)       # r.time: CUDA-event seconds of the 200 iterations
    """

    def __init__(self, stokes: StokesArrays, pt_stokes, grid, flow_bcs, ρg, K, G, dt, igg: Optional[IGG] = None, *, nout=10 ** 9):
        igg = igg or IGG()
        grid = _grid_of(stokes, grid, igg)
        self._opts = build_opts(pt_stokes, grid._di.center, dt, flow_bcs, igg.n_g(stokes.ni), iterMax=10 ** 9, nout=nout)
        self._keep = va_slots(stokes, ρg, K, G)
        self._fs = build_fields(self._keep, stokes.ni)
        self._open = False

    def __enter__(self):
        _abi.check(_abi.lib().jr_stokes3d_VA_begin(context(), C.byref(self._fs), C.byref(self._opts)))
        self._open = True
        return self

    def step(self, niter: int, observe_last: bool = False):
        res = _abi.StokesResult()
        _abi.check(_abi.lib().jr_stokes3d_VA_step(context(), int(niter), int(bool(observe_last)), C.byref(res)))
        return SimpleNamespace(iter=int(res.iter), time=float(res.time_s), kernel_launches=int(res.kernel_launches))

    def __exit__(self, *exc):
        if self._open:
            self._open = False
            _abi.check(_abi.lib().jr_stokes3d_VA_end(context()))
        return False


def plan_info():
    """Facts about the fused plan of the last 3D-VA solve on this device (tile rows, z-chunks, constant-ρg elision …)."""
    info = (C.c_int32 * 8)()
    _abi.check(_abi.lib().jr_stokes3d_VA_plan_info(context(), info))
    keys = ("BY", "nchunk", "rhog_const", "finite_dt", "PX", "PY", "PZ", "slack")
    return dict(zip(keys, [int(v) for v in info]))


def flow_bcs_(stokes, bcs: AbstractFlowBoundaryConditions):
    """flow_bcs!(stokes, bcs) — src/ext/CUDA/3D.jl:195-218 → BoundaryConditions.jl:65-100."""
    if isinstance(backend(stokes), CPUBackendTrait):
        raise RuntimeError("flow_bcs_: host arrays; this package only provides the B200 backend")
    A = stokes.U if isinstance(bcs, DisplacementBoundaryConditions) else stokes.V
    comps = list(A)
    if len(stokes.ni) == 2:
        _abi.check(_abi.lib().jr_flow_bcs2d(context(), data_ptr(comps[0]), data_ptr(comps[1]), _abi.i32x(list(stokes.ni) + [1]),
                                             _abi.i32x(bcs.flags("free_slip")), _abi.i32x(bcs.flags("no_slip")),
                                             _abi.i32x(bcs.flags("periodic"))))
        return
    _abi.check(_abi.lib().jr_flow_bcs3d(context(), data_ptr(comps[0]), data_ptr(comps[1]), data_ptr(comps[2]),
                                         _abi.i32x(stokes.ni), _abi.i32x(bcs.flags("free_slip")),
                                         _abi.i32x(bcs.flags("no_slip")), _abi.i32x(bcs.flags("periodic"))))


def compute_maxloc_(B, A, window=(1, 1, 1)):
    """compute_maxloc!(B, A; window) — src/Utils.jl:409-461."""
    if not (is_device_array(A) and is_device_array(B)):
        raise RuntimeError("compute_maxloc_: B200 arrays required")
    if A.dim() != 3:
        raise NotImplementedError
    _abi.check(_abi.lib().jr_maxloc3d(context(), data_ptr(B), data_ptr(A), _abi.i32x(A.shape), _abi.i32x(window)))


def velocity2displacement_(stokes, dt):
    """velocity2displacement!(stokes, dt) — src/types/displacement.jl:1-29."""
    for u, v in zip(stokes.U, stokes.V):
        if v is not None:
            _abi.check(_abi.lib().jr_scale_copy(context(), data_ptr(u), data_ptr(v), float(dt), int(np.prod(v.shape))))
    _abi.check(_abi.lib().jr_context_synchronize(context()))


def displacement2velocity_(stokes, dt):
    """displacement2velocity!(stokes, dt) — src/types/displacement.jl:33-60 (V = U * inv(dt))."""
    for u, v in zip(stokes.U, stokes.V):
        if v is not None:
            _abi.check(_abi.lib().jr_scale_copy(context(), data_ptr(v), data_ptr(u), 1.0 / float(dt), int(np.prod(v.shape))))
    _abi.check(_abi.lib().jr_context_synchronize(context()))


def sumsq_interior(A, interior: bool = True) -> float:
    """Σ A[2:end-1,…]^2 (interior) or Σ A^2: the local part of norm_mpi (src/Utils.jl:698-701)."""
    out = C.c_double()
    shp = list(A.shape) + [1] * (3 - A.dim())
    _abi.check(_abi.lib().jr_sumsq(context(), data_ptr(A), _abi.i32x(shp), int(interior), C.byref(out)))
    return out.value


def residual_norms3d_(stokes, igg: Optional[IGG] = None):
    """the four residual norms the 3D loops sample every `nout` iterations (Stokes3D.jl:125-142): ‖R‖₂ of the interior of Rx, Ry, Rz
    and of RP (norm_mpi: all-reduced sums of squares), divided by the reference's normalisers (quirk Q4)"""
    ss = [sumsq_interior(stokes.R.Rx), sumsq_interior(stokes.R.Ry), sumsq_interior(stokes.R.Rz), sumsq_interior(stokes.R.RP, False)]
    if igg is not None and igg.nprocs > 1:
        from .comm import _allreduce

        ss = _allreduce(ss, 0)
    gx, gy, gz = (igg or IGG()).n_g(stokes.ni)
    den = ((gx - 2) * (gy - 1) * (gz - 1), (gx - 1) * (gy - 2) * (gz - 1), (gx - 1) * (gy - 1) * (gz - 2), gx * gy * gz)
    return tuple(math.sqrt(v) / d for v, d in zip(ss, den))


def norm_interior(A, interior: bool = True) -> float:
    return math.sqrt(sumsq_interior(A, interior))


# ------------------------------------------------------------------------------------------------------------------
# 2D variants
def _common_checks(stokes, flow_bcs, arrays, *, allow_displacement=False):
    if isinstance(backend(stokes), CPUBackendTrait):
        raise RuntimeError("solve_: StokesArrays live on the host (CPUBackend). This package only provides the B200 "
                           "backend; the CPU solver is JustRelax.jl's own. No CPU fallback.")
    if not isinstance(flow_bcs, AbstractFlowBoundaryConditions):
        raise TypeError(f"Unknown boundary conditions type: {type(flow_bcs)}")
    if isinstance(flow_bcs, DisplacementBoundaryConditions) and not allow_displacement:
        raise NotImplementedError("DisplacementBoundaryConditions are supported by the multiphase 2D solve (2D-VC) only")
    for a in arrays:
        if a is not None and not is_device_array(a):
            raise ValueError("array arguments must be B200 arrays (use PTArray(B200Backend)(x))")


def _single_rank2d(igg):
    """the 2D solvers have no halo exchange / all-reduced norms (libjrb200 returns JR_ERR_UNSUPPORTED too): refuse loudly
    instead of returning rank-local answers"""
    if igg is not None and (igg.nprocs > 1 or any(getattr(igg, "periods", (0, 0, 0)))):
        raise NotImplementedError(f"the 2D Stokes solvers of the B200 backend run on one non-periodic rank only (igg.nprocs = {igg.nprocs}, "
                                  f"periods = {tuple(getattr(igg, 'periods', (0, 0, 0)))})")


def _print_hist2(out, igg, verbose):
    if verbose and igg.me == 0:
        for c in range(len(out.err_evo1)):
            print("Iteration = %d, err = %1.3e [norm_Rx=%1.3e, norm_Ry=%1.3e, norm_∇V=%1.3e]" % (
                out.err_evo2[c], out.err_evo1[c], out.norm_Rx[c], out.norm_Ry[c], out.norm_divV[c]))


def _solve2d_V2(stokes, pt_stokes, di, flow_bcs, ρg, G, K, dt, igg: Optional[IGG] = None, *, kwargs=None):
    """2D visco-elastic solve with G, K arrays (variant 2D-V2) — src/stokes/Stokes2D.jl:181-325."""
    kw = dict(iterMax=10e3, nout=500, b_width=(4, 4, 1), verbose=True)
    kw.update(kwargs or {})
    _common_checks(stokes, flow_bcs, (*ρg, G, K))
    igg = igg or IGG()
    _single_rank2d(igg)
    grid = _grid_of(stokes, di, igg)
    opts = build_opts(pt_stokes, grid._di.center, dt, flow_bcs, igg.n_g(stokes.ni), iterMax=kw["iterMax"], nout=kw["nout"])
    fs = build_fields(va_slots(stokes, ρg, K, G), stokes.ni)
    hist = _Hist(int(opts.iterMax // max(opts.nout, 1)) + 3, _abi.StokesResult)
    st = _abi.lib().jr_stokes2d_solve_V2(context(), C.byref(fs), C.byref(opts), C.byref(hist.res))
    if st == _abi.JR_ERR_NAN:
        raise RuntimeError("NaN(s)")
    _abi.check(st)
    out = hist.named(2)
    _print_hist2(out, igg, kw.get("verbose"))
    return out


def iterate2d_V2_(stokes, pt_stokes, di, flow_bcs, ρg, G, K, dt, niter: int, igg: Optional[IGG] = None):
    """exactly `niter` PT iterations of variant 2D-V2 (benchmark / fixed-iteration parity)"""
    igg = igg or IGG()
    _single_rank2d(igg)
    grid = _grid_of(stokes, di, igg)
    opts = build_opts(pt_stokes, grid._di.center, dt, flow_bcs, igg.n_g(stokes.ni), iterMax=niter, nout=max(niter, 1))
    fs = build_fields(va_slots(stokes, ρg, K, G), stokes.ni)
    res = _abi.StokesResult()
    _abi.check(_abi.lib().jr_stokes2d_iterate_V2(context(), C.byref(fs), C.byref(opts), int(niter), C.byref(res)))
    return SimpleNamespace(iter=int(res.iter), time=float(res.time_s), kernel_launches=int(res.kernel_launches))


def vc_inputs(rheology, phase_ratios: PhaseRatios, *, free_surface: float = 0.0, ndim: int = 2):
    """lower rheology::NTuple{N,MaterialParams} to the flat table (raises UnsupportedRheology outside the subset) and bind the
    phase-ratio arrays → jr_vc_inputs"""
    rows = _rheology.lower_stokes(rheology, ndim)
    arr = (_abi.StokesPhase * len(rows))()
    for i, r in enumerate(rows):
        for k, v in r.items():
            setattr(arr[i], k, (C.c_double * len(v))(*v) if isinstance(v, (list, tuple)) else v)
    vc = _abi.VcInputs()
    # compute_gravity(ConstantGravity) is a Number: only the last ρg component is filled (BuoyancyForces.jl:70-71)
    vc.nphase, vc.g_scalar, vc.phases, vc.free_surface = len(rows), 1, arr, float(free_surface)
    g = _rheology.gravity_of(rheology)
    for q in range(3):
        vc.g[q] = float(g[q])
    for nm in ("center", "vertex", "xy", "yz", "xz"):
        a = getattr(phase_ratios, nm, None)
        if a is not None:
            if not is_device_array(a):
                raise ValueError("phase ratios must be B200 arrays")
            if a.shape[-1] != len(rows):
                raise ValueError(f"phase_ratios.{nm} holds {a.shape[-1]} phases, rheology has {len(rows)}")
            setattr(vc, "ph_" + nm, data_ptr(a))
    vc._keep = (arr, phase_ratios)
    return vc


def _args_items(args) -> dict:
    if args is None:
        return {}
    if isinstance(args, dict):
        return dict(args)
    if hasattr(args, "_asdict"):
        return dict(args._asdict())
    return dict(vars(args))


def vc_slots(stokes, ρg, args) -> dict:
    """field slots of a VC solve.  `args` is the reference's NamedTuple (Stokes2D.jl:577-599, Stokes3D.jl:447-466): T (ni.+2, cell
    centres with one ghost layer) and P feed the density / viscosity laws, ΔT (ni) switches compute_P! to the thermal-stress form
    (PressureKernels.jl:128-149,197-206), dt is carried by the miniapps but never read by the lowered laws, and perturbation_C is
    accepted and unused exactly as in the reference (its only consumer, the keyword of plastic_params_phase StressUpdate.jl:146-176,
    is never handed `args` by the stress kernels of this version: StressKernels.jl:1037,1084 call it without keywords).  Every other
    key — melt_fraction (PressureKernels.jl:151-176), ϕ, … — selects behaviour this backend does not have: it raises instead of
    being dropped."""
    d = stokes.slots()
    d["rhogx"], d["rhogy"] = ρg[0], ρg[1]
    if len(ρg) > 2:
        d["rhogz"] = ρg[2]
    items = _args_items(args)
    known = {"T": "T", "P": "Pargs", "ΔT": "dTargs", "dT": "dTargs"}
    for k, v in items.items():
        if k in ("dt", "perturbation_C"):
            continue
        if k not in known:
            raise NotImplementedError(f"args.{k} is not supported by the B200 backend (supported keys: T, P, ΔT, dt, perturbation_C); "
                                      "refusing to ignore it")
        if v is None:
            continue
        if not is_device_array(v):
            raise ValueError(f"args.{k} must be a B200 array (use PTArray(B200Backend)(x))")
        d[known[k]] = v
    if "dTargs" in d and tuple(d["dTargs"].shape) not in (tuple(stokes.ni), tuple(n + 2 for n in stokes.ni)):
        raise ValueError(f"args.ΔT must have the size of the cell grid {tuple(stokes.ni)} or of thermal.ΔT (ni .+ 2), got {tuple(d['dTargs'].shape)}")
    return d


def _dT_ghosted(stokes, slots) -> bool:
    """args.ΔT given as thermal.ΔT (ni .+ 2), the way the reference's scripts pass it (test_thermalstresses.jl:321, Blob3D.jl:280): the kernel
    then indexes it ΔT[I...] without an offset, exactly like compute_P_kernel! (PressureKernels.jl:143-146)"""
    a = slots.get("dTargs")
    return a is not None and tuple(a.shape) == tuple(n + 2 for n in stokes.ni)


def _vc_opts(stokes, pt_stokes, grid, flow_bcs, dt, igg, kw):
    return build_opts(pt_stokes, grid._di.center, dt, flow_bcs, igg.n_g(stokes.ni), iterMax=kw["iterMax"], nout=kw["nout"],
                      viscosity_relaxation=kw["viscosity_relaxation"], λ_relaxation=kw["λ_relaxation"],
                      viscosity_cutoff=kw["viscosity_cutoff"], iterMin=kw["iterMin"], strain_increment=kw.get("strain_increment", False))


def _solve2d_VC(stokes, pt_stokes, di, flow_bcs, ρg, phase_ratios, rheology, args, dt, igg: Optional[IGG] = None, *, kwargs=None):
    """2D multiphase visco-elasto-plastic solve (variant 2D-VC) — src/stokes/Stokes2D.jl:577-866."""
    kw = dict(iterMax=50e3, iterMin=1e2, viscosity_relaxation=1e-2, λ_relaxation=0.2, free_surface=False, nout=500, b_width=(4, 4, 0),
              verbose=True, viscosity_cutoff=(-math.inf, math.inf), strain_increment=False)
    kw.update(kwargs or {})
    _common_checks(stokes, flow_bcs, ρg, allow_displacement=True)
    igg = igg or IGG()
    _single_rank2d(igg)
    grid = _grid_of(stokes, di, igg)
    opts = _vc_opts(stokes, pt_stokes, grid, flow_bcs, dt, igg, kw)
    vc = vc_inputs(rheology, phase_ratios, free_surface=float(dt) * float(kw["free_surface"]) if kw["free_surface"] else 0.0)
    slots = vc_slots(stokes, ρg, args)
    opts.dT_ghosted = int(_dT_ghosted(stokes, slots))
    fs = build_fields(slots, stokes.ni)
    hist = _Hist(int(opts.iterMax // max(opts.nout, 1)) + 3, _abi.StokesResult)
    st = _abi.lib().jr_stokes2d_solve_VC(context(), C.byref(fs), C.byref(opts), C.byref(vc), C.byref(hist.res))
    if st == _abi.JR_ERR_NAN:
        raise RuntimeError("NaN(s)")  # Stokes2D.jl:836
    _abi.check(st)
    out = hist.named(2)
    _print_hist2(out, igg, kw.get("verbose"))
    return out


def iterate2d_VC_(stokes, pt_stokes, di, flow_bcs, ρg, phase_ratios, rheology, args, dt, niter: int, igg: Optional[IGG] = None, *, finish=False,
                  kwargs=None):
    """pre-loop initialisation + exactly `niter` iterations of variant 2D-VC (+ the exit kernels when finish)"""
    kw = dict(iterMax=niter, iterMin=0, viscosity_relaxation=1e-2, λ_relaxation=0.2, free_surface=False, nout=max(niter, 1),
              viscosity_cutoff=(-math.inf, math.inf), strain_increment=False)
    kw.update(kwargs or {})
    igg = igg or IGG()
    _single_rank2d(igg)
    grid = _grid_of(stokes, di, igg)
    opts = _vc_opts(stokes, pt_stokes, grid, flow_bcs, dt, igg, kw)
    vc = vc_inputs(rheology, phase_ratios, free_surface=float(dt) * float(kw["free_surface"]) if kw["free_surface"] else 0.0)
    slots = vc_slots(stokes, ρg, args)
    opts.dT_ghosted = int(_dT_ghosted(stokes, slots))
    fs = build_fields(slots, stokes.ni)
    res = _abi.StokesResult()
    _abi.check(_abi.lib().jr_stokes2d_iterate_VC(context(), C.byref(fs), C.byref(opts), C.byref(vc), int(niter), int(finish), C.byref(res)))
    return SimpleNamespace(iter=int(res.iter), time=float(res.time_s), kernel_launches=int(res.kernel_launches))


def compute_viscosity_(stokes, phase_ratios, args, rheology, cutoff=(-math.inf, math.inf), *, relaxation=1.0):
    """compute_viscosity!(stokes, phase_ratios, args, rheology, cutoff; relaxation) — src/rheology/Viscosity.jl:67-106."""
    if len(stokes.ni) != 2:
        from .stokes3d_vc import compute_viscosity3d_

        return compute_viscosity3d_(stokes, phase_ratios, args, rheology, cutoff, relaxation=relaxation)
    vc = vc_inputs(rheology, phase_ratios)
    o = _abi.StokesOpts()
    o.visc_cutoff_lo, o.visc_cutoff_hi = float(cutoff[0]), float(cutoff[1])
    fs = build_fields(vc_slots(stokes, (stokes.P, stokes.P), args), stokes.ni)
    _abi.check(_abi.lib().jr_compute_viscosity2d(context(), C.byref(fs), C.byref(o), C.byref(vc), float(relaxation)))


def compute_ρg_(ρg, phase_ratios, rheology, args, stokes):
    """compute_ρg!(ρg, phase_ratios, rheology, args) — src/rheology/BuoyancyForces.jl:74-95 (stokes only supplies the grid size)."""
    if len(stokes.ni) != 2:
        from .stokes3d_vc import compute_rhog3d_

        return compute_rhog3d_(ρg, phase_ratios, rheology, args, stokes)
    vc = vc_inputs(rheology, phase_ratios)
    fs = build_fields(vc_slots(stokes, ρg, args), stokes.ni)
    _abi.check(_abi.lib().jr_compute_rhog2d(context(), C.byref(fs), C.byref(vc)))


def tensor_invariant_(T, ni):
    """tensor_invariant!(A::SymmetricTensor) — src/stokes/StressKernels.jl:442-480 (2D)."""
    if len(ni) != 2:
        from .stokes3d_vc import tensor_invariant3d_

        return tensor_invariant3d_(T, ni)
    _abi.check(_abi.lib().jr_tensor_invariant2d(context(), data_ptr(T.II), data_ptr(T.xx), data_ptr(T.yy), data_ptr(T.xy),
                                                 _abi.i32x(list(ni) + [1])))


def accumulate_tensor_(II, A, dt, ni):
    """accumulate_tensor!(II, A::SymmetricTensor, dt) — src/stokes/StressKernels.jl:364-408 (II += second_invariant_staggered(A) · dt)."""
    n3 = _abi.i32x(list(ni) + [1] * (3 - len(ni)))
    if len(ni) == 2:
        _abi.check(_abi.lib().jr_accumulate_tensor2d(context(), data_ptr(II), data_ptr(A.xx), data_ptr(A.yy), data_ptr(A.xy), n3, float(dt)))
    else:
        _abi.check(_abi.lib().jr_accumulate_tensor3d(context(), data_ptr(II), data_ptr(A.xx), data_ptr(A.yy), data_ptr(A.zz), data_ptr(A.yz),
                                                      data_ptr(A.xz), data_ptr(A.xy), n3, float(dt)))


def accumulate_vol_(EVol_pl, ε_vol_pl, dt):
    """accumulate_vol!(EVol_pl, ε_vol_pl, dt) — src/stokes/StressKernels.jl:422-438."""
    _abi.check(_abi.lib().jr_accumulate_vol(context(), data_ptr(EVol_pl), data_ptr(ε_vol_pl), int(np.prod(EVol_pl.shape)), float(dt)))


def compute_dt_(stokes, di, dt_diff=math.inf, igg: Optional[IGG] = None):
    """compute_dt(stokes, di[, dt_diff][, igg]) — src/Utils.jl:492-519: min(dt_diff, 0.9 · min_d(di[d] · inv(maximum(abs.(V_d))))), with
    maximum_mpi when an IGG is given."""
    best = math.inf
    for d, v in zip(di, stokes.V):
        if v is None:
            continue
        out = C.c_double()
        _abi.check(_abi.lib().jr_absmax(context(), data_ptr(v), int(np.prod(v.shape)), int(igg is not None and igg.nprocs > 1), C.byref(out)))
        m = out.value
        x = float(d) * (math.inf if m == 0.0 else 1.0 / m)
        best = x if math.isnan(x) else min(best, x)   # mapreduce(min): NaN propagates
    return min(float(dt_diff), best * 0.9)


# ------------------------------------------------------------------------------------------------------------------
# per-time-step kernels between the Stokes and the thermal loops
def _n3(ni):
    return _abi.i32x(list(ni) + [1] * (3 - len(ni)))


def velocity2vertex_(Vv, V, ni):
    """velocity2vertex!(Vx_v, Vy_v[, Vz_v], Vx, Vy[, Vz]) — src/Interpolations.jl:212-248; Vv / V: sequences of B200 arrays"""
    nd = len(ni)
    z = lambda seq, q: data_ptr(seq[q]) if nd == 3 else None
    _abi.check(_abi.lib().jr_velocity2vertex(context(), nd, _n3(ni), _n3(Vv[0].shape), data_ptr(Vv[0]), data_ptr(Vv[1]), z(Vv, 2),
                                              data_ptr(V[0]), data_ptr(V[1]), z(V, 2)))


def velocity2center_(Vc, V, ni):
    """velocity2center!(Vx_c, Vy_c[, Vz_c], Vx, Vy[, Vz]) — src/Interpolations.jl:257-289"""
    nd = len(ni)
    z = lambda seq, q: data_ptr(seq[q]) if nd == 3 else None
    _abi.check(_abi.lib().jr_velocity2center(context(), nd, _n3(ni), _n3(Vc[0].shape), data_ptr(Vc[0]), data_ptr(Vc[1]), z(Vc, 2),
                                              data_ptr(V[0]), data_ptr(V[1]), z(V, 2)))


def compute_lithostatic_pressure_(P, ρg, dz, igg: Optional[IGG] = None):
    """compute_lithostatic_pressure!(P, ρg, dz[, igg]) — src/Utils.jl:541-617.  dz: a number or a B200 vector of cell heights."""
    if tuple(P.shape) != tuple(ρg.shape):
        raise ValueError(f"`P` and `ρg` must span the same cells, got axes {tuple(P.shape)} and {tuple(ρg.shape)}")   # DimensionMismatch
    nd = P.dim()
    if igg is not None and igg.dims[nd - 1] > 1 and igg.periods[nd - 1]:   # Utils.jl:580-583 (test_lithostatic_pressure2D_MPI.jl:138-141)
        raise RuntimeError("the lithostatic pressure of a column that is periodic along the vertical direction is undefined")
    dzv = None
    if not isinstance(dz, (int, float)):
        if int(np.prod(dz.shape)) != P.shape[-1]:
            raise ValueError(f"`dz` must hold one height per cell, got {int(np.prod(dz.shape))} heights for {P.shape[-1]} cells")
        dzv = data_ptr(dz)
    across = igg is not None
    nvert = int(igg.nxyz[nd - 1]) if (across and igg.nxyz is not None) else int(P.shape[-1])
    st = _abi.lib().jr_lithostatic_pressure(context(), nd, _n3(P.shape), data_ptr(P), data_ptr(ρg), 0.0 if dzv else float(dz), dzv, int(across), nvert)
    if st == _abi.JR_ERR_ARG:
        raise ValueError(_abi.lib().jr_last_error().decode())   # ArgumentError("the vertical direction is split across MPI ranks; …")
    _abi.check(st)
    _abi.check(_abi.lib().jr_context_synchronize(context()))
    return P


def compute_shear_heating_(thermal, stokes, *rest):
    """compute_shear_heating!(thermal, stokes, rheology, dt) / (thermal, stokes, phase_ratios, rheology, dt)
    — src/thermal_diffusion/ShearHeating.jl:14-72"""
    if len(rest) == 3:
        phase_ratios, rheology, dt = rest
    else:
        (rheology, dt), phase_ratios = rest, None
    rh = tuple(rheology) if isinstance(rheology, (tuple, list)) else (rheology,)
    if phase_ratios is None and len(rh) != 1:
        raise ValueError("a multi-phase rheology needs the phase ratios")
    rows = _rheology.lower_stokes(rh)
    arr = (_abi.StokesPhase * len(rows))()
    for i, r in enumerate(rows):
        for k, v in r.items():
            setattr(arr[i], k, (C.c_double * len(v))(*v) if isinstance(v, (list, tuple)) else v)
    vc = _abi.VcInputs()
    vc.nphase, vc.g_scalar, vc.phases = len(rows), 1, arr
    if phase_ratios is not None:
        vc.ph_center = data_ptr(phase_ratios.center)
    chi = (C.c_double * len(rows))(*_rheology.shear_heating_coefficients(rh))
    fs = build_fields(stokes.slots(), stokes.ni)
    _abi.check(_abi.lib().jr_compute_shear_heating(context(), C.byref(fs), C.byref(vc), chi, float(dt), data_ptr(thermal.shear_heating)))
