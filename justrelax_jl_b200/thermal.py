"""Host side of the thermal PT solver — the Python twin of the methods a Julia `JustRelaxB200Ext` adds for
`heatdiffusion_PT!`, `thermal_bcs!`, `PTThermalCoeffs`, `update_thermal_coeffs!` (src/ext/CUDA/3D.jl:55-175, 220-226,
383-385).  Signatures follow the reference:

    heatdiffusion_PT_(thermal, pt_thermal, thermal_bc, K, ρCp, dt, grid|di; kwargs=dict(igg, iterMax, nout, verbose))
    heatdiffusion_PT_(thermal, pt_thermal, thermal_bc, rheology, args, dt, grid|di; kwargs=dict(igg, phase, stokes, ...))
        (src/thermal_diffusion/DiffusionPT_solver.jl:34-48, 181-197; one keyword literally named `kwargs`, quirk Q1)
    → SimpleNamespace(iter_count, norm_ResT)      (:148, :304)

Only marshalling happens here; the numerics run in libjrb200 (csrc/thermal.cu).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Optional

import numpy as np

from . import _abi
from .rheology import MaterialParams, lower_thermal
from .stokes import context
from .types import (require_uniform, B200Backend, CPUBackendTrait, Geometry, IGG, PTArray, TemperatureBoundaryConditions, ThermalArrays, backend,
                    data_ptr, is_device_array, legacy_uniform_grid, zeros)

FACES = ("left", "right", "front", "back", "top", "bot")


class PTThermalCoeffs:
    """PTThermalCoeffs — src/types/heat_diffusion.jl:30-44; constructors src/thermal_diffusion/DiffusionPT_coefficients.jl:
      PTThermalCoeffs(backend, K, ρCp, dt, di, li; ϵ, CFL)                          :11-26
      PTThermalCoeffs(backend, rheology, phase_ratios, args, dt, ni, di, li; ϵ, CFL)  :37-65
      PTThermalCoeffs(backend, rheology, args, dt, ni, di, li; ϵ, CFL)               :75-103
    θr_dτ, dτ_ρ are B200 arrays."""

    def __init__(self, backend_t, *a, ϵ: float = 1.0e-8, CFL: Optional[float] = None):
        CFL = 0.9 / math.sqrt(3) if CFL is None else CFL
        if backend_t is not B200Backend:
            raise ValueError("PTThermalCoeffs: this package only provides the B200 backend")
        self.CFL, self.ϵ = float(CFL), float(ϵ)
        if len(a) == 5 and not isinstance(a[0], (MaterialParams, tuple, list)):
            K, ρCp, dt, di, li = a
            self.Vpdτ = min(di) * CFL
            self.max_lxyz = max(li)
            self.max_lxyz2 = self.max_lxyz ** 2
            # Re = π + √(π² + ρCp·L²/K/dt); θr_dτ = L/Vpdτ/Re; dτ_ρ = Vpdτ·L/K/Re   (broadcasts on the device, :20-23)
            Re = math.pi + (math.pi * math.pi + ρCp * self.max_lxyz2 / K / dt).sqrt()
            self.θr_dτ = self.max_lxyz / self.Vpdτ / Re
            self.dτ_ρ = self.Vpdτ * self.max_lxyz / K / Re
            return
        if len(a) == 7:
            rheology, phase_ratios, args, dt, ni, di, li = a
        elif len(a) == 6:
            rheology, args, dt, ni, di, li = a
            phase_ratios = None
        else:
            raise TypeError("PTThermalCoeffs: unsupported argument list")
        self.Vpdτ = min(di) * CFL
        self.max_lxyz = max(li)
        self.max_lxyz2 = self.max_lxyz ** 2
        self.θr_dτ, self.dτ_ρ = zeros(B200Backend, *ni), zeros(B200Backend, *ni)
        update_thermal_coeffs_(self, rheology, phase_ratios, args, dt)


def _bc_opts(o, bc: Optional[TemperatureBoundaryConditions]):
    if bc is None:
        return
    for q, f in enumerate(FACES):
        o.no_flux[q] = int(bool(bc.no_flux.get(f, False)))
        o.periodic[q] = int(bool(bc.periodic.get(f, False)))
        cv = bc.constant_value.get(f, False)
        o.cv_active[q] = int(cv is not False)
        o.cv_value[q] = float(cv) if cv is not False else 0.0          # `true` acts as the number 1 (quirk Q16)
        cf = bc.constant_flux.get(f, False)
        o.cf_active[q] = int(not isinstance(cf, bool))                 # !isa(bc_flux.left, Bool)
        o.cf_value[q] = float(cf) if not isinstance(cf, bool) else 0.0


def _opts(pt, _di, dt, bc, *, form, rows=(), iterMax=50e3, nout=1e3):
    o = _abi.ThermalOpts()
    for q in range(3):
        o._di[q] = float(_di[q]) if q < len(_di) else 0.0
    o.dt, o.eps, o.iterMax, o.nout = float(dt), float(pt.ϵ), int(iterMax), int(nout)
    o.max_lxyz, o.Vpdtau, o.form = float(pt.max_lxyz), float(pt.Vpdτ), int(form)
    arr = (_abi.ThermalPhase * max(len(rows), 1))()
    for i, r in enumerate(rows):
        for k, v in r.items():
            setattr(arr[i], k, v)
    o.nphase, o.phases = len(rows), arr
    o._keep = arr
    _bc_opts(o, bc)
    if bc is not None and bc.dirichlet is not None and isinstance(bc.dirichlet[0], (int, float)):
        o.dir_const = float(bc.dirichlet[0])
    return o


def _fields(thermal: ThermalArrays, pt, **extra):
    fs = _abi.ThermalFields()
    ni = thermal.ni
    fs.ndim = len(ni)
    for q in range(3):
        fs.n[q] = ni[q] if q < len(ni) else 1
    src = dict(T=thermal.T, Told=thermal.Told, dT=thermal.ΔT, qTx=thermal.qTx, qTy=thermal.qTy, qTz=thermal.qTz, qTx2=thermal.qTx2,
               qTy2=thermal.qTy2, qTz2=thermal.qTz2, H=thermal.H, shear_heating=thermal.shear_heating, adiabatic=thermal.adiabatic,
               ResT=thermal.ResT, theta_r_dtau=pt.θr_dτ, dtau_rho=pt.dτ_ρ)
    src.update(extra)
    for nm, a in src.items():
        if a is not None:
            if not is_device_array(a):
                raise ValueError(f"heatdiffusion_PT_: `{nm}` must be a B200 array (use PTArray(B200Backend)(x)); no CPU fallback")
            setattr(fs, nm, data_ptr(a))
    fs._keep = src
    return fs


def _dirichlet_extra(bc):
    """(constant | value array, mask) → device arrays on the ghosted grid (src/boundaryconditions/Dirichlet.jl:127-136)"""
    if bc is None or bc.dirichlet is None:
        return {}
    const, mask = bc.dirichlet
    if mask is None:
        return {}
    out = dict(dir_mask=mask if is_device_array(mask) else PTArray(B200Backend)(mask))
    if const is None:  # DirichletBoundaryCondition(A): value = A, mask = A != 0
        A = out["dir_mask"]
        out["dir_value"] = A
        out["dir_mask"] = (A != 0).to(A.dtype)
    return out


def _phase_extra(phase):
    """PhaseRatios-like object with .center/.Vx/.Vy/.Vz given as B200 arrays of shape (nodes..., nphase) laid out
    [phase][node] (see phases.PhaseRatios), or None."""
    if phase is None:
        return {}
    return dict(phase_c=phase.center, phase_x=phase.Vx, phase_y=phase.Vy, phase_z=getattr(phase, "Vz", None))


def _grid_of(thermal, grid_or_di, igg):
    if isinstance(grid_or_di, Geometry):
        require_uniform(grid_or_di, "heatdiffusion_PT!")
        return grid_or_di
    di = grid_or_di.center if hasattr(grid_or_di, "center") else grid_or_di
    if any(np.ndim(x) > 0 for x in di):
        raise NotImplementedError("heatdiffusion_PT!: vector grid spacings (non-uniform grids) are outside the B200 backend's subset")
    return legacy_uniform_grid(thermal.ni, tuple(di), igg)


def heatdiffusion_PT_(thermal: ThermalArrays, pt_thermal, thermal_bc, a4, a5, dt, grid, *, kwargs=None, _niter=None):
    """heatdiffusion_PT!(thermal, pt_thermal, thermal_bc, K, ρCp, dt, grid; kwargs) or
    heatdiffusion_PT!(thermal, pt_thermal, thermal_bc, rheology, args, dt, grid; kwargs)."""
    kw = dict(igg=None, phase=None, stokes=None, b_width=(4, 4, 4), iterMax=50e3, nout=1e3, verbose=True)
    kw.update(kwargs or {})
    if isinstance(backend(thermal), CPUBackendTrait):
        raise RuntimeError("heatdiffusion_PT_: ThermalArrays live on the host (CPUBackend); this package only provides the B200 "
                           "backend. No CPU fallback.")
    if not isinstance(thermal_bc, TemperatureBoundaryConditions):
        raise TypeError(f"Unknown boundary conditions type: {type(thermal_bc)}")
    grid = _grid_of(thermal, grid, kw["igg"])
    stokes_P = stokes_P0 = None
    if isinstance(a4, (MaterialParams, tuple, list)):
        rheology, args = a4, a5
        rows = lower_thermal(rheology)
        if kw["phase"] is None and len(rows) != 1:
            raise ValueError("a rheology tuple needs `phase` (phase ratios)")
        P = args.get("P") if isinstance(args, dict) else getattr(args, "P", None)
        fs = _fields(thermal, pt_thermal, P=P, **_dirichlet_extra(thermal_bc), **_phase_extra(kw["phase"]))
        o = _opts(pt_thermal, grid._di.center, dt, thermal_bc, form=1, rows=rows, iterMax=kw["iterMax"], nout=kw["nout"])
        if kw["stokes"] is not None:
            stokes_P, stokes_P0 = data_ptr(kw["stokes"].P), data_ptr(kw["stokes"].P0)
    else:
        K, ρCp = a4, a5
        fs = _fields(thermal, pt_thermal, K=K, rhoCp=ρCp, **_dirichlet_extra(thermal_bc))
        o = _opts(pt_thermal, grid._di.center, dt, thermal_bc, form=0, iterMax=kw["iterMax"], nout=kw["nout"])
    res = _abi.ThermalResult()
    if _niter is not None:
        _abi.check(_abi.lib().jr_thermal_iterate(context(), C.byref(fs), C.byref(o), int(_niter), C.byref(res)))
        return SimpleNamespace(iter=int(res.iter), err=float(res.err), time=float(res.time_s), kernel_launches=int(res.kernel_launches))
    cap = int(o.iterMax // max(o.nout, 1)) + 2
    nr, ic = np.zeros(cap), np.zeros(cap, dtype=np.int64)
    res.cap, res.norm_ResT, res.iter_count = cap, nr.ctypes.data_as(C.POINTER(C.c_double)), ic.ctypes.data_as(C.POINTER(C.c_int64))
    _abi.check(_abi.lib().jr_heatdiffusion_PT(context(), C.byref(fs), C.byref(o), stokes_P, stokes_P0, C.byref(res)))
    n = int(res.nhist)
    if kw["verbose"] and (kw["igg"] is None or kw["igg"].me == 0):
        for q in range(n):
            print("iter = %d, err = %1.3e " % (ic[q], nr[q]))
    return SimpleNamespace(iter_count=ic[:n].copy(), norm_ResT=nr[:n].copy(), iter=int(res.iter), time=float(res.time_s),
                           kernel_launches=int(res.kernel_launches))


def thermal_iterate_(thermal, pt_thermal, thermal_bc, a4, a5, dt, grid, niter: int, *, kwargs=None):
    """exactly `niter` PT iterations, no convergence test, Told untouched (fixed-iteration parity / benchmark)"""
    return heatdiffusion_PT_(thermal, pt_thermal, thermal_bc, a4, a5, dt, grid, kwargs=kwargs, _niter=niter)


def thermal_bcs_(thermal, bcs: TemperatureBoundaryConditions):
    """thermal_bcs!(thermal, bcs) — src/ext/CUDA/3D.jl:220-226 → BoundaryConditions.jl:39-54.  `thermal` may be a
    ThermalArrays or the ghosted temperature array itself."""
    T = thermal.T if isinstance(thermal, ThermalArrays) else thermal
    if not is_device_array(T):
        raise RuntimeError("thermal_bcs_: host arrays; this package only provides the B200 backend")
    ni = [s - 2 for s in T.shape]
    o = _abi.ThermalOpts()
    _bc_opts(o, bcs)
    _abi.check(_abi.lib().jr_thermal_bcs(context(), data_ptr(T), len(ni), _abi.i32x(ni + [1] * (3 - len(ni))), C.byref(o)))


def update_thermal_coeffs_(pt_thermal, rheology, phase_ratios, args, dt):
    """update_thermal_coeffs!(pt_thermal, rheology, [phase_ratios,] args, dt) — DiffusionPT_coefficients.jl:164-208."""
    rows = lower_thermal(rheology)
    T = args["T"] if isinstance(args, dict) else args.T
    P = args.get("P") if isinstance(args, dict) else getattr(args, "P", None)
    ni = tuple(pt_thermal.θr_dτ.shape)
    fs = _abi.ThermalFields()
    fs.ndim = len(ni)
    for q in range(3):
        fs.n[q] = ni[q] if q < len(ni) else 1
    # only T, P, the phase ratios and the two outputs are read by the kernel; the remaining required slots alias T
    keep = dict(T=T, P=P, theta_r_dtau=pt_thermal.θr_dτ, dtau_rho=pt_thermal.dτ_ρ, **_phase_extra(phase_ratios))
    for nm in ("Told", "dT", "qTx", "qTy", "qTz", "qTx2", "qTy2", "qTz2", "H", "shear_heating", "adiabatic", "ResT"):
        keep.setdefault(nm, T)
    for nm, a in keep.items():
        if a is not None:
            setattr(fs, nm, data_ptr(a))
    o = _opts(pt_thermal, (1.0,) * len(ni), dt, None, form=1, rows=rows)
    _abi.check(_abi.lib().jr_thermal_pt_arrays(context(), C.byref(fs), C.byref(o)))
