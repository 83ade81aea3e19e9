"""ctypes wrapper of the CPU ORACLE (oracle/libjr_oracle.so) — test infrastructure, NOT product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjr_oracle.so")
_lib = None


def build(force: bool = False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_field_name.restype = C.c_char_p
        L.orc_mini3.restype = C.c_double
        L.orc_mini2.restype = C.c_double
        L.orc_sumsq_interior.restype = C.c_double
        L.orc_vc3_begin.restype = C.c_void_p
        _lib = L
    return _lib


def field_names():
    L = lib()
    return [L.orc_field_name(i).decode() for i in range(L.orc_field_count())]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def mini3(name, A, _d, i, j, k):
    A = np.asfortranarray(A, dtype=np.float64)
    return lib().orc_mini3(name.encode(), _dp(A), *map(C.c_int, A.shape), C.c_double(_d), C.c_int(i), C.c_int(j), C.c_int(k))


def mini2(name, A, _d, i, j):
    A = np.asfortranarray(A, dtype=np.float64)
    shp = A.shape if A.ndim == 2 else (A.shape[0], 1)
    return lib().orc_mini2(name.encode(), _dp(A), *map(C.c_int, shp), C.c_double(_d), C.c_int(i), C.c_int(j))


def make_fields(slots: dict, ni):
    """slots: name -> column-major float64 numpy array (kept alive by the caller)."""
    names = field_names()

    class Fields(C.Structure):
        _fields_ = [("ndim", C.c_int32), ("n", C.c_int32 * 3), ("f", C.c_void_p * len(names))]

    fs = Fields()
    fs.ndim = len(ni)
    for d in range(3):
        fs.n[d] = int(ni[d]) if d < len(ni) else 1
    for i, nm in enumerate(names):
        a = slots.get(nm)
        if a is None:
            fs.f[i] = None
        else:
            assert a.dtype == np.float64 and a.flags.f_contiguous, nm
            fs.f[i] = a.ctypes.data
    unknown = set(slots) - set(names)
    assert not unknown, unknown
    return fs


def alloc_stokes(ni, init: dict | None = None) -> dict:
    """Host StokesArrays (all slots the solvers touch), zero-initialised like constructors/stokes.jl,
    η = η_vep = ηv = 1, then overwritten by `init`."""
    nd = len(ni)
    z = lambda *s: np.zeros(s, order="F")
    d = {}
    if nd == 3:
        nx, ny, nz = ni
        c, v = (nx, ny, nz), (nx + 1, ny + 1, nz + 1)
        shapes = dict(Vx=(nx + 1, ny + 2, nz + 2), Vy=(nx + 2, ny + 1, nz + 2), Vz=(nx + 2, ny + 2, nz + 1),
                      xy=(nx + 1, ny + 1, nz), yz=(nx, ny + 1, nz + 1), xz=(nx + 1, ny, nz + 1),
                      Rx=(nx - 1, ny, nz), Ry=(nx, ny - 1, nz), Rz=(nx, ny, nz - 1))
    else:
        nx, ny = ni
        c, v = (nx, ny), (nx + 1, ny + 1)
        shapes = dict(Vx=(nx + 1, ny + 2), Vy=(nx + 2, ny + 1), xy=(nx + 1, ny + 1), Rx=(nx - 1, ny), Ry=(nx, ny - 1))
    for nm in ("P", "P0", "divV", "Q", "EII_pl", "EVol_pl", "e_vol_pl", "eta_vep", "etatau", "RP", "divU", "lam", "dPpsi",
               "eta", "rhogx", "rhogy", "K", "G"):
        d[nm] = z(*c)
    d["eta"][...] = 1.0
    d["eta_vep"][...] = 1.0
    d["etav"] = np.ones(v, order="F")
    d["lamv"] = z(*v)
    d["Vx"], d["Vy"], d["Ux"], d["Uy"] = z(*shapes["Vx"]), z(*shapes["Vy"]), z(*shapes["Vx"]), z(*shapes["Vy"])
    d["Rx"], d["Ry"] = z(*shapes["Rx"]), z(*shapes["Ry"])
    comps_c = ("xx", "yy", "zz") if nd == 3 else ("xx", "yy")
    shear = ("yz", "xz", "xy") if nd == 3 else ("xy",)
    for pre, suf in (("t", ""), ("t", "_o"), ("e", ""), ("p", ""), ("d", "")):
        for cc in comps_c:
            d[f"{pre}{cc}{suf}"] = z(*c)
        for sh in shear:
            d[f"{pre}{sh}{suf}"] = z(*shapes[sh])
            d[f"{pre}{sh}{suf}_c"] = z(*c)
        d[f"{pre}II{suf}"] = z(*c)
    if nd == 3:
        d["Vz"], d["Uz"], d["Rz"], d["rhogz"] = z(*shapes["Vz"]), z(*shapes["Vz"]), z(*shapes["Rz"]), z(*c)
        d["wyz"], d["wxz"], d["wxy"] = z(*shapes["yz"]), z(*shapes["xz"]), z(*shapes["xy"])
    else:
        d["wxy"] = z(*shapes["xy"])
        for nm in ("txx_v", "tyy_v", "txx_o_v", "tyy_o_v"):
            d[nm] = z(*v)
    d["T"] = z(*(n + 2 for n in ni))      # args.T (ghosted) and args.P of the VC variants
    d["Pargs"] = z(*c)
    if init and "dTargs" in init:         # args.ΔT: only present when the caller passes it (NULL slot = the plain compute_P! form)
        d["dTargs"] = z(*c)
    if init:
        for k, a in init.items():
            assert k in d, k
            assert d[k].shape == a.shape, (k, d[k].shape, a.shape)
            d[k][...] = a
    return d


class StokesOpts(C.Structure):
    _fields_ = [
        ("r", C.c_double), ("theta_dtau", C.c_double), ("eta_dtau", C.c_double),
        ("eps_rel", C.c_double), ("eps_abs", C.c_double),
        ("_di", C.c_double * 3), ("dt", C.c_double),
        ("iterMax", C.c_int64), ("nout", C.c_int64),
        ("n_g", C.c_int32 * 3),
        ("free_slip", C.c_int32 * 6), ("no_slip", C.c_int32 * 6), ("periodic", C.c_int32 * 6),
        ("viscosity_relaxation", C.c_double), ("lambda_relaxation", C.c_double),
        ("visc_cutoff_lo", C.c_double), ("visc_cutoff_hi", C.c_double),
        ("iterMin", C.c_int64), ("strain_rate_ni_only", C.c_int32), ("strain_increment", C.c_int32), ("displacement_bcs", C.c_int32),
        ("dT_ghosted", C.c_int32),
    ]


class StokesResult(C.Structure):
    _fields_ = [
        ("iter", C.c_int64), ("nhist", C.c_int64), ("err", C.c_double),
        ("err_evo1", C.POINTER(C.c_double)), ("err_evo2", C.POINTER(C.c_int64)),
        ("norm_Rx", C.POINTER(C.c_double)), ("norm_Ry", C.POINTER(C.c_double)),
        ("norm_Rz", C.POINTER(C.c_double)), ("norm_divV", C.POINTER(C.c_double)),
    ]


def make_opts(pt, _di, dt, flags: dict, n_g, *, iterMax, nout, viscosity_relaxation=1e-2, lambda_relaxation=0.2,
              viscosity_cutoff=(-np.inf, np.inf), iterMin=100, strain_rate_ni_only=0, strain_increment=0, displacement_bcs=0, dT_ghosted=0):
    """pt: object with r, θ_dτ, ηdτ, ϵ_rel, ϵ_abs; flags: dict(free_slip=[6], no_slip=[6], periodic=[6])."""
    o = StokesOpts()
    o.r, o.theta_dtau, o.eta_dtau, o.eps_rel, o.eps_abs = pt.r, pt.θ_dτ, pt.ηdτ, pt.ϵ_rel, pt.ϵ_abs
    for d in range(3):
        o._di[d] = float(_di[d]) if d < len(_di) else 0.0
        o.n_g[d] = int(n_g[d]) if d < len(n_g) else 1
    o.dt = float(dt)
    o.iterMax, o.nout = int(iterMax), int(nout)
    for nm in ("free_slip", "no_slip", "periodic"):
        arr = getattr(o, nm)
        for q in range(6):
            arr[q] = int(flags.get(nm, [0] * 6)[q])
    o.viscosity_relaxation, o.lambda_relaxation = viscosity_relaxation, lambda_relaxation
    o.visc_cutoff_lo, o.visc_cutoff_hi = viscosity_cutoff
    o.iterMin, o.strain_rate_ni_only = int(iterMin), int(strain_rate_ni_only)
    o.strain_increment, o.displacement_bcs, o.dT_ghosted = int(strain_increment), int(displacement_bcs), int(dT_ghosted)
    return o


class Hist:
    def __init__(self, cap):
        self.err_evo1 = np.zeros(cap)
        self.err_evo2 = np.zeros(cap, dtype=np.int64)
        self.norm_Rx, self.norm_Ry, self.norm_Rz, self.norm_divV = (np.zeros(cap) for _ in range(4))
        r = StokesResult()
        r.err_evo1, r.err_evo2 = _dp(self.err_evo1), self.err_evo2.ctypes.data_as(C.POINTER(C.c_int64))
        r.norm_Rx, r.norm_Ry, r.norm_Rz, r.norm_divV = _dp(self.norm_Rx), _dp(self.norm_Ry), _dp(self.norm_Rz), _dp(self.norm_divV)
        self.res = r

    def out(self):
        n = int(self.res.nhist)
        return dict(iter=int(self.res.iter), err_evo1=self.err_evo1[:n].copy(), err_evo2=self.err_evo2[:n].copy(),
                    norm_Rx=self.norm_Rx[:n].copy(), norm_Ry=self.norm_Ry[:n].copy(), norm_Rz=self.norm_Rz[:n].copy(),
                    norm_divV=self.norm_divV[:n].copy())


def solve3d_VA(slots, ni, opts):
    fs = make_fields(slots, ni)
    h = Hist(int(opts.iterMax // max(opts.nout, 1)) + 3)
    st = lib().orc_solve3d_VA(C.byref(fs), C.byref(opts), C.byref(h.res))
    out = h.out()
    out["status"] = st
    return out


def iterate3d_VA(slots, ni, opts, niter):
    fs = make_fields(slots, ni)
    return lib().orc_iterate3d_VA(C.byref(fs), C.byref(opts), C.c_int64(niter))


def sumsq(A, interior):
    A = np.asfortranarray(A)
    shp = list(A.shape) + [1] * (3 - A.ndim)
    return lib().orc_sumsq_interior(_dp(A), *map(C.c_int, shp), C.c_int(int(interior)))


# ------------------------------------------------------------------------------------------------------------
# thermal (oracle/thermal.c)
class ThermalFields(C.Structure):
    _names = ("T", "Told", "dT", "qTx", "qTy", "qTz", "qTx2", "qTy2", "qTz2", "H", "shear_heating", "adiabatic", "ResT",
              "theta_r_dtau", "dtau_rho", "K", "rhoCp", "P", "dir_mask", "dir_value", "phase_c", "phase_x", "phase_y", "phase_z")
    _fields_ = [("ndim", C.c_int32), ("n", C.c_int32 * 3)] + [(nm, C.c_void_p) for nm in _names]


class ThermalPhase(C.Structure):
    _fields_ = [("rho_kind", C.c_int32), ("has_Hr", C.c_int32), ("rho0", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
                ("T0", C.c_double), ("P0", C.c_double), ("Cp", C.c_double), ("k", C.c_double), ("Hr", C.c_double),
                ("k_kind", C.c_int32), ("_pad", C.c_int32), ("k_a", C.c_double), ("k_b", C.c_double), ("k_c", C.c_double), ("k_d", C.c_double)]


class ThermalOpts(C.Structure):
    _fields_ = [("_di", C.c_double * 3), ("dt", C.c_double), ("eps", C.c_double), ("iterMax", C.c_int64), ("nout", C.c_int64),
                ("max_lxyz", C.c_double), ("Vpdtau", C.c_double), ("form", C.c_int32), ("nphase", C.c_int32),
                ("phases", C.POINTER(ThermalPhase)), ("dir_const", C.c_double),
                ("no_flux", C.c_int32 * 6), ("cv_active", C.c_int32 * 6), ("cf_active", C.c_int32 * 6), ("periodic", C.c_int32 * 6),
                ("cv_value", C.c_double * 6), ("cf_value", C.c_double * 6)]


class ThermalResult(C.Structure):
    _fields_ = [("iter", C.c_int64), ("nhist", C.c_int64), ("cap", C.c_int64), ("err", C.c_double),
                ("norm_ResT", C.POINTER(C.c_double)), ("iter_count", C.POINTER(C.c_int64))]


def alloc_thermal(ni, init: dict | None = None) -> dict:
    """Host ThermalArrays + PTThermalCoeffs arrays (zero-initialised, constructors/heat_diffusion.jl:38-120)."""
    z = lambda *s: np.zeros(s, order="F")
    g = tuple(n + 2 for n in ni)
    d = dict(T=z(*g), Told=z(*g), dT=z(*g))
    for nm in ("H", "shear_heating", "adiabatic", "ResT", "theta_r_dtau", "dtau_rho"):
        d[nm] = z(*ni)
    for a, nm in enumerate(("qTx", "qTy", "qTz")[:len(ni)]):
        e = tuple(n + (1 if b == a else 0) for b, n in enumerate(ni))
        d[nm], d[nm + "2"] = z(*e), z(*e)
    if init:
        for k, a in init.items():
            d[k] = np.asfortranarray(a, dtype=np.float64).copy(order="F")
    return d


def thermal_fields(slots: dict, ni):
    fs = ThermalFields()
    fs.ndim = len(ni)
    for q in range(3):
        fs.n[q] = int(ni[q]) if q < len(ni) else 1
    for nm in ThermalFields._names:
        a = slots.get(nm)
        if a is not None:
            assert a.dtype == np.float64 and a.flags.f_contiguous, nm
            setattr(fs, nm, a.ctypes.data)
    return fs


FACES = ("left", "right", "front", "back", "top", "bot")


def thermal_opts(*, _di, dt, eps, iterMax, nout, max_lxyz, Vpdtau, form, phases=(), bc=None, dir_const=0.0):
    """bc: object with dicts no_flux / constant_value / constant_flux / periodic keyed by face name
    (False = inactive; True counts as the number 1 for constant_value — quirk Q16)."""
    o = ThermalOpts()
    for q in range(3):
        o._di[q] = float(_di[q]) if q < len(_di) else 0.0
    o.dt, o.eps, o.iterMax, o.nout = float(dt), float(eps), int(iterMax), int(nout)
    o.max_lxyz, o.Vpdtau, o.form, o.dir_const = float(max_lxyz), float(Vpdtau), int(form), float(dir_const)
    arr = (ThermalPhase * max(len(phases), 1))()
    for i, p in enumerate(phases):
        for k, v in p.items():
            setattr(arr[i], k, (C.c_double * len(v))(*v) if isinstance(v, (list, tuple)) else v)
    o.nphase, o.phases = len(phases), arr
    o._keep = arr
    for q, f in enumerate(FACES):
        if bc is None:
            continue
        o.no_flux[q] = int(bool(bc.no_flux.get(f, False)))
        o.periodic[q] = int(bool(bc.periodic.get(f, False)))
        cv = bc.constant_value.get(f, False)
        o.cv_active[q] = int(cv is not False)
        o.cv_value[q] = float(cv) if cv is not False else 0.0
        cf = bc.constant_flux.get(f, False)
        o.cf_active[q] = int(not isinstance(cf, bool))  # !isa(bc_flux.left, Bool)  DiffusionPT_kernels.jl:13
        o.cf_value[q] = float(cf) if not isinstance(cf, bool) else 0.0
    return o


def heatdiffusion_PT(slots, ni, opts, stokes_P=None, stokes_P0=None):
    fs = thermal_fields(slots, ni)
    cap = int(opts.iterMax // max(opts.nout, 1)) + 2
    nr, ic = np.zeros(cap), np.zeros(cap, dtype=np.int64)
    r = ThermalResult()
    r.cap, r.norm_ResT, r.iter_count = cap, _dp(nr), ic.ctypes.data_as(C.POINTER(C.c_int64))
    lib().orc_heatdiffusion_PT(C.byref(fs), C.byref(opts), None if stokes_P is None else _dp(stokes_P),
                               None if stokes_P0 is None else _dp(stokes_P0), C.byref(r))
    n = int(r.nhist)
    return dict(iter=int(r.iter), iter_count=ic[:n].copy(), norm_ResT=nr[:n].copy(), err=float(r.err))


# ------------------------------------------------------------------------------------------------------------
# 2D Stokes (oracle/stokes2d.c) and the multiphase (VC) inputs
class StokesPhase(C.Structure):
    _fields_ = [("eta", C.c_double), ("G", C.c_double), ("Kb", C.c_double), ("has_pl", C.c_int32), ("rho_kind", C.c_int32),
                ("C", C.c_double), ("sinphi", C.c_double), ("cosphi", C.c_double), ("sinpsi", C.c_double), ("eta_vp", C.c_double),
                ("rho0", C.c_double), ("alpha", C.c_double), ("beta", C.c_double), ("T0", C.c_double), ("P0", C.c_double),
                ("soft_C_kind", C.c_int32), ("_pad", C.c_int32), ("soft_C", C.c_double * 6)]


class VcInputs(C.Structure):
    _fields_ = [("nphase", C.c_int32), ("g_scalar", C.c_int32), ("phases", C.POINTER(StokesPhase)), ("g", C.c_double * 3),
                ("ph_center", C.c_void_p), ("ph_vertex", C.c_void_p), ("ph_xy", C.c_void_p), ("ph_yz", C.c_void_p), ("ph_xz", C.c_void_p),
                ("free_surface", C.c_double)]


def vc_inputs(rows, g, ratios: dict, *, g_scalar=True, free_surface=0.0, Phase=StokesPhase, Inputs=VcInputs):
    """rows: list of dicts with the StokesPhase fields; ratios: name -> column-major array [node..., phase]."""
    arr = (Phase * len(rows))()
    for i, r in enumerate(rows):
        for k, v in r.items():
            setattr(arr[i], k, (C.c_double * len(v))(*v) if isinstance(v, (list, tuple)) else v)
    vc = Inputs()
    vc.nphase, vc.g_scalar, vc.phases, vc.free_surface = len(rows), int(g_scalar), arr, float(free_surface)
    for q in range(3):
        vc.g[q] = float(g[q])
    keep = [arr]
    for nm in ("center", "vertex", "xy", "yz", "xz"):
        a = ratios.get(nm)
        if a is not None:
            if isinstance(a, np.ndarray):
                assert a.dtype == np.float64 and a.flags.f_contiguous, nm
                setattr(vc, "ph_" + nm, a.ctypes.data)
            else:
                setattr(vc, "ph_" + nm, a)
            keep.append(a)
    vc._keep = keep
    return vc


def solve2d_V2(slots, ni, opts):
    fs = make_fields(slots, ni)
    h = Hist(int(opts.iterMax // max(opts.nout, 1)) + 3)
    st = lib().orc_solve2d_V2(C.byref(fs), C.byref(opts), C.byref(h.res))
    out = h.out()
    out["status"] = st
    return out


def iterate2d_V2(slots, ni, opts, niter):
    fs = make_fields(slots, ni)
    return lib().orc_iterate2d_V2(C.byref(fs), C.byref(opts), C.c_int64(niter))


def solve2d_VC(slots, ni, opts, vc):
    fs = make_fields(slots, ni)
    h = Hist(int(opts.iterMax // max(opts.nout, 1)) + 3)
    st = lib().orc_solve2d_VC(C.byref(fs), C.byref(opts), C.byref(vc), C.byref(h.res))
    out = h.out()
    out["status"] = st
    return out


def iterate2d_VC(slots, ni, opts, vc, niter, finish=False):
    fs = make_fields(slots, ni)
    return lib().orc_iterate2d_VC(C.byref(fs), C.byref(opts), C.byref(vc), C.c_int64(niter), C.c_int(int(finish)))


def solve3d_VC(slots, ni, opts, vc):
    fs = make_fields(slots, ni)
    h = Hist(int(opts.iterMax // max(opts.nout, 1)) + 3)
    st = lib().orc_solve3d_VC(C.byref(fs), C.byref(opts), C.byref(vc), C.byref(h.res))
    out = h.out()
    out["status"] = st
    return out


def iterate3d_VC(slots, ni, opts, vc, niter, finish=False):
    fs = make_fields(slots, ni)
    return lib().orc_iterate3d_VC(C.byref(fs), C.byref(opts), C.byref(vc), C.c_int64(niter), C.c_int(int(finish)))


def tensor_invariant2d(xx, yy, xy):
    II = np.zeros(xx.shape, order="F")
    lib().orc_tensor_invariant2d(_dp(II), _dp(xx), _dp(yy), _dp(xy), C.c_int(xx.shape[0]), C.c_int(xx.shape[1]))
    return II


def phase_ratios_from_arrays(phase_arrays, xci, xvi):
    """update_phase_ratios_{2,3}D!(phase_ratios, phase_arrays, xci, xvi) → dict of (nodes..., nphase) column-major arrays"""
    ni = phase_arrays[0].shape
    nd, N = len(ni), len(phase_arrays)
    n3 = (C.c_int32 * 3)(*[int(ni[d]) if d < nd else 1 for d in range(3)])
    ph = [np.asfortranarray(a, dtype=np.float64) for a in phase_arrays]
    xc = [np.ascontiguousarray(x, dtype=np.float64) for x in xci]
    xv = [np.ascontiguousarray(x, dtype=np.float64) for x in xvi]
    pp = lambda arrs: (C.POINTER(C.c_double) * max(len(arrs), 3))(*[_dp(a) for a in arrs])
    shp = dict(center=ni, vertex=tuple(n + 1 for n in ni))
    for a, nm in enumerate(("Vx", "Vy", "Vz")[:nd]):
        shp[nm] = tuple(n + (1 if b == a else 0) for b, n in enumerate(ni))
    if nd == 3:
        nx, ny, nz = ni
        shp.update(xy=(nx + 1, ny + 1, nz), yz=(nx, ny + 1, nz + 1), xz=(nx + 1, ny, nz + 1))
    out = {k: np.zeros(v + (N,), order="F") for k, v in shp.items()}
    g = lambda k: _dp(out[k]) if k in out else None
    lib().orc_phase_ratios_from_arrays(C.c_int(nd), n3, C.c_int(N), pp(ph), pp(xc), pp(xv), g("center"), g("vertex"), g("Vx"), g("Vy"), g("Vz"),
                                       g("xy"), g("yz"), g("xz"))
    return out
