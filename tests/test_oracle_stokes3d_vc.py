"""CPU oracle of the 3D multiphase visco-elasto-plastic Stokes PT loop (variant 3D-VC, src/stokes/Stokes3D.jl:447-668), no GPU.

The reference pins no numbers for this variant (test/test_shearband3D_MPI.jl:186-246 only runs it), so the restatement is
checked three ways:
 - the 3D shear-band setup of that test converges and yields (plastic multiplier active, τII on the Drucker-Prager envelope);
 - for a single non-plastic phase with uniform viscosity the converged 3D-VC solution equals the converged 3D-VA solution of
   the SolVi-pinned oracle driven by the same buoyancy field (both discretisations coincide there);
 - one iteration on a random multiphase state against an independent numpy / pure-Python restatement of the kernels.
"""
import ctypes as C
import math

import numpy as np

from justrelax_jl_b200 import rheology as R, setups
from justrelax_jl_b200.types import Geometry, PTStokesCoeffs, VelocityBoundaryConditions
from util import bc_flags

FS = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)


def test_shearband3d_converges_and_yields(oracle):
    s = setups.shearband3d(12)
    d = oracle.alloc_stokes(s.ni, s.fields)
    rows = R.lower_stokes(s.rheology)
    vc = oracle.vc_inputs(rows, R.gravity_of(s.rheology), s.ratios, g_scalar=True)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_viscosity3d(C.byref(fs), C.byref(opts), C.byref(vc), C.c_double(1.0))   # compute_viscosity!  test_shearband3D_MPI.jl:131
    oracle.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)
    txx_max = []
    for _ in range(8):
        out = oracle.solve3d_VC(d, s.ni, opts, vc)
        assert out["status"] == 0 and out["iter"] < s.kwargs["iterMax"]
        assert out["err_evo1"][-1] < 1.0e-5 * max(1.0, out["err_evo1"][0]) or out["err_evo1"][-1] < s.pt_stokes.ϵ_abs
        txx_max.append(d["txx"].max())
    # elastic build-up towards the viscous limit 2 η ε = 2, capped by the yield stress C cosϕ + P sinϕ
    assert txx_max[0] < txx_max[1] < txx_max[2]
    assert d["lam"].max() > 0 and d["EII_pl"].max() > 0
    F = d["tII"] - (1.6 + 0.5 * d["P"])               # F = τII − C cosϕ − P sinϕ  (≤ η_vp-regularised overshoot where yielding)
    assert F.max() < 0.05 and F[d["lam"] > 0].min() > -0.05
    for c in ("xx", "yy", "zz", "yz", "xz", "xy"):
        assert np.array_equal(d["t" + c], d["t" + c + "_o"])


def test_vc_matches_va_for_single_viscoelastic_phase(oracle):
    n = 10
    ni, li = (n, n, n), (1.0, 1.0, 1.0)
    grid = Geometry(ni, li)
    pt = PTStokesCoeffs(li, grid.di.center, ϵ_rel=1e-12, ϵ_abs=1e-11, CFL=0.9 / math.sqrt(3.1))
    xc, yc, zc = grid.xci
    T = np.zeros((n + 2,) * 3, order="F")
    T[1:-1, 1:-1, 1:-1] = np.sin(math.pi * xc)[:, None, None] * np.cos(math.pi * yc)[None, :, None] * np.sin(math.pi * zc)[None, None, :]
    el = R.ConstantElasticity(G=2.0, Kb=3.0)
    rheo = (R.SetMaterialParams(Phase=1, Density=R.T_Density(ρ0=1.0, α=0.5, T0=0.0), Gravity=R.ConstantGravity(g=1.0),
                                CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0), el)), Elasticity=el),)
    ratios = {nm: np.ones(sh + (1,), order="F") for nm, sh in dict(center=ni, xy=(n + 1, n + 1, n), yz=(n, n + 1, n + 1), xz=(n + 1, n, n + 1)).items()}
    dt = 0.5
    rng = np.random.default_rng(3)
    old = {f"t{c}_o": np.asfortranarray(rng.uniform(-0.1, 0.1, size=ni)) for c in ("xx", "yy", "zz")}
    # 3D-VC
    d = oracle.alloc_stokes(ni, dict(T=T, **old))
    vc = oracle.vc_inputs(R.lower_stokes(rheo), R.gravity_of(rheo), ratios, g_scalar=True)
    opts = oracle.make_opts(pt, grid._di.center, dt, FS, ni, iterMax=40000, nout=200, viscosity_relaxation=1.0)
    out = oracle.solve3d_VC(d, ni, opts, vc)
    assert out["status"] == 0 and out["iter"] < 40000
    # 3D-VA with the same body force and moduli as arrays
    rho = 1.0 * (1.0 - 0.5 * (T[1:-1, 1:-1, 1:-1] - 0.0))
    e = oracle.alloc_stokes(ni, dict(rhogz=np.asfortranarray(rho * 1.0), G=np.full(ni, 2.0, order="F"), K=np.full(ni, 3.0, order="F"), **old))
    assert np.array_equal(d["rhogz"], e["rhogz"])
    out2 = oracle.solve3d_VA(e, ni, opts)
    assert out2["status"] == 0 and out2["iter"] < 40000
    for nm in ("Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy"):
        scale = np.abs(e[nm]).max()
        assert np.abs(d[nm] - e[nm]).max() <= 2e-7 * scale, nm


def _harm4(a, b, c, d_):
    return 4 / (1 / a + 1 / b + 1 / c + 1 / d_)


def test_vc_iteration_matches_independent_restatement(oracle):
    """one 3D-VC iteration on a random three-phase state vs numpy (∇V, θ, ε, η, non-plastic edges) and a pure-Python evaluation of the
    Drucker-Prager return mapping at sampled centres"""
    s = setups.random_vc3d((7, 6, 5))
    ni = s.ni
    nx, ny, nz = ni
    d = oracle.alloc_stokes(ni, s.fields)
    ref = {k: v.copy(order="F") for k, v in d.items()}
    rows = R.lower_stokes(s.rheology)
    vc = oracle.vc_inputs(rows, R.gravity_of(s.rheology), s.ratios, g_scalar=True)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, FS, ni, iterMax=1, nout=1, viscosity_cutoff=s.kwargs["viscosity_cutoff"])
    oracle.iterate3d_VC(d, ni, opts, vc, 1)
    pt, dt = s.pt_stokes, s.dt
    _dx, _dy, _dz = s.grid._di.center
    rc = s.ratios["center"]
    # pre-loop: compute_viscosity! (ν = 1) → η0 ; loop: ητ = maxloc(η0), η1 = relax(η0)
    eta_p = np.array([r["eta"] for r in rows])
    single = rc > 0.999
    eta_mix = 1.0 / (np.where(rc != 0, rc / eta_p, 0.0)).sum(-1)
    first = np.argmax(single, -1)
    eta_new = np.clip(np.where(single.any(-1), eta_p[first], eta_mix), *s.kwargs["viscosity_cutoff"])
    eta0 = eta_new
    pe = np.pad(eta0, 1, mode="edge")
    ett = np.max([pe[a:a + nx, b:b + ny, c:c + nz] for a in range(3) for b in range(3) for c in range(3)], axis=0)
    assert np.array_equal(d["etatau"], ett)
    eta1 = np.clip((1 - 1e-2) * eta0 + 1e-2 * eta_new, *s.kwargs["viscosity_cutoff"])
    assert np.allclose(d["eta"], eta1, rtol=1e-15, atol=0)
    Vx, Vy, Vz = ref["Vx"], ref["Vy"], ref["Vz"]
    dVx = (Vx[1:, 1:-1, 1:-1] - Vx[:-1, 1:-1, 1:-1]) * _dx
    dVy = (Vy[1:-1, 1:, 1:-1] - Vy[1:-1, :-1, 1:-1]) * _dy
    dVz = (Vz[1:-1, 1:-1, 1:] - Vz[1:-1, 1:-1, :-1]) * _dz
    divV = dVx + dVy + dVz
    assert np.allclose(d["divV"], divV, rtol=0, atol=1e-13)
    Gp, Kp = np.array([r["G"] for r in rows]), np.array([r["Kb"] for r in rows])
    with np.errstate(invalid="ignore"):
        G = np.where(rc != 0, rc * Gp, 0.0).sum(-1)
        K = np.where(rc != 0, rc * Kp, 0.0).sum(-1)
    P0 = ref["P"]                                         # @copy P0 P
    psi = 1.0 / (1.0 / ett + 1.0 / (G * dt)) * pt.r / pt.θ_dτ
    theta = ((P0 / (K * dt) - divV + ref["Q"] / dt) * psi + P0) / (1 + psi / (K * dt))
    RP = -(P0 - P0) / (K * dt) - divV + ref["Q"] / dt
    assert np.allclose(d["RP"], RP, rtol=0, atol=1e-12)
    exx = dVx - divV / 3
    assert np.allclose(d["exx"], exx, rtol=0, atol=1e-13)
    # strain rate only over ni (quirk Q20): far edge planes keep their previous content
    eyz = ref["eyz"].copy()
    eyz[:, :ny, :nz] = 0.5 * (_dz * (Vy[1:-1, :-1, 1:-1] - Vy[1:-1, :-1, :-2]) + _dy * (Vz[1:-1, 1:-1, :-1] - Vz[1:-1, :-2, :-1]))[:, :ny, :nz]
    assert np.allclose(d["eyz"], eyz, rtol=0, atol=1e-13)
    assert np.array_equal(d["eyz"][:, ny, :], ref["eyz"][:, ny, :]) and np.array_equal(d["eyz"][:, :, nz], ref["eyz"][:, :, nz])
    # ---- centre return mapping, pure Python at sampled cells
    rng = np.random.default_rng(0)
    n_pl = 0
    for _ in range(60):
        i, j, k = (int(rng.integers(0, n)) for n in ni)
        r = rc[i, j, k]
        Gc = sum(r[p] * rows[p]["G"] for p in range(3) if r[p] != 0)
        Kc = sum(r[p] * rows[p]["Kb"] for p in range(3) if r[p] != 0)
        is_pl = any(r[p] != 0 and rows[p]["has_pl"] for p in range(3))
        eta_reg = sum(r[p] * rows[p]["eta_vp"] for p in range(3) if r[p] != 0 and rows[p]["has_pl"])
        et = d["eta"][i, j, k]
        _Gdt = 1.0 / (Gc * dt)
        dtr = 1.0 / (pt.θ_dτ + et * _Gdt + 1.0)
        eij = [d["exx"][i, j, k], d["eyy"][i, j, k], d["ezz"][i, j, k],
               0.25 * d["eyz"][i, j:j + 2, k:k + 2].sum(), 0.25 * d["exz"][i:i + 2, j, k:k + 2].sum(), 0.25 * d["exy"][i:i + 2, j:j + 2, k].sum()]
        tij = [ref[nm][i, j, k] for nm in ("txx", "tyy", "tzz", "tyz_c", "txz_c", "txy_c")]
        tijo = [ref[nm][i, j, k] for nm in ("txx_o", "tyy_o", "tzz_o", "tyz_o_c", "txz_o_c", "txy_o_c")]
        dtau = [(-(t - to) * et * _Gdt - t + 2 * et * e) * dtr for t, to, e in zip(tij, tijo, eij)]
        tr = [t + q for t, q in zip(tij, dtau)]
        II = lambda a: math.sqrt(0.5 * (a[0] ** 2 + a[1] ** 2 + a[2] ** 2) + a[3] ** 2 + a[4] ** 2 + a[5] ** 2)
        tII = II(tr)
        Pr = theta[i, j, k]
        Fy = sum(r[p] * ((tII - rows[p]["C"] * rows[p]["cosphi"] - Pr * rows[p]["sinphi"]) if rows[p]["has_pl"] else tII) for p in range(3) if r[p] != 0)
        wpl = sum(r[p] for p in range(3) if r[p] != 0 and rows[p]["has_pl"])
        dQdP = -sum(r[p] * rows[p]["sinpsi"] for p in range(3) if r[p] != 0 and rows[p]["has_pl"])
        dFdP = -sum(r[p] * rows[p]["sinphi"] for p in range(3) if r[p] != 0 and rows[p]["has_pl"])
        vol = 0.0 if math.isinf(Kc) else Kc * dt * dFdP * dQdP
        if is_pl and tII != 0 and Fy > 0:
            n_pl += 1
            lam = 0.2 * Fy / (et * dtr + eta_reg + vol)
            new = [t - 2 * et * (lam * wpl * 0.5 * t / tII) * dtr for t in tr]
            Pc = Pr - (0.0 if math.isinf(Kc) else Kc * dt * lam * dQdP)
            assert abs(d["lam"][i, j, k] - lam) <= 1e-12 * abs(lam)
        else:
            new, Pc = tr, Pr
            assert d["lam"][i, j, k] == 0.0
        for nm, v in zip(("txx", "tyy", "tzz", "tyz_c", "txz_c", "txy_c"), new):
            assert abs(d[nm][i, j, k] - v) <= 1e-12 * max(1.0, abs(v)), nm
        assert abs(d["P"][i, j, k] - Pc) <= 1e-12 * max(1.0, abs(Pc))
        assert abs(d["tII"][i, j, k] - II(new)) <= 1e-12
    assert 5 < n_pl < 60
    # ---- yz edges where no phase is plastic: τyz += dτ with harmonic η and clamped 4-cell averages
    ryz = s.ratios["yz"]
    pc = lambda A: np.pad(A, ((0, 0), (1, 1), (1, 1)), mode="edge")
    a, b, c_, e_ = pc(d["eta"])[:, :-1, :-1], pc(d["eta"])[:, 1:, :-1], pc(d["eta"])[:, :-1, 1:], pc(d["eta"])[:, 1:, 1:]
    etav = _harm4(a, b, c_, e_)
    with np.errstate(invalid="ignore"):
        Gv = np.where(ryz != 0, ryz * Gp, 0.0).sum(-1)
    _Gv = 1.0 / (Gv * dt)
    dtr = 1.0 / (pt.θ_dτ + etav * _Gv + 1.0)
    dtyz = dtr * (2 * etav * d["eyz"] - (ref["tyz"] - ref["tyz_o"]) * etav * _Gv - ref["tyz"])
    nonpl = (ryz[..., 0] == 0)
    assert nonpl.sum() > 20
    assert np.allclose(d["tyz"][nonpl], (ref["tyz"] + dtyz)[nonpl], rtol=0, atol=1e-13)
    assert np.all(d["pyz"][nonpl] == 0)
    assert np.abs(d["pyz"]).max() > 0                      # some edges yielded


def test_thermal_stress_pressure_form_identity(oracle):
    """compute_P_kernel! with ΔT (PressureKernels.jl:128-149,197-206): RP gains exactly α·ΔT/dt with α = Σ ratio·α_phase (the α of the
    density law, 0 for ConstantDensity: test/test_rheology.jl:57-116), and with ΔT ≡ 0 the result is the plain form's bit for bit"""
    import ctypes as C

    from justrelax_jl_b200 import rheology as R, setups

    ni = (9, 8, 7)
    s = setups.random_vc3d(ni, seed=3)
    flags = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)
    vc = oracle.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), s.ratios)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, flags, s.ni, iterMax=1, nout=1, viscosity_cutoff=s.kwargs["viscosity_cutoff"])
    dT = np.asfortranarray(np.random.default_rng(1).uniform(-30.0, 30.0, size=ni))
    out = {}
    for tag, arr in (("none", None), ("zero", np.zeros(ni, order="F")), ("dT", dT)):
        d = oracle.alloc_stokes(s.ni, s.fields)
        d["Pargs"] = d["P"]
        if arr is not None:
            d["dTargs"] = arr
        oracle.iterate3d_VC(d, s.ni, opts, vc, 1, finish=False)
        out[tag] = d
    assert np.array_equal(out["none"]["RP"], out["zero"]["RP"]) and np.array_equal(out["none"]["P"], out["zero"]["P"])
    alpha = s.ratios["center"][..., 0] * 3.0e-2     # only phase 1 carries a T-dependent density (α = 3e-2)
    assert np.allclose(out["dT"]["RP"] - out["none"]["RP"], alpha * dT / s.dt, rtol=1e-9, atol=1e-12)
    assert np.abs(out["dT"]["P"] - out["none"]["P"]).max() > 1e-3


def test_unknown_args_key_fails_loudly():
    """args keys the backend does not consume (melt_fraction: PressureKernels.jl:151-176; perturbation_C: StressUpdate.jl:146-176) raise"""
    import pytest

    from justrelax_jl_b200 import CPUBackend, StokesArrays
    from justrelax_jl_b200.stokes import vc_slots

    st = StokesArrays(CPUBackend, 4, 4, 4)
    for bad in ("melt_fraction", "ϕ", "anything_else"):
        with pytest.raises(NotImplementedError, match="refusing to ignore"):
            vc_slots(st, (st.P, st.P, st.P), {bad: st.P})
    assert "dTargs" not in vc_slots(st, (st.P, st.P, st.P), dict(dt=0.1, perturbation_C=st.P))   # accepted and unused, as in the reference


def in_plane_invariant(oracle, txx, tzz, txz, j=1):
    """tensor_invariant!(stokes.τ) of the 2D test on the (x, z) plane y = j of the 3D fields"""
    return oracle.tensor_invariant2d(np.asfortranarray(txx[:, j, :]), np.asfortranarray(tzz[:, j, :]), np.asfortranarray(txz[:, j, :]))


def test_extruded_shearband_pins_3d_vc_on_the_2d_reference_golden(oracle):
    """EXTERNAL PIN of the 3D-VC restatement: the reference's 2D shear-band test (test/test_shearband2D.jl, Drucker-Prager + elasticity +
    two phases) extruded along y and solved with the 3D multiphase solver must land on the 2D golden values
    (test/test_shearband2D.jl:197-201: extrema(τII) ≈ (1.5128689768248313, 1.6415759440014273) atol 1e-3, maximum(τxx) ≈ 1.6376258215356436
    atol 1e-4) — plane strain differs from the 2D kernels only by the out-of-plane deviatoric stress (|τyy| ≤ 0.02 here)."""
    s = setups.shearband3d_extruded(32, 4)
    d = oracle.alloc_stokes(s.ni, s.fields)
    vc = oracle.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), s.ratios)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_viscosity3d(C.byref(fs), C.byref(opts), C.byref(vc), C.c_double(1.0))
    oracle.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)
    txx_max = []
    for _ in range(s.nt):
        out = oracle.solve3d_VC(d, s.ni, opts, vc)
        assert out["status"] == 0 and out["err_evo1"][-1] < 1.0e-6
        txx_max.append(d["txx"].max())
    tII = in_plane_invariant(oracle, d["txx"], d["tzz"], d["txz"])
    assert abs(tII.min() - 1.5128689768248313) < 1.0e-3, tII.min()
    assert abs(tII.max() - 1.6415759440014273) < 1.0e-3, tII.max()
    assert abs(txx_max[-1] - 1.6376258215356436) < 1.0e-4, txx_max[-1]
    assert d["EII_pl"].max() > 0 and d["lam"].max() > 0
    assert np.abs(d["txx"] - d["txx"][:, :1, :]).max() == 0.0 and np.abs(d["tyy"]).max() < 0.03      # y-invariant; small out-of-plane stress
