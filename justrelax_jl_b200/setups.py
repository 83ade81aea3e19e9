"""Synthetic model setups (the callers either side of the hot path) restated from the reference's
miniapps/tests so that tests and bench.py run THE configurations BASELINE.json names.

Everything here is host-side numpy (setup runs once; it is not the hot path).  Each function returns
host arrays keyed by ABI slot name plus the scalar parameters, so the same inputs can be handed to the
B200 backend and to the CPU oracle.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

from .types import Geometry, IGG, PTStokesCoeffs, VelocityBoundaryConditions


def _smooth3(A: np.ndarray, fact: float) -> np.ndarray:
    """smooth!  miniapps/benchmarks/stokes3D/solvi/SolVi3D.jl:9-12 (ParallelStencil @inn/@d2_*i)."""
    A2 = A.copy(order="F")
    c = A[1:-1, 1:-1, 1:-1]
    d2x = (A[2:, 1:-1, 1:-1] - c) - (c - A[:-2, 1:-1, 1:-1])
    d2y = (A[1:-1, 2:, 1:-1] - c) - (c - A[1:-1, :-2, 1:-1])
    d2z = (A[1:-1, 1:-1, 2:] - c) - (c - A[1:-1, 1:-1, :-2])
    A2[1:-1, 1:-1, 1:-1] = c + 1.0 / 6.1 / fact * (d2x + d2y + d2z)
    return A2


def solvi3d(nx=31, ny=31, nz=31, *, Δη=1.0e-3, lx=1.0e1, ly=1.0e1, lz=1.0e1, rc=1.0e0, εbg=1.0e0, igg: IGG | None = None,
            smooth_passes: int = 10, update_halo=None, divfree: bool = False, global_coords: bool = False):
    """3D SolVi inclusion benchmark (config 4): miniapps/benchmarks/stokes3D/solvi/SolVi3D.jl:45-130.

    Returns host fields (numpy, column-major) and parameters for variant 3D-VA:
    η with a low-viscosity sphere smoothed 10×, G = 1, Kb = Inf, dt = Inf, pure-shear velocity
    (pureshear_bc!, src/boundaryconditions/pure_shear.jl:15-32 incl. quirk Q18), free slip on all six faces,
    ρg = 0, PTStokesCoeffs(li, di; CFL = 1/√3), kwargs = (iterMax = 5000, nout = 100).

    Multi-rank (igg.nprocs > 1): like the miniapp, di = li / (nx_g, ny_g, nz_g), the inclusion test uses LOCAL indices
    (SolVi3D.jl:14-25), and `update_halo(η)` (in place, host array) is called after every smoothing pass (:38-42).
    """
    igg = igg or IGG()
    ni = (nx, ny, nz)
    li = (lx, ly, lz)
    grid = Geometry(ni, li, origin=(0.0, 0.0, 0.0), igg=igg)
    di = grid.di.center
    dx, dy, dz = di
    pt_stokes = PTStokesCoeffs(li, di, CFL=1 / math.sqrt(3))

    # viscosity  SolVi3D.jl:14-45  (local indices, exactly as the reference kernel)
    ix = np.arange(nx, dtype=np.float64)[:, None, None]
    iy = np.arange(ny, dtype=np.float64)[None, :, None]
    iz = np.arange(nz, dtype=np.float64)[None, None, :]
    rad = np.sqrt((ix * dx + 0.5 * dx - 0.5 * lx) ** 2 + (iy * dy + 0.5 * dy - 0.5 * ly) ** 2 + (iz * dz + 0.5 * dz - 0.5 * lz) ** 2)
    if global_coords:
        # NOT the miniapp: inclusion placed by GLOBAL cell-centre coordinates, so that the decomposed viscosity field is
        # one consistent global field (tests that compare a decomposed solve with the single-block solve)
        xc, yc, zc = grid.xci
        rad = np.sqrt((xc[:, None, None] - 0.5 * lx) ** 2 + (yc[None, :, None] - 0.5 * ly) ** 2 + (zc[None, None, :] - 0.5 * lz) ** 2)
    η = np.full(ni, 1.0, order="F")
    η[rad <= rc] = Δη
    for _ in range(smooth_passes):
        η = np.asfortranarray(_smooth3(η, 1.0))
        if update_halo is not None:
            update_halo(η)
    η = np.asfortranarray(η)

    xv, yv, zv = grid.xvi
    Vx = np.zeros((nx + 1, ny + 2, nz + 2), order="F")
    Vy = np.zeros((nx + 2, ny + 1, nz + 2), order="F")
    Vz = np.zeros((nx + 2, ny + 2, nz + 1), order="F")
    Vx[:, 1:-1, 1:-1] = (εbg * xv)[:, None, None]
    # Q18: the reference uses the x-vertex coordinates for Vy (only well-formed when nx == ny)
    Vy[1:-1, :, 1:-1] = (εbg * (xv if nx == ny else yv))[None, :, None]
    Vz[1:-1, 1:-1, :] = (-εbg * zv)[None, None, :]
    if divfree:
        # NOT the reference's pureshear_bc!: a divergence-free variant (Vy from the y vertices, Vz = −2 εbg z) for which
        # the continuity residual can converge, used by tests that need the loop to END on its tolerance
        Vy[1:-1, :, 1:-1] = (εbg * yv)[None, :, None]
        Vz[1:-1, 1:-1, :] = (-2.0 * εbg * zv)[None, None, :]

    flow_bcs = VelocityBoundaryConditions(
        free_slip=dict(left=True, right=True, top=True, bot=True, back=True, front=True),
        no_slip=dict(left=False, right=False, top=False, bot=False, back=False, front=False),
    )
    fields = dict(Vx=Vx, Vy=Vy, Vz=Vz, eta=η, G=np.full(ni, 1.0, order="F"), K=np.full(ni, np.inf, order="F"),
                  rhogx=np.zeros(ni, order="F"), rhogy=np.zeros(ni, order="F"), rhogz=np.zeros(ni, order="F"))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=igg, pt_stokes=pt_stokes, flow_bcs=flow_bcs, dt=math.inf,
                           fields=fields, kwargs=dict(iterMax=5000, nout=100, verbose=False))


def burstedde3d(n=16, *, β=10.0):
    """test/test_stokes_burstedde.jl:28-40 + miniapps/benchmarks/stokes3D/burstedde/Burstedde.jl (variant 3D-VA, dt = Inf, G = K = Inf):
    the manufactured solution of Burstedde et al. (2013) on the unit cube — η = exp(1 − β(x(1−x) + y(1−y) + z(1−z))), body force from the
    analytical velocity / pressure, the analytical velocity prescribed on every face and ghost layer (no face is free-slip or no-slip, so
    flow_bcs! leaves them untouched), net boundary flux removed, PTStokesCoeffs(li, di; CFL = 1/√3), kwargs = (iterMax = 100e3, nout = 1e3)."""
    ni, li = (n, n, n), (1.0, 1.0, 1.0)
    grid = Geometry(ni, li, origin=(0.0, 0.0, 0.0))
    di = grid.di.center
    (xc, yc, zc), (xv, yv, zv) = grid.xci, grid.xvi
    X, Y, Z = np.meshgrid(xc, yc, zc, indexing="ij")
    η = np.exp(1 - β * (X * (1 - X) + Y * (1 - Y) + Z * (1 - Z)))
    dηdx, dηdy, dηdz = -β * (1 - 2 * X) * η, -β * (1 - 2 * Y) * η, -β * (1 - 2 * Z) * η
    fx = ((Y * Z + 3 * X ** 2 * Y ** 3 * Z) - η * (2 + 6 * X * Y)) - dηdx * (2 + 4 * X + 2 * Y + 6 * X ** 2 * Y) \
        - dηdy * (X + X ** 3 + Y + 2 * X * Y ** 2) - dηdz * (-3 * Z - 10 * X * Y * Z)
    fy = ((X * Z + 3 * X ** 3 * Y ** 2 * Z) - η * (2 + 2 * X ** 2 + 2 * Y ** 2)) - dηdx * (X + X ** 3 + Y + 2 * X * Y ** 2) \
        - dηdy * (2 + 2 * X + 4 * Y + 4 * X ** 2 * Y) - dηdz * (-3 * Z - 5 * X ** 2 * Z)
    fz = ((X * Y + X ** 3 * Y ** 3) - η * (-10 * Y * Z)) - dηdx * (-3 * Z - 10 * X * Y * Z) - dηdy * (-3 * Z - 5 * X ** 2 * Z) \
        - dηdz * (-4 - 6 * X - 6 * Y - 10 * X ** 2 * Y)
    vx = lambda x, y: x + x ** 2 + x * y + x ** 3 * y
    vy = lambda x, y: y + x * y + y ** 2 + x ** 2 * y ** 2
    vz = lambda x, y, z: -2 * z - 3 * x * z - 3 * y * z - 5 * x ** 2 * y * z
    # ghosted centre coordinates: LinRange(xci[1] − d, xci[end] + d, n + 2)   Burstedde.jl:44-48
    gc = [np.linspace(c[0] - d, c[-1] + d, c.size + 2) for c, d in zip((xc, yc, zc), di)]

    def shell(shape, values):
        A = np.zeros(shape, order="F")
        m = np.zeros(shape, dtype=bool)
        for ax in range(3):
            idx = [slice(None)] * 3
            for side in (0, -1):
                idx[ax] = side
                m[tuple(idx)] = True
        A[m] = values[m]
        return A

    Vx = shell((n + 1, n + 2, n + 2), vx(xv[:, None, None], gc[1][None, :, None]) * np.ones((1, 1, n + 2)))
    Vy = shell((n + 2, n + 1, n + 2), vy(gc[0][:, None, None], yv[None, :, None]) * np.ones((1, 1, n + 2)))
    Vz = shell((n + 2, n + 2, n + 1), vz(gc[0][:, None, None], gc[1][None, :, None], zv[None, None, :]))
    # remove_net_flux!  Burstedde.jl:100-122
    dx, dy, dz = di
    Ax, Ay, Az = dy * dz, dx * dz, dx * dy
    flux = ((Vx[-1, 1:-1, 1:-1].sum() - Vx[0, 1:-1, 1:-1].sum()) * Ax + (Vy[1:-1, -1, 1:-1].sum() - Vy[1:-1, 0, 1:-1].sum()) * Ay
            + (Vz[1:-1, 1:-1, -1].sum() - Vz[1:-1, 1:-1, 0].sum()) * Az)
    δ = flux / (2 * (n * n * Ax + n * n * Ay + n * n * Az))
    Vx[0, :, :] += δ; Vx[-1, :, :] -= δ
    Vy[:, 0, :] += δ; Vy[:, -1, :] -= δ
    Vz[:, :, 0] += δ; Vz[:, :, -1] -= δ
    none = dict(left=False, right=False, top=False, bot=False, back=False, front=False)
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(none), no_slip=dict(none))
    F = np.asfortranarray
    fields = dict(Vx=Vx, Vy=Vy, Vz=Vz, eta=F(η), G=np.full(ni, np.inf, order="F"), K=np.full(ni, np.inf, order="F"),
                  rhogx=F(-fx), rhogy=F(-fy), rhogz=F(-fz))

    def error_norms(Vx_, Vy_, Vz_, P_):
        """vizBurstedde.jl error_norms: √(Σ e² ΔV) against the analytical fields, pressures with their mean removed"""
        dV = dx * dy * dz
        L2 = lambda e: math.sqrt(float(np.sum(e * e)) * dV)
        ax_ = vx(xv[:, None, None], yc[None, :, None]) * np.ones((1, 1, n))
        ay_ = vy(xc[:, None, None], yv[None, :, None]) * np.ones((1, 1, n))
        az_ = vz(xc[:, None, None], yc[None, :, None], zv[None, None, :])
        p = X * Y * Z + X ** 3 * Y ** 3 * Z - 5 / 32
        return (L2((P_ - P_.mean()) - (p - p.mean())), L2(Vx_[:, 1:-1, 1:-1] - ax_), L2(Vy_[1:-1, :, 1:-1] - ay_), L2(Vz_[1:-1, 1:-1, :] - az_))

    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=PTStokesCoeffs(li, di, CFL=1 / math.sqrt(3)), flow_bcs=flow_bcs,
                           dt=math.inf, fields=fields, error_norms=error_norms, kwargs=dict(iterMax=100.0e3, nout=1.0e3, verbose=False))


def taylor_green3d(n=16):
    """test/test_stokes_taylor_green.jl:29-40 + miniapps/benchmarks/stokes3D/taylor_green/TaylorGreen.jl (variant 3D-VA, η = 1, dt = Inf,
    G = K = Inf): the FVCA8 Taylor-Green Stokes solution on the unit cube — V = (−2 cos2πx sin2πy sin2πz, sin2πx cos2πy sin2πz,
    sin2πx sin2πy cos2πz), P = −6π sin2πx sin2πy sin2πz, ρg_x = 36π² cos2πx sin2πy sin2πz; the analytical velocity is prescribed on every
    face and ghost layer (no free-slip / no-slip face), PTStokesCoeffs(li, di; CFL = 1/√3), kwargs = (iterMax = 100e3, nout = 1e3)."""
    ni, li = (n, n, n), (1.0, 1.0, 1.0)
    grid = Geometry(ni, li, origin=(0.0, 0.0, 0.0))
    di = grid.di.center
    (xc, yc, zc), (xv, yv, zv) = grid.xci, grid.xvi
    tp = 2 * math.pi
    X, Y, Z = np.meshgrid(xc, yc, zc, indexing="ij")
    vx = lambda x, y, z: -2 * np.cos(tp * x) * np.sin(tp * y) * np.sin(tp * z)
    vy = lambda x, y, z: np.sin(tp * x) * np.cos(tp * y) * np.sin(tp * z)
    vz = lambda x, y, z: np.sin(tp * x) * np.sin(tp * y) * np.cos(tp * z)
    gc = [np.linspace(c[0] - d, c[-1] + d, c.size + 2) for c, d in zip((xc, yc, zc), di)]

    def shell(values):
        A = np.zeros(values.shape, order="F")
        m = np.zeros(values.shape, dtype=bool)
        for ax in range(3):
            idx = [slice(None)] * 3
            for side in (0, -1):
                idx[ax] = side
                m[tuple(idx)] = True
        A[m] = values[m]
        return A

    Vx = shell(vx(xv[:, None, None], gc[1][None, :, None], gc[2][None, None, :]))
    Vy = shell(vy(gc[0][:, None, None], yv[None, :, None], gc[2][None, None, :]))
    Vz = shell(vz(gc[0][:, None, None], gc[1][None, :, None], zv[None, None, :]))
    none = dict(left=False, right=False, top=False, bot=False, back=False, front=False)
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(none), no_slip=dict(none))
    F = np.asfortranarray
    fx = 36 * math.pi ** 2 * np.cos(tp * X) * np.sin(tp * Y) * np.sin(tp * Z)
    fields = dict(Vx=Vx, Vy=Vy, Vz=Vz, eta=np.ones(ni, order="F"), G=np.full(ni, np.inf, order="F"), K=np.full(ni, np.inf, order="F"),
                  rhogx=F(fx), rhogy=np.zeros(ni, order="F"), rhogz=np.zeros(ni, order="F"))
    dx, dy, dz = di

    def error_norms(Vx_, Vy_, Vz_, P_):
        dV = dx * dy * dz
        L2 = lambda e: math.sqrt(float(np.sum(e * e)) * dV)
        p = -6 * math.pi * np.sin(tp * X) * np.sin(tp * Y) * np.sin(tp * Z)
        return (L2((P_ - P_.mean()) - (p - p.mean())), L2(Vx_[:, 1:-1, 1:-1] - vx(xv[:, None, None], yc[None, :, None], zc[None, None, :])),
                L2(Vy_[1:-1, :, 1:-1] - vy(xc[:, None, None], yv[None, :, None], zc[None, None, :])),
                L2(Vz_[1:-1, 1:-1, :] - vz(xc[:, None, None], yc[None, :, None], zv[None, None, :])))

    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=PTStokesCoeffs(li, di, CFL=1 / math.sqrt(3)), flow_bcs=flow_bcs,
                           dt=math.inf, fields=fields, error_norms=error_norms, kwargs=dict(iterMax=100.0e3, nout=1.0e3, verbose=False))


def random_stokes3d(ni, seed=20261017, *, dt=0.7, finite_K=True, const_rhog=None):
    """Seeded random state for kernel-level parity fuzzing of variant 3D-VA:
    V, τ ~ U(−1,1), P ~ U(0,1), η ~ 10^U(−3,0), G ~ U(0.5,2), K ~ U(1,4) (or Inf), ρg ~ U(−1,1)."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = ni
    U = lambda *s: np.asfortranarray(rng.uniform(-1.0, 1.0, size=s))
    f = dict(
        Vx=U(nx + 1, ny + 2, nz + 2), Vy=U(nx + 2, ny + 1, nz + 2), Vz=U(nx + 2, ny + 2, nz + 1),
        P=np.asfortranarray(rng.uniform(0, 1, size=ni)), P0=np.asfortranarray(rng.uniform(0, 1, size=ni)),
        Q=np.asfortranarray(rng.uniform(-0.1, 0.1, size=ni)),
        txx=U(*ni), tyy=U(*ni), tzz=U(*ni), tyz=U(nx, ny + 1, nz + 1), txz=U(nx + 1, ny, nz + 1), txy=U(nx + 1, ny + 1, nz),
        txx_o=U(*ni), tyy_o=U(*ni), tzz_o=U(*ni), tyz_o=U(nx, ny + 1, nz + 1), txz_o=U(nx + 1, ny, nz + 1),
        txy_o=U(nx + 1, ny + 1, nz),
        eta=np.asfortranarray(10.0 ** rng.uniform(-3, 0, size=ni)),
        G=np.asfortranarray(rng.uniform(0.5, 2.0, size=ni)),
        K=np.asfortranarray(rng.uniform(1.0, 4.0, size=ni)) if finite_K else np.full(ni, np.inf, order="F"),
        rhogx=U(*ni), rhogy=U(*ni), rhogz=U(*ni),
    )
    if const_rhog is not None:  # spatially constant body force (exercises the constant-field elision of the fused kernel)
        for k, v in zip(("rhogx", "rhogy", "rhogz"), const_rhog):
            f[k] = np.full(ni, float(v), order="F")
    li = (1.0, 1.3, 0.9)
    grid = Geometry(ni, li)
    pt = PTStokesCoeffs(li, grid.di.center)
    return SimpleNamespace(ni=tuple(ni), li=li, di=grid.di.center, grid=grid, igg=IGG(), pt_stokes=pt, dt=dt, fields=f)


# ------------------------------------------------------------------------------------------------------------
def pt_thermal_coeffs_arrays(K, ρCp, dt, di, li, *, ϵ=1.0e-8, CFL=None):
    """PTThermalCoeffs(K, ρCp, dt, di, li; ϵ, CFL) — src/thermal_diffusion/DiffusionPT_coefficients.jl:17-26 (host setup)."""
    CFL = 0.9 / math.sqrt(3) if CFL is None else CFL
    Vpdτ = min(di) * CFL
    max_lxyz = max(li)
    max_lxyz2 = max_lxyz ** 2
    Re = math.pi + np.sqrt(math.pi * math.pi + ρCp * max_lxyz2 / K / dt)
    θr_dτ = np.asfortranarray(max_lxyz / Vpdτ / Re)
    dτ_ρ = np.asfortranarray(Vpdτ * max_lxyz / K / Re)
    return SimpleNamespace(CFL=CFL, ϵ=ϵ, max_lxyz=max_lxyz, max_lxyz2=max_lxyz2, Vpdτ=Vpdτ, θr_dτ=θr_dτ, dτ_ρ=dτ_ρ)


def diffusion2d(nx=32, ny=32, *, lx=100.0e3, ly=100.0e3, ρ0=3.3e3, Cp0=1.2e3, K0=3.0):
    """Config 1 — test/test_diffusion2D.jl:46-125: 2D thermal diffusion, rheology form with a single MaterialParams
    (PT_Density(ρ0=3.1e3, β=0, T0=0, α=1.5e-5), ConstantHeatCapacity(Cp0), ConstantConductivity(K0)), H = 1e-6,
    T(z) linear 1600–1900 K + 100 K disc of radius 10 km, top 300 K / bottom 3500 K, sides no-flux (and
    constant_value = true, i.e. the number 1, overridden by no-flux: quirk Q16), dt = 50 kyr, 20 steps."""
    from .types import TemperatureBoundaryConditions

    kyr = 1.0e3 * 3600 * 24 * 365.25
    dt = 50 * kyr
    ni, li = (nx, ny), (lx, ly)
    di = tuple(l / n for l, n in zip(li, ni))
    grid = Geometry(ni, li, origin=(0.0, -ly))
    xc, yc = grid.xci
    T = np.zeros((nx + 2, ny + 2), order="F")
    T[:, 1:-1] = (yc * (1900.0 - 1600.0) / yc.min() + 1600.0)[None, :]       # init_T!  :29-32
    bc = TemperatureBoundaryConditions(no_flux=dict(left=True, right=True, top=False, bot=False),
                                       constant_value=dict(left=True, right=True, top=300.0, bot=3500.0))
    ρCp = np.full(ni, Cp0 * ρ0, order="F")
    K = np.full(ni, K0, order="F")
    pt = pt_thermal_coeffs_arrays(K, ρCp, dt, di, li, CFL=0.95 / math.sqrt(2.1))
    pert = ((xc[:, None] - lx / 2) ** 2 + (yc[None, :] + ly / 2) ** 2) <= 10.0e3 ** 2                # elliptical_perturbation!
    phases = [dict(rho_kind=1, has_Hr=0, rho0=3.1e3, alpha=1.5e-5, beta=0.0, T0=0.0, P0=0.0, Cp=Cp0, k=K0, Hr=0.0)]
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, dt=dt, nt=20, T=T, bc=bc, pt=pt, perturbation=pert, δT=100.0,
                           H=np.full(ni, 1.0e-6, order="F"), P=np.zeros(ni, order="F"), phases=phases, K=K, ρCp=ρCp,
                           kwargs=dict(iterMax=50e3, nout=1e3, verbose=False))


def diffusion3d(n=32, *, l=100.0e3, ρ0=3.3e3, Cp0=1.2e3, K0=3.0):
    """test/test_diffusion3D.jl:52-141: the 3D twin of config 1 — rheology form with a single MaterialParams (PT_Density(ρ0=3.1e3, β=0, T0=0,
    α=1.5e-5), ConstantHeatCapacity, ConstantConductivity), H = 1e-6, T(z) linear 1600–1900 K + 100 K sphere of radius 10 km, top 300 K /
    bottom 3500 K, sides no-flux, PTThermalCoeffs(K, ρCp, dt, di, li; CFL = 0.95/√3.1), dt = 50 kyr, 10 steps, default kwargs, no thermal_bcs!
    before the loop.  (The reference keeps this test's assertions commented out, :143-156; its golden numbers are used as a soft pin.)"""
    from .types import TemperatureBoundaryConditions

    kyr = 1.0e3 * 3600 * 24 * 365.25
    dt = 50 * kyr
    ni, li = (n, n, n), (l, l, l)
    di = tuple(x / n for x in li)
    grid = Geometry(ni, li, origin=(0.0, 0.0, -l))
    xc, yc, zc = grid.xci
    T = np.zeros((n + 2, n + 2, n + 2), order="F")
    T[:, :, 1:-1] = (zc * (1900.0 - 1600.0) / zc.min() + 1600.0)[None, None, :]          # init_T! over (1:nx+2, 1:ny+2, 1:nz)  :30-33
    bc = TemperatureBoundaryConditions(no_flux=dict(left=True, right=True, top=False, bot=False, front=True, back=True),
                                       constant_value=dict(left=True, right=True, top=300.0, bot=3500.0, front=True, back=True))
    ρCp = np.full(ni, Cp0 * ρ0, order="F")
    K = np.full(ni, K0, order="F")
    pt = pt_thermal_coeffs_arrays(K, ρCp, dt, di, li, CFL=0.95 / math.sqrt(3.1))
    pert = ((xc[:, None, None] - l / 2) ** 2 + (yc[None, :, None] - l / 2) ** 2 + (zc[None, None, :] + l / 2) ** 2) <= 10.0e3 ** 2
    phases = [dict(rho_kind=1, has_Hr=0, rho0=3.1e3, alpha=1.5e-5, beta=0.0, T0=0.0, P0=0.0, Cp=Cp0, k=K0, Hr=0.0)]
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, dt=dt, nt=10, T=T, bc=bc, pt=pt, perturbation=pert, δT=100.0,
                           H=np.full(ni, 1.0e-6, order="F"), P=np.zeros(ni, order="F"), phases=phases, K=K, ρCp=ρCp,
                           kwargs=dict(iterMax=50e3, nout=1e3, verbose=False))


# ------------------------------------------------------------------------------------------------------------
def _smooth2(A, fact):
    """smooth!  miniapps/benchmarks/stokes2D/solcx/SolCx.jl:6-11"""
    A2 = A.copy(order="F")
    c = A[1:-1, 1:-1]
    A2[1:-1, 1:-1] = c + 1.0 / 4.1 / fact * (((A[2:, 1:-1] - c) - (c - A[:-2, 1:-1])) + ((A[1:-1, 2:] - c) - (c - A[1:-1, :-2])))
    return A2


def solcx2d(nx=64, ny=64, *, Δη=1.0e6, lx=1.0, ly=1.0):
    """Config 2 — miniapps/benchmarks/stokes2D/solcx/SolCx.jl:54-145 (variant 2D-V2): η = 1 | Δη across x = 0.5, five
    smoothing passes with edge copies, ρg_y = −sin(πy) cos(πx), G = K = Inf, solve with dt = 0.1, free slip,
    PTStokesCoeffs(li, di; CFL = 1/√2.1, ϵ_abs = 1e-8, ϵ_rel = 1e-9), kwargs = (iterMax = 500e3, nout = 5e3)."""
    ni, li = (nx, ny), (lx, ly)
    grid = Geometry(ni, li, origin=(0.0, 0.0))
    di = grid.di.center
    pt = PTStokesCoeffs(li, di, CFL=1 / math.sqrt(2.1), ϵ_abs=1.0e-8, ϵ_rel=1.0e-9)
    xc, yc = grid.xci
    η = np.asfortranarray(np.where(xc[:, None] <= 0.5, 1.0, Δη) * np.ones((1, ny)))
    ρ = np.asfortranarray(-np.sin(math.pi * yc[None, :]) * np.cos(math.pi * xc[:, None]))
    η2 = η.copy(order="F")
    for _ in range(5):
        η2 = _smooth2(η, 1.0) if True else η2
        # the reference smooths η into η2 (whose interior is overwritten) and then copies the edges
        η2[0, :], η2[-1, :] = η2[1, :], η2[-2, :]
        η2[:, 0], η2[:, -1] = η2[:, 1], η2[:, -2]
        η, η2 = η2, η
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True))
    fields = dict(eta=np.asfortranarray(η), rhogx=np.zeros(ni, order="F"), rhogy=np.asfortranarray(ρ * 1.0), G=np.full(ni, np.inf, order="F"),
                  K=np.full(ni, np.inf, order="F"))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=pt, flow_bcs=flow_bcs, dt=0.1, fields=fields,
                           kwargs=dict(iterMax=500e3, nout=5e3, verbose=False))


def solkz2d(nx=32, ny=32, *, Δη=1.0e6):
    """test/test_stokes_solkz.jl:26-37 + miniapps/benchmarks/stokes2D/solkz/SolKz.jl:4-101 (variant 2D-V2): η = exp(ln(Δη)·y) (six decades
    bottom to top), ρg_y = −sin(2y) cos(3πx), G = K = Inf, dt = 0.1, free slip, PTStokesCoeffs(li, di; Re = 5π, CFL = 1/√2.1),
    kwargs = (iterMax = 150e3, nout = 1e3); the reference test wants err_evo1[end] < 1e-8."""
    ni, li = (nx, ny), (1.0, 1.0)
    grid = Geometry(ni, li, origin=(0.0, 0.0))
    di = grid.di.center
    xc, yc = grid.xci
    η = np.asfortranarray(np.exp(math.log(Δη) * yc)[None, :] * np.ones((nx, 1)))
    ρ = np.asfortranarray(-np.sin(2 * yc)[None, :] * np.cos(3 * math.pi * xc)[:, None])
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True))
    fields = dict(eta=η, rhogx=np.zeros(ni, order="F"), rhogy=ρ * 1.0, G=np.full(ni, np.inf, order="F"), K=np.full(ni, np.inf, order="F"))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=PTStokesCoeffs(li, di, Re=5 * math.pi, CFL=1 / math.sqrt(2.1)),
                           flow_bcs=flow_bcs, dt=0.1, fields=fields, kwargs=dict(iterMax=150.0e3, nout=1.0e3, verbose=False))


def elastic_buildup2d(n=32, *, lx=100.0e3, ly=100.0e3, endtime=10, η0=1.0e21, εbg=1.0e-14, G=10.0e9):
    """test/test_stokes_elastic_buildup.jl:24-53 + miniapps/benchmarks/stokes2D/elastic_buildup/Elastic_BuildUp.jl:19-108 (variant 2D-V2 with
    finite G and dt, K = Inf): uniform η0, pure shear εbg (pureshear_bc!, src/boundaryconditions/pure_shear.jl:1-9), free slip, no gravity,
    PTStokesCoeffs(li, di; ϵ_abs = ϵ_rel = 1e-6, CFL = 1/√2.1), kwargs = (iterMax = 150e3, nout = 1000); time steps of 0.05 kyr below
    10 kyr (1 kyr above).  Analytic stress: τ(t) = 2 εbg η0 (1 − exp(−G t / η0)); the reference test wants the mean relative error of
    max|τyy| over the steps ≤ 5e-3."""
    ni, li = (n, n), (lx, ly)
    grid = Geometry(ni, li, origin=(0.0, 0.0))
    di = grid.di.center
    pt = PTStokesCoeffs(li, di, ϵ_abs=1.0e-6, ϵ_rel=1.0e-6, CFL=1 / math.sqrt(2.1))
    (xc, yc), (xv, yv) = grid.xci, grid.xvi
    Vx = np.zeros((n + 1, n + 2), order="F")
    Vy = np.zeros((n + 2, n + 1), order="F")
    Vx[:, 1:-1] = (εbg * xv)[:, None] * np.ones((1, n))
    Vy[1:-1, :] = np.ones((n, 1)) * (-εbg * yv)[None, :]
    kyr = 1.0e3 * 365.25 * 3600 * 24
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True))
    fields = dict(Vx=Vx, Vy=Vy, eta=np.full(ni, η0, order="F"), rhogx=np.zeros(ni, order="F"), rhogy=np.zeros(ni, order="F"),
                  G=np.full(ni, G, order="F"), K=np.full(ni, np.inf, order="F"))
    dt_of = lambda t: 0.05 * kyr if t < 10 * kyr else 1.0 * kyr
    solution = lambda t: 2 * εbg * η0 * (1 - math.exp(-G * t / η0))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=pt, flow_bcs=flow_bcs, fields=fields, ttot=endtime * kyr,
                           dt_of=dt_of, solution=solution, kwargs=dict(iterMax=150.0e3, nout=1000, verbose=False))


def shearband2d(n=32):
    """Config 3 — test/test_shearband2D.jl:61-192 (variant 2D-VC): unit square, two phases (matrix G = 1, inclusion r = 0.1 with
    G = 0.5, Kb = 4), η = 1, DruckerPrager_regularised(C = 1.6/cosd(30), ϕ = 30, η_vp = 8e-3, Ψ = 0), dt = 0.25, pure shear
    εbg = 1, free slip, PTStokesCoeffs(li, di; ϵ_rel = 1e-6, CFL = 0.75/√2.1), kwargs = (iterMax = 50e3, nout = 100), 10 steps."""
    from . import rheology as R

    ni, li = (n, n), (1.0, 1.0)
    grid = Geometry(ni, li, origin=(0.0, 0.0))
    di = grid.di.center
    η0, G0, εbg = 1.0, 1.0, 1.0
    Gi = G0 / (6.0 - 4.0)
    dt = η0 / G0 / 4.0
    ϕ = 30
    cosd = math.cos(math.radians(ϕ))
    visc = R.LinearViscous(η=η0)
    pl = R.DruckerPrager_regularised(C=1.6 / cosd, ϕ=ϕ, η_vp=8.0e-3, Ψ=0)
    el_bg, el_inc = R.ConstantElasticity(G=G0, Kb=4), R.ConstantElasticity(G=Gi, Kb=4)
    rheology = (R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=0.0), Gravity=R.ConstantGravity(g=0.0),
                                    CompositeRheology=R.CompositeRheology((visc, el_bg, pl)), Elasticity=el_bg),
                R.SetMaterialParams(Phase=2, Density=R.ConstantDensity(ρ=0.0), Gravity=R.ConstantGravity(g=0.0),
                                    CompositeRheology=R.CompositeRheology((visc, el_inc, pl)), Elasticity=el_inc))

    def init_phases(x, y):  # test_shearband2D.jl:37-52
        out = ((x[:, None] - 0.5) ** 2 + (y[None, :] - 0.5) ** 2) > 0.1 ** 2
        r = np.zeros(out.shape + (2,), order="F")
        r[..., 0], r[..., 1] = out, ~out
        return r

    xc, yc = grid.xci
    xv, yv = grid.xvi
    ratios = dict(center=init_phases(xc, yc), vertex=init_phases(xv, yv))
    pt = PTStokesCoeffs(li, di, ϵ_rel=1.0e-6, CFL=0.75 / math.sqrt(2.1))
    Vx = np.asfortranarray((xv * εbg)[:, None] * np.ones((1, n + 2)))
    Vy = np.asfortranarray(np.ones((n + 2, 1)) * (-yv * εbg)[None, :])
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True),
                                          no_slip=dict(left=False, right=False, top=False, bot=False))
    fields = dict(Vx=Vx, Vy=Vy, T=np.zeros((n + 2, n + 2), order="F"))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=pt, flow_bcs=flow_bcs, dt=dt, fields=fields, rheology=rheology,
                           ratios=ratios, nt=10, kwargs=dict(verbose=False, iterMax=50.0e3, nout=1.0e2, viscosity_cutoff=(-math.inf, math.inf)))


def shearband2d_softening(n=32):
    """test/test_shearband2D_softening.jl:61-192: the shear-band setup with cohesion softening — DruckerPrager_regularised(C = 1.6/cosd(30),
    ϕ = 30, η_vp = 8e-3, Ψ = 0, softening_C = NonLinearSoftening(ξ₀ = 1.6, Δ = 0.8)), dt = η0/G0/4/5 = 0.05, five time steps, args carry
    ΔT = 0 (the thermal-stress form of compute_P! with α = 0)."""
    from . import rheology as R

    s = shearband2d(n)
    cosd = math.cos(math.radians(30))
    pl = R.DruckerPrager_regularised(C=1.6 / cosd, ϕ=30, η_vp=8.0e-3, Ψ=0, softening_C=R.NonLinearSoftening(ξ0=1.6, Δ=0.8))
    mats = []
    for m in s.rheology:
        els = tuple(pl if isinstance(e, R.DruckerPrager_regularised) else e for e in m.CompositeRheology.elements)
        mats.append(R.SetMaterialParams(Phase=m.Phase, Density=m.Density, Gravity=m.Gravity, CompositeRheology=R.CompositeRheology(els), Elasticity=m.Elasticity))
    s.rheology = tuple(mats)
    s.dt = s.dt / 5
    s.nt = 5
    s.fields["dTargs"] = np.zeros(s.ni, order="F")
    s.solution = lambda t: 2.0 * 1.0 * 1.0 * (1.0 - math.exp(-1.0 * t / 1.0))   # solution(ε, t, G, η) of the test script
    return s


def sinking_block2d(n=32, *, nsub=8, center_weights="uniform"):
    """test/test_sinking_block.jl:93-200 (variant 2D-VC with buoyancy, SI units): 500 km square, mantle (LinearViscous η = 1e21,
    ConstantDensity 3200) with a 100 km square block (η = 1e23, ρ = 3300) centred at x = 250 km, depth 100 km; no elasticity (G = Kb = Inf),
    g = 9.81, lithostatic initial pressure, free slip, dt = 1, PTStokesCoeffs(li, di; ϵ_rel = 1e-5, CFL = 0.95/√2.1),
    kwargs = (iterMax = 150e3, nout = 1e3).  The reference builds the phase ratios from JustPIC particles (20–40 per cell); here they are
    the volume fractions of nsub² sub-samples per cell (hat-weighted at the vertices)."""
    from . import rheology as R

    ly = 500.0e3
    ni, li = (n, n), (ly, ly)
    grid = Geometry(ni, li, origin=(0.0, -ly))
    di = grid.di.center
    rheology = (R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=3.2e3), Gravity=R.ConstantGravity(g=9.81),
                                    CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0e21),))),
                R.SetMaterialParams(Phase=2, Density=R.ConstantDensity(ρ=3.3e3), Gravity=R.ConstantGravity(g=9.81),
                                    CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0e23),))))
    xc_a, depth_a, r_a = 250.0e3, 100.0e3, 50.0e3
    inside = lambda X, Y: ((X - xc_a) ** 2 <= r_a ** 2) & ((-Y - depth_a) ** 2 <= r_a ** 2)

    def ratios(x, y, hat):
        off = (np.arange(nsub) + 0.5) / nsub - 0.5
        if hat:
            off = off * 2.0
        acc, wsum = np.zeros((x.size, y.size)), 0.0
        for ox in off:
            for oy in off:
                # JustPIC weights every particle of a cell with the bilinear shape function of the node it contributes to — at the
                # centres too (weights 1 … ½ per dimension): center_weights = "bilinear" emulates that, "uniform" is the volume fraction
                w = (1.0 - abs(ox)) * (1.0 - abs(oy)) if (hat or center_weights == "bilinear") else 1.0
                X, Y = np.meshgrid(x + ox * di[0], y + oy * di[1], indexing="ij")
                acc += w * inside(X, Y)
                wsum += w
        f2 = acc / wsum
        return _onehot([1.0 - f2, f2])

    (xc, yc), (xv, yv) = grid.xci, grid.xvi
    rat = dict(center=np.asfortranarray(ratios(xc, yc, False)), vertex=np.asfortranarray(ratios(xv, yv, True)))
    rho = rat["center"][..., 0] * 3.2e3 + rat["center"][..., 1] * 3.3e3
    rhogy = np.asfortranarray(rho * 9.81)                                   # compute_ρg!(ρg[2], phase_ratios, rheology, args)
    P = np.asfortranarray(rhogy * np.abs(yc)[None, :])                      # init_P!: P = ρg·|z|
    pt = PTStokesCoeffs(li, di, ϵ_rel=1.0e-5, CFL=0.95 / math.sqrt(2.1))
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True),
                                          no_slip=dict(left=False, right=False, top=False, bot=False))
    fields = dict(P=P, rhogy=rhogy, T=np.ones((n + 2, n + 2), order="F"))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=pt, flow_bcs=flow_bcs, dt=1.0, fields=fields, rheology=rheology,
                           ratios=rat, kwargs=dict(verbose=False, iterMax=150.0e3, nout=1.0e3, viscosity_cutoff=(-math.inf, math.inf)))


# ------------------------------------------------------------------------------------------------------------
def _onehot(mask_list):
    """stack boolean masks into a ratio array (nodes..., nphases), column-major"""
    r = np.zeros(mask_list[0].shape + (len(mask_list),), order="F")
    for p, m in enumerate(mask_list):
        r[..., p] = m
    return r


def _stag_coords3(grid):
    """node coordinates of the four 3D phase-ratio locations (JustPIC PhaseRatios: center, xy, yz, xz)"""
    (xc, yc, zc), (xv, yv, zv) = grid.xci, grid.xvi
    return dict(center=(xc, yc, zc), xy=(xv, yv, zc), yz=(xc, yv, zv), xz=(xv, yc, zv), vertex=(xv, yv, zv))


def shearband3d(n=16):
    """3D shear band — test/test_shearband3D_MPI.jl:75-210 (variant 3D-VC): unit cube, two phases (matrix G = 1, spherical inclusion
    r = 0.1 with G = 0.5; ν = 0.5 ⇒ Kb = Inf), η = 1, DruckerPrager_regularised(C = 1.6/cosd(30), ϕ = 30, η_vp = 1.25e-2, Ψ = 0),
    dt = 0.25, Vx = x εbg, Vz = −z εbg, free slip, PTStokesCoeffs(li, di; ϵ_rel = 1e-5, Re = 3, r = 0.7, CFL = 0.9/√3.1),
    kwargs = (iterMax = 150e3, nout = 1e3).  Phase ratios are sampled at the staggered nodes (grid-based phases) instead of from particles."""
    from . import rheology as R

    ni, li = (n, n, n), (1.0, 1.0, 1.0)
    grid = Geometry(ni, li, origin=(0.0, 0.0, 0.0))
    di = grid.di.center
    η0, G0, εbg = 1.0, 1.0, 1.0
    Gi = G0 / (6.0 - 4.0)
    dt = η0 / G0 / 4.0
    cosd = math.cos(math.radians(30))
    visc = R.LinearViscous(η=η0)
    pl = R.DruckerPrager_regularised(C=1.6 / cosd, ϕ=30, η_vp=1.25e-2, Ψ=0)
    el_bg, el_inc = R.ConstantElasticity(G=G0, ν=0.5), R.ConstantElasticity(G=Gi, ν=0.5)
    rheology = (R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=0.0), Gravity=R.ConstantGravity(g=0.0),
                                    CompositeRheology=R.CompositeRheology((visc, el_bg, pl)), Elasticity=el_bg),
                R.SetMaterialParams(Phase=2, Density=R.ConstantDensity(ρ=0.0), Gravity=R.ConstantGravity(g=0.0),
                                    CompositeRheology=R.CompositeRheology((visc, el_inc, pl)), Elasticity=el_inc))
    ratios = {}
    for nm, (x, y, z) in _stag_coords3(grid).items():
        out = ((x[:, None, None] - 0.5) ** 2 + (y[None, :, None] - 0.5) ** 2 + (z[None, None, :] - 0.5) ** 2) > 0.1 ** 2
        ratios[nm] = _onehot([out, ~out])
    pt = PTStokesCoeffs(li, di, ϵ_rel=1.0e-5, Re=3.0, r=0.7, CFL=0.9 / math.sqrt(3.1))
    xv, yv, zv = grid.xvi
    Vx = np.asfortranarray(np.broadcast_to((xv * εbg)[:, None, None], (n + 1, n + 2, n + 2)).copy())
    Vz = np.asfortranarray(np.broadcast_to((-zv * εbg)[None, None, :], (n + 2, n + 2, n + 1)).copy())
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True, back=True, front=True),
                                          no_slip=dict(left=False, right=False, top=False, bot=False, back=False, front=False))
    fields = dict(Vx=Vx, Vz=Vz, T=np.zeros((n + 2,) * 3, order="F"))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=pt, flow_bcs=flow_bcs, dt=dt, fields=fields, rheology=rheology,
                           ratios=ratios, nt=3, kwargs=dict(verbose=False, iterMax=150.0e3, nout=1.0e3, viscosity_cutoff=(-math.inf, math.inf)))


def shearband3d_extruded(n=32, ny=4):
    """The reference's 2D shear-band test (test/test_shearband2D.jl:61-192, see shearband2d) extruded along y and run through the 3D
    multiphase solver (variant 3D-VC): 2D (x, y) ↦ 3D (x, z), ny cubic cells deep, cylindrical inclusion, free slip on every face,
    Vx = x εbg, Vz = −z εbg, Vy = 0, the 2D test's rheology (Kb = 4, η_vp = 8e-3), dt and tolerance.  Plane strain in 3D carries the
    out-of-plane deviatoric stress τyy the 2D kernels do not have, so the 2D golden is reproduced up to that term (≈ 1 %)."""
    from . import rheology as R

    ni, li = (n, ny, n), (1.0, ny / n, 1.0)
    grid = Geometry(ni, li, origin=(0.0, 0.0, 0.0))
    di = grid.di.center
    s2 = shearband2d(n)
    ratios = {}
    for nm, (x, y, z) in _stag_coords3(grid).items():
        out = np.broadcast_to((((x[:, None, None] - 0.5) ** 2 + (z[None, None, :] - 0.5) ** 2) > 0.1 ** 2), (len(x), len(y), len(z)))
        ratios[nm] = _onehot([out, ~out])
    pt = PTStokesCoeffs((1.0, 1.0, 1.0), di, ϵ_rel=1.0e-6, CFL=0.75 / math.sqrt(3.1))
    xv, yv, zv = grid.xvi
    Vx = np.asfortranarray(np.broadcast_to((xv * 1.0)[:, None, None], (n + 1, ny + 2, n + 2)).copy())
    Vz = np.asfortranarray(np.broadcast_to((-zv * 1.0)[None, None, :], (n + 2, ny + 2, n + 1)).copy())
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True, back=True, front=True),
                                          no_slip=dict(left=False, right=False, top=False, bot=False, back=False, front=False))
    fields = dict(Vx=Vx, Vz=Vz, T=np.zeros((n + 2, ny + 2, n + 2), order="F"))
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=IGG(), pt_stokes=pt, flow_bcs=flow_bcs, dt=s2.dt, fields=fields, rheology=s2.rheology,
                           ratios=ratios, nt=10, kwargs=dict(verbose=False, iterMax=50.0e3, nout=1.0e2, viscosity_cutoff=(-math.inf, math.inf)))


def convection3d(nx=32, ny=32, nz=32, *, igg: IGG | None = None, plastic=True):
    """Config 5 — 3D thermal convection with grid-based phases, modelled on miniapps/convection/RisingBlob3D/Blob3D.jl:132-201,
    213-300: three phases — crust (LinearViscous + ConstantElasticity + DruckerPrager_regularised, PT_Density),
    a hot low-density blob (LinearViscous + ConstantElasticity, PT_Density) and a weak top layer (LinearViscous, ConstantDensity) —
    ConstantHeatCapacity / ConstantConductivity per phase, free slip, T fixed at top and bottom, no flux on the sides.
    The conductivities differ only mildly between phases and the thermal PT CFL is 0.8/√3.1: the reference's PT scheme sizes dτ_ρ from the
    cell conductivity but fluxes with the 2-cell average, so a conductivity jump ≥ 1.15 at CFL = 0.95/√3.1 exceeds the 3D stability bound
    (observed: the oracle itself diverges at 64³ with k = 1 | 0.5 | 5).
    Deviations from the miniapp (stated in DESIGN.md): non-dimensional O(1) parameters instead of GEO_units scaling, no
    NonLinearSoftening / latent heat / shear heating, phases sampled on the staggered grid instead of from particles.
    Returns Stokes (3D-VC) and thermal (rheology form with phase ratios) inputs for one coupled time step."""
    from . import rheology as R
    from .types import TemperatureBoundaryConditions

    igg = igg or IGG()
    ni, li = (nx, ny, nz), (1.0, 1.0, 1.0)
    grid = Geometry(ni, li, origin=(0.0, 0.0, -1.0), igg=igg)
    di = grid.di.center
    cosd = math.cos(math.radians(30))
    pl = R.DruckerPrager_regularised(C=2.0 / cosd, ϕ=30, η_vp=1.0e-2, Ψ=0)
    el = R.ConstantElasticity(G=10.0, ν=0.25)
    el_blob = R.ConstantElasticity(G=5.0, ν=0.25)
    crust = (R.LinearViscous(η=1.0), el, pl) if plastic else (R.LinearViscous(η=1.0), el)
    rheology = (
        R.SetMaterialParams(Phase=1, Density=R.PT_Density(ρ0=1.0, α=3.0e-2, β=1.0e-3, T0=0.0, P0=0.0), HeatCapacity=R.ConstantHeatCapacity(Cp=1.0),
                            Conductivity=R.ConstantConductivity(k=1.0), CompositeRheology=R.CompositeRheology(crust),
                            Gravity=R.ConstantGravity(g=10.0), Elasticity=el),
        R.SetMaterialParams(Phase=2, Density=R.PT_Density(ρ0=0.9, α=3.0e-2, β=1.0e-3, T0=0.0, P0=0.0), HeatCapacity=R.ConstantHeatCapacity(Cp=1.0),
                            Conductivity=R.ConstantConductivity(k=0.9), CompositeRheology=R.CompositeRheology((R.LinearViscous(η=0.1), el_blob)),
                            Gravity=R.ConstantGravity(g=10.0), Elasticity=el_blob),
        R.SetMaterialParams(Phase=3, Density=R.ConstantDensity(ρ=0.8), HeatCapacity=R.ConstantHeatCapacity(Cp=1.0),
                            Conductivity=R.ConstantConductivity(k=1.1), CompositeRheology=R.CompositeRheology((R.LinearViscous(η=0.01),)),
                            Gravity=R.ConstantGravity(g=10.0)),
    )

    def phase_masks(x, y, z):
        X, Y, Z = x[:, None, None], y[None, :, None], z[None, None, :]
        air = np.broadcast_to(Z > -0.1, (x.size, y.size, z.size))
        blob = (((X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z + 0.6) ** 2) <= 0.15 ** 2) & ~air
        return [~air & ~blob, blob, air]

    locs = _stag_coords3(grid)
    (xc, yc, zc), (xv, yv, zv) = grid.xci, grid.xvi
    locs.update(Vx=(xv, yc, zc), Vy=(xc, yv, zc), Vz=(xc, yc, zv))
    ratios = {nm: _onehot(phase_masks(*xyz)) for nm, xyz in locs.items()}
    # temperature (ghosted, ni.+2): conductive profile 0 (top) … 1 (bottom) + blob anomaly; ghost cells by edge copy
    Tc = (-zc)[None, None, :] * np.ones((nx, ny, 1))
    Tc = Tc + 0.2 * ratios["center"][..., 1]
    T = np.asfortranarray(np.pad(Tc, 1, mode="edge"))
    thermal_bc = TemperatureBoundaryConditions(no_flux=dict(left=True, right=True, front=True, back=True, top=False, bot=False),
                                               constant_value=dict(left=False, right=False, front=False, back=False, top=0.0, bot=1.0))
    flow_bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, top=True, bot=True, back=True, front=True),
                                          no_slip=dict(left=False, right=False, top=False, bot=False, back=False, front=False))
    pt = PTStokesCoeffs(li, di, ϵ_rel=1.0e-5, CFL=0.9 / math.sqrt(3.1))
    fields = dict(T=T)
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, igg=igg, pt_stokes=pt, flow_bcs=flow_bcs, thermal_bc=thermal_bc, dt=0.05, fields=fields,
                           rheology=rheology, ratios=ratios, T=T,
                           kwargs=dict(verbose=False, iterMax=150.0e3, nout=1.0e3, viscosity_cutoff=(1.0e-3, 1.0e3)),
                           thermal_kwargs=dict(iterMax=150.0e3, nout=1.0e3, verbose=False), thermal_CFL=0.8 / math.sqrt(3.1))


def random_vc3d(ni, nphase=3, seed=20261017, *, dt=0.4, mixed=True):
    """Seeded random state for kernel-level parity fuzzing of variant 3D-VC: V, τ ~ U(−1,1), P ~ U(0,1), η ~ 10^U(−2,0),
    phase ratios Dirichlet(1,…) with exact zeros and ones mixed in; phases: plastic + compressible, elastic only, viscous only."""
    from . import rheology as R

    rng = np.random.default_rng(seed)
    nx, ny, nz = ni
    U = lambda *s: np.asfortranarray(rng.uniform(-1.0, 1.0, size=s))
    cshape = dict(center=ni, xy=(nx + 1, ny + 1, nz), yz=(nx, ny + 1, nz + 1), xz=(nx + 1, ny, nz + 1))
    f = dict(Vx=U(nx + 1, ny + 2, nz + 2), Vy=U(nx + 2, ny + 1, nz + 2), Vz=U(nx + 2, ny + 2, nz + 1),
             P=np.asfortranarray(rng.uniform(0, 1, size=ni)), Q=np.asfortranarray(rng.uniform(-0.1, 0.1, size=ni)),
             eta=np.asfortranarray(10.0 ** rng.uniform(-2, 0, size=ni)), EII_pl=np.asfortranarray(rng.uniform(0, 0.1, size=ni)),
             T=np.asfortranarray(rng.uniform(0, 1, size=(nx + 2, ny + 2, nz + 2))), rhogx=U(*ni), rhogy=U(*ni), rhogz=U(*ni))
    for pre in ("t", "e"):
        for c in ("xx", "yy", "zz"):
            f[f"{pre}{c}"] = U(*ni) * (0.3 if pre == "t" else 1.0)
        for c in ("yz", "xz", "xy"):
            f[f"{pre}{c}"] = U(*cshape[c]) * (0.3 if pre == "t" else 1.0)
    for c in ("xx", "yy", "zz"):
        f[f"t{c}_o"] = U(*ni) * 0.3
    for c in ("yz", "xz", "xy"):
        f[f"t{c}_o"] = U(*cshape[c]) * 0.3
        f[f"t{c}_c"] = U(*ni) * 0.3
        f[f"t{c}_o_c"] = U(*ni) * 0.3
    ratios = {}
    for nm, sh in cshape.items():
        r = rng.dirichlet(np.ones(nphase), size=sh)
        if mixed:
            pick = rng.integers(0, 3, size=sh)            # 0: keep mixture, 1: one-hot, 2: two-phase mixture with an exact zero
            hot = np.eye(nphase)[rng.integers(0, nphase, size=sh)]
            two = r.copy()
            two[..., -1] = 0.0
            two /= two.sum(axis=-1, keepdims=True)
            r = np.where((pick == 1)[..., None], hot, np.where((pick == 2)[..., None], two, r))
        else:
            r = np.eye(nphase)[rng.integers(0, nphase, size=sh)]
        ratios[nm] = np.asfortranarray(r)
    cosd = math.cos(math.radians(30))
    pl = R.DruckerPrager_regularised(C=0.05 / cosd, ϕ=30, η_vp=1.0e-2, Ψ=5)
    el1, el2 = R.ConstantElasticity(G=1.0, Kb=4.0), R.ConstantElasticity(G=0.5, ν=0.5)
    mats = [
        R.SetMaterialParams(Phase=1, Density=R.PT_Density(ρ0=1.0, α=3.0e-2, β=1.0e-2, T0=0.1, P0=0.0), Gravity=R.ConstantGravity(g=1.0),
                            CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0), el1, pl)), Elasticity=el1),
        R.SetMaterialParams(Phase=2, Density=R.ConstantDensity(ρ=0.5), Gravity=R.ConstantGravity(g=1.0),
                            CompositeRheology=R.CompositeRheology((R.LinearViscous(η=0.1), el2)), Elasticity=el2),
        R.SetMaterialParams(Phase=3, Density=R.ConstantDensity(ρ=0.1), Gravity=R.ConstantGravity(g=1.0),
                            CompositeRheology=R.CompositeRheology((R.LinearViscous(η=0.01),))),
    ][:nphase]
    li = (1.0, 1.3, 0.9)
    grid = Geometry(ni, li)
    pt = PTStokesCoeffs(li, grid.di.center, CFL=0.9 / math.sqrt(3.1))
    return SimpleNamespace(ni=tuple(ni), li=li, di=grid.di.center, grid=grid, igg=IGG(), pt_stokes=pt, dt=dt, fields=f, ratios=ratios,
                           rheology=tuple(mats), kwargs=dict(viscosity_cutoff=(1.0e-3, 1.0e2)))


def diffusion_multiphase(nd=3, n=32, *, nsub=8):
    """test/test_diffusion3D_multiphase.jl:88-203 (nd = 3) and test/test_diffusion2D_multiphase.jl:81-181 (nd = 2): heatdiffusion_PT! in the
    rheology form with TWO phases (PT_Density ρ0 = 3.0e3 | 3.3e3, Cp = 1.2e3, k = 3, ConstantRadioactiveHeat 1e-6 | 1e-7; phase 2 inside a
    sphere/disc of radius 10 km at the domain centre), T(z) linear 1600–1900 K + 100 K inside the same sphere, top 300 K / bottom 3500 K,
    sides no-flux, dt = 50 kyr.  The reference builds the phase ratios from JustPIC particles (20–40 per cell); here they are the
    (hat-weighted at the face nodes) volume fractions from nsub^nd sub-samples per cell — the goldens hold to the reference tolerance."""
    from .types import TemperatureBoundaryConditions

    kyr = 1.0e3 * 3600 * 24 * 365.25
    dt = 50 * kyr
    ni, li = (n,) * nd, (100.0e3,) * nd
    origin = (0.0,) * (nd - 1) + (-li[-1],)
    grid = Geometry(ni, li, origin=origin)
    di = grid.di.center
    xci = grid.xci
    zc = xci[-1]
    T = np.zeros(tuple(m + 2 for m in ni), order="F")
    prof = zc * (1900.0 - 1600.0) / zc.min() + 1600.0
    centre = tuple(0.5 * l for l in li[:-1]) + (-0.5 * li[-1],)
    r = 10.0e3
    inside = lambda *X: sum((x - c) ** 2 for x, c in zip(X, centre)) <= r ** 2
    mesh = lambda cs: np.meshgrid(*cs, indexing="ij")
    pert = inside(*mesh(xci))
    if nd == 3:
        T[:, :, 1:-1] = prof[None, None, :]          # init_T! over (1:nx+2, 1:ny+2, 1:nz)
        bc = TemperatureBoundaryConditions(no_flux=dict(left=True, right=True, top=False, bot=False, front=True, back=True),
                                           constant_value=dict(left=True, right=True, top=300.0, bot=3500.0, front=True, back=True))
        H, ϵ, CFL, nt = 1.0e-6, 1.0e-8, 0.95 / math.sqrt(3.1), 10
        kwargs = dict(iterMax=10.0e3, nout=1.0e2, verbose=False)
    else:
        T[:, 1:-1] = prof[None, :]
        bc = TemperatureBoundaryConditions(no_flux=dict(left=True, right=True, top=False, bot=False),
                                           constant_value=dict(left=True, right=True, top=300.0, bot=3500.0))
        H, ϵ, CFL, nt = 0.0, 1.0e-5, 0.95 / math.sqrt(2), 20
        kwargs = dict(iterMax=1.0e3, nout=10, verbose=False)

    def ratios(coords, hat):
        """fraction of phase 2 around every node: uniform sub-samples of the cell (centres) or hat-weighted over ±d (face nodes)"""
        off = (np.arange(nsub) + 0.5) / nsub - 0.5
        if hat:
            off = off * 2.0
        acc = np.zeros(tuple(c.size for c in coords))
        wsum = 0.0
        for idx in np.ndindex(*(nsub,) * nd):
            o = [off[i] for i in idx]
            w = float(np.prod([1.0 - abs(v) for v in o])) if hat else 1.0
            X = mesh([c + o[q] * di[q] for q, c in enumerate(coords)])
            acc += w * inside(*X)
            wsum += w
        f2 = acc / wsum
        return _onehot([1.0 - f2, f2])

    xvi = grid.xvi
    locs = dict(center=xci)
    for a, nm in enumerate(("Vx", "Vy", "Vz")[:nd]):
        locs[nm] = tuple(xvi[q] if q == a else xci[q] for q in range(nd))
    phase = {nm: np.asfortranarray(ratios(c, nm != "center")) for nm, c in locs.items()}
    rows = [dict(rho_kind=1, has_Hr=1, rho0=3.0e3, alpha=1.5e-5, beta=0.0, T0=0.0, P0=0.0, Cp=1.2e3, k=3.0, Hr=1.0e-6),
            dict(rho_kind=1, has_Hr=1, rho0=3.3e3, alpha=1.5e-5, beta=0.0, T0=0.0, P0=0.0, Cp=1.2e3, k=3.0, Hr=1.0e-7)]
    ρCp = np.full(ni, 1.2e3 * 3.3e3, order="F")
    K = np.full(ni, 3.0, order="F")
    pt = pt_thermal_coeffs_arrays(K, ρCp, dt, di, li, ϵ=ϵ, CFL=CFL)
    return SimpleNamespace(ni=ni, li=li, di=di, grid=grid, dt=dt, nt=nt, T=T, bc=bc, pt=pt, perturbation=pert, δT=100.0, phase=phase, phases=rows,
                           H=np.full(ni, H, order="F"), P=np.zeros(ni, order="F"), kwargs=kwargs, thermal_bcs_first=(nd == 2))
