"""GeoParams.jl material laws — the supported subset — and their lowering to the flat per-phase tables libjrb200 takes.

The reference hands `rheology::NTuple{N,MaterialParams}` (GeoParams.jl structs) to its kernels, which dispatch on the
law types point-wise (src/rheology/*.jl, src/thermal_diffusion/DiffusionPT_GeoParams.jl).  The B200 backend lowers the
tuple ONCE per solve to plain rows (`jr_thermal_phase`, `jr_stokes_phase` in include/jrb200.h); a law outside the subset
raises at lowering time — there is no fallback (SURVEY.md §8b, Appendix C lists the struct fields read).

Names, keyword arguments and defaults follow GeoParams 0.7 (SI values, no unit handling: pass numbers).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple


class UnsupportedRheology(ValueError):
    """raised when a material law is outside the subset the device table supports (JR_ERR_UNSUPPORTED on the C side)"""


# ---- density ----------------------------------------------------------------------------------------------------
@dataclass
class ConstantDensity:
    ρ: float = 2900.0


@dataclass
class PT_Density:
    """ρ = ρ0 (1 − α (T − T0) + β (P − P0))"""
    ρ0: float = 2900.0
    α: float = 3e-5
    β: float = 1e-9
    T0: float = 0.0
    P0: float = 0.0


@dataclass
class T_Density:
    """ρ = ρ0 (1 − α (T − T0))"""
    ρ0: float = 2900.0
    α: float = 3e-5
    T0: float = 273.15


# ---- thermal ------------------------------------------------------------------------------------------------------
@dataclass
class ConstantHeatCapacity:
    Cp: float = 1050.0


@dataclass
class ConstantConductivity:
    k: float = 3.0


@dataclass
class TP_Conductivity:
    """GeoParams TP_Conductivity: k(T, P) = (a + b / (T + c)) · (1 + d · P)  (T in the units of the model, P likewise; the miniapps pass
    d per Pa: miniapps/convection/Particles3D/Layered_rheology.jl:45-57, benchmarks/stokes2D/shear_heating/Shearheating_rheology.jl:9-17).
    Defaults = GeoParams' (a = 1.18 W/m/K, b = 474 W/m, c = 77 K, d = 0)."""
    a: float = 1.18
    b: float = 474.0
    c: float = 77.0
    d: float = 0.0


@dataclass
class ConstantRadioactiveHeat:
    H_r: float = 1e-6


@dataclass
class ConstantShearheating:
    """GeoParams ConstantShearheating(Χ): H_s = Χ·τij(εij − εij_el)"""
    Χ: float = 0.0


@dataclass
class ConstantGravity:
    g: float = 9.81


# ---- creep / elasticity / plasticity ---------------------------------------------------------------------------------
@dataclass
class LinearViscous:
    η: float = 1e20


@dataclass
class ConstantElasticity:
    """G shear modulus, ν Poisson ratio, Kb bulk modulus (GeoParams: Kb = 2G(1+ν)/(3(1−2ν)) unless given)"""
    G: float = 5e10
    ν: float = 0.5
    Kb: Optional[float] = None

    def __post_init__(self):
        if self.Kb is None:
            den = 3 * (1 - 2 * self.ν)
            self.Kb = math.inf if den == 0 else 2 * self.G * (1 + self.ν) / den


@dataclass
class LinearSoftening:
    """GeoParams LinearSoftening((min_value, max_value), (lo, hi)): max_value below lo, min_value above hi, linear in between"""
    min_max_values: tuple = (0.0, 0.0)
    lo_hi: tuple = (0.0, 1.0)

    def params(self):
        mn, mx = (float(v) for v in self.min_max_values)
        lo, hi = (float(v) for v in self.lo_hi)
        slope = (mx - mn) / (lo - hi)
        return 1, [lo, hi, mx, mn, slope, mx - slope * lo]


@dataclass
class NonLinearSoftening:
    """GeoParams NonLinearSoftening(ξ₀, Δ, μ = 1, σ = 0.5): ξ₀ − ½·Δ·erfc(−(x − μ)/σ)"""
    ξ0: float = 0.0
    Δ: float = 0.0
    μ: float = 1.0
    σ: float = 0.5

    def params(self):
        return 2, [float(self.ξ0), float(self.Δ), float(self.μ), float(self.σ), 0.0, 0.0]


@dataclass
class DruckerPrager_regularised:
    """F = τII − C cosϕ − P sinϕ; ϕ, Ψ in degrees; η_vp regularisation viscosity"""
    C: float = 10e6
    ϕ: float = 30.0
    Ψ: float = 0.0
    η_vp: float = 1e20
    softening_C: object = None
    softening_ϕ: object = None

    @property
    def sinϕ(self):
        return math.sin(math.radians(self.ϕ))

    @property
    def cosϕ(self):
        return math.cos(math.radians(self.ϕ))

    @property
    def sinΨ(self):
        return math.sin(math.radians(self.Ψ))


@dataclass
class DruckerPrager(DruckerPrager_regularised):
    η_vp: float = 0.0


@dataclass
class CompositeRheology:
    elements: Tuple

    def __init__(self, elements):
        self.elements = tuple(elements)


@dataclass
class MaterialParams:
    Phase: int = 1
    Density: object = None
    HeatCapacity: object = None
    Conductivity: object = None
    RadioactiveHeat: object = None
    ShearHeat: object = None
    CompositeRheology: object = None
    Gravity: object = field(default_factory=ConstantGravity)
    Elasticity: object = None


def SetMaterialParams(**kw) -> MaterialParams:
    return MaterialParams(**kw)


# ---- lowering -----------------------------------------------------------------------------------------------------------
def _as_tuple(rheology) -> Sequence[MaterialParams]:
    return tuple(rheology) if isinstance(rheology, (tuple, list)) else (rheology,)


def lower_thermal(rheology):
    """rows of jr_thermal_phase: dict(rho_kind, has_Hr, rho0, alpha, beta, T0, P0, Cp, k, Hr, k_kind, k_a, k_b, k_c, k_d)"""
    rows = []
    for p in _as_tuple(rheology):
        ρ = p.Density
        if isinstance(ρ, ConstantDensity):
            row = dict(rho_kind=0, rho0=ρ.ρ, alpha=0.0, beta=0.0, T0=0.0, P0=0.0)
        elif isinstance(ρ, PT_Density):
            row = dict(rho_kind=1, rho0=ρ.ρ0, alpha=ρ.α, beta=ρ.β, T0=ρ.T0, P0=ρ.P0)
        elif isinstance(ρ, T_Density):
            row = dict(rho_kind=2, rho0=ρ.ρ0, alpha=ρ.α, beta=0.0, T0=ρ.T0, P0=0.0)
        else:
            raise UnsupportedRheology(f"density law {type(ρ).__name__} is outside the supported subset")
        if not isinstance(p.HeatCapacity, ConstantHeatCapacity):
            raise UnsupportedRheology(f"heat-capacity law {type(p.HeatCapacity).__name__} is outside the supported subset")
        row["Cp"] = p.HeatCapacity.Cp
        row.update(k=0.0, k_kind=0, k_a=0.0, k_b=0.0, k_c=0.0, k_d=0.0)
        if isinstance(p.Conductivity, ConstantConductivity):
            row["k"] = p.Conductivity.k
        elif isinstance(p.Conductivity, TP_Conductivity):
            κ = p.Conductivity
            row.update(k_kind=1, k_a=float(κ.a), k_b=float(κ.b), k_c=float(κ.c), k_d=float(κ.d))
        else:
            raise UnsupportedRheology(f"conductivity law {type(p.Conductivity).__name__} is outside the supported subset")
        if p.RadioactiveHeat is None:
            row["has_Hr"], row["Hr"] = 0, 0.0
        elif isinstance(p.RadioactiveHeat, ConstantRadioactiveHeat):
            row["has_Hr"], row["Hr"] = 1, p.RadioactiveHeat.H_r
        else:
            raise UnsupportedRheology(f"radioactive-heat law {type(p.RadioactiveHeat).__name__} is outside the supported subset")
        rows.append(row)
    return rows


def _elasticity_of(p: MaterialParams):
    if p.Elasticity is not None:
        return p.Elasticity
    if p.CompositeRheology is not None:
        for e in p.CompositeRheology.elements:
            if isinstance(e, ConstantElasticity):
                return e
    return None


def _modulus(x):
    """get_shear_modulus / get_bulk_modulus: Inf when NaN or zero (src/rheology/GeoParams.jl:1-15)"""
    return math.inf if (x is None or x != x or x == 0) else float(x)


def lower_stokes(rheology, ndim: int = 2):
    """rows of jr_stokes_phase: dict(eta, G, Kb, has_pl, rho_kind, C, sinphi, cosphi, sinpsi, eta_vp, rho0, alpha, beta, T0, P0, soft_C_kind,
    soft_C).  Supported: LinearViscous (exactly one creep element), ConstantElasticity, DruckerPrager[_regularised] (the FIRST plastic element
    wins, StressUpdate.jl:131-144) with cohesion softening (Linear/NonLinearSoftening of C with the accumulated plastic strain,
    StressUpdate.jl:305-332 — 2D solves only), Constant/PT_/T_Density."""
    rows = []
    for p in _as_tuple(rheology):
        if p.CompositeRheology is None:
            raise UnsupportedRheology("MaterialParams without a CompositeRheology")
        visc = [e for e in p.CompositeRheology.elements if isinstance(e, LinearViscous)]
        other = [e for e in p.CompositeRheology.elements
                 if not isinstance(e, (LinearViscous, ConstantElasticity, DruckerPrager_regularised))]
        if other:
            raise UnsupportedRheology(f"rheological element {type(other[0]).__name__} is outside the supported subset")
        if len(visc) != 1:
            raise UnsupportedRheology("exactly one LinearViscous element per phase is supported")
        el = _elasticity_of(p)
        row = dict(eta=float(visc[0].η), G=_modulus(el.G if el else None), Kb=_modulus(el.Kb if el else None))
        pls = [e for e in p.CompositeRheology.elements if isinstance(e, DruckerPrager_regularised)]
        if pls:
            pl = pls[0]
            if pl.softening_ϕ is not None:
                raise UnsupportedRheology("friction-angle softening is outside the supported subset (SURVEY §8f-1)")
            kind, par = 0, [0.0] * 6
            if pl.softening_C is not None:
                if not isinstance(pl.softening_C, (LinearSoftening, NonLinearSoftening)):
                    raise UnsupportedRheology(f"softening law {type(pl.softening_C).__name__} is outside the supported subset")
                if ndim != 2:
                    raise UnsupportedRheology("cohesion softening is supported by the 2D multiphase solve only (SURVEY §8f-1)")
                kind, par = pl.softening_C.params()
            row.update(has_pl=1, C=float(pl.C), sinphi=pl.sinϕ, cosphi=pl.cosϕ, sinpsi=pl.sinΨ, eta_vp=float(pl.η_vp), soft_C_kind=kind, soft_C=par)
        else:
            row.update(has_pl=0, C=0.0, sinphi=0.0, cosphi=0.0, sinpsi=0.0, eta_vp=0.0, soft_C_kind=0, soft_C=[0.0] * 6)
        ρ = p.Density
        if ρ is None or isinstance(ρ, ConstantDensity):
            row.update(rho_kind=0, rho0=(ρ.ρ if ρ else 0.0), alpha=0.0, beta=0.0, T0=0.0, P0=0.0)
        elif isinstance(ρ, PT_Density):
            row.update(rho_kind=1, rho0=ρ.ρ0, alpha=ρ.α, beta=ρ.β, T0=ρ.T0, P0=ρ.P0)
        elif isinstance(ρ, T_Density):
            row.update(rho_kind=2, rho0=ρ.ρ0, alpha=ρ.α, beta=0.0, T0=ρ.T0, P0=0.0)
        else:
            raise UnsupportedRheology(f"density law {type(ρ).__name__} is outside the supported subset")
        rows.append(row)
    return rows


def gravity_of(rheology):
    """compute_gravity(first(rheology)) (BuoyancyForces.jl:25,56): ConstantGravity(g) acts along the LAST axis (0, 0, g)."""
    g = _as_tuple(rheology)[0].Gravity
    gv = g.g if g is not None else 0.0
    return (0.0, 0.0, float(gv))


def shear_heating_coefficients(rheology):
    """Χ per phase (ShearHeat = ConstantShearheating(Χ); absent → 0, like GeoParams' empty ShearHeat tuple)"""
    out = []
    for p in _as_tuple(rheology):
        sh = p.ShearHeat
        if sh is None:
            out.append(0.0)
        elif isinstance(sh, ConstantShearheating):
            out.append(float(sh.Χ))
        else:
            raise UnsupportedRheology(f"shear-heating law {type(sh).__name__} is outside the supported subset")
    return out
