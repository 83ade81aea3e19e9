"""Lowering of the supported GeoParams subset to the flat device table (no GPU): the parameter look-ups the reference pins in
test/test_rheology.jl:97-116 (get_bulk_modulus ≈ 5e10, get_shear_modulus ≈ 1e10, Inf when no elastic element, α of T_/PT_Density ≈ 3e-5,
α of ConstantDensity == 0) and the loud failure for laws outside the subset (north star: no fallback)."""
import math

import pytest

from justrelax_jl_b200 import rheology as R


def test_parameter_lookups_match_reference_kats():
    elastic = R.ConstantElasticity(G=1.0e10, Kb=5.0e10)
    mat_T = R.SetMaterialParams(Phase=1, Density=R.T_Density(ρ0=2900.0, α=3.0e-5, T0=273.0), Elasticity=elastic,
                                CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0e21), elastic)))
    mat_PT = R.SetMaterialParams(Phase=2, Density=R.PT_Density(ρ0=2900.0, α=3.0e-5, β=1.0e-9, T0=273.0, P0=0.0),
                                 CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0e21), elastic)), Elasticity=elastic)
    mat_no_elastic = R.SetMaterialParams(Phase=3, Density=R.ConstantDensity(ρ=3000.0),
                                         CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0e21),)))
    rows = R.lower_stokes((mat_T, mat_PT, mat_no_elastic))
    assert rows[0]["Kb"] == pytest.approx(5.0e10) and rows[0]["G"] == pytest.approx(1.0e10)      # test_rheology.jl:107-108
    assert rows[0]["alpha"] == pytest.approx(3.0e-5) and rows[0]["rho_kind"] == 2                   # :61-62
    assert rows[1]["alpha"] == pytest.approx(3.0e-5) and rows[1]["rho_kind"] == 1                   # :63-64
    assert rows[2]["Kb"] == math.inf and rows[2]["G"] == math.inf                                   # :114-115  (=== Inf)
    assert rows[2]["alpha"] == 0 and rows[2]["rho_kind"] == 0                                       # :57-58
    assert all(r["eta"] == 1.0e21 and r["has_pl"] == 0 for r in rows)


def test_drucker_prager_row_and_first_plastic_element_wins():
    pl1 = R.DruckerPrager_regularised(C=1.6, ϕ=30, η_vp=8.0e-3, Ψ=5)
    pl2 = R.DruckerPrager(C=9.9, ϕ=10, Ψ=0)
    el = R.ConstantElasticity(G=1.0, Kb=4.0)
    m = R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=1.0), Elasticity=el,
                            CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0), el, pl1, pl2)))
    (row,) = R.lower_stokes((m,))
    assert row["has_pl"] == 1 and row["C"] == 1.6 and row["eta_vp"] == 8.0e-3                      # StressUpdate.jl:131-144
    assert row["sinphi"] == pytest.approx(math.sin(math.radians(30))) and row["cosphi"] == pytest.approx(math.cos(math.radians(30)))
    assert row["sinpsi"] == pytest.approx(math.sin(math.radians(5)))


def test_laws_outside_the_subset_fail_loudly():
    class DislocationCreep:   # stands for any GeoParams creep law that is not lowered
        pass

    m = R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=1.0), CompositeRheology=R.CompositeRheology((DislocationCreep(),)))
    with pytest.raises(R.UnsupportedRheology):
        R.lower_stokes((m,))
    two = R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=1.0),
                              CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0), R.LinearViscous(η=2.0))))
    with pytest.raises(R.UnsupportedRheology):
        R.lower_stokes((two,))
