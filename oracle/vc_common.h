/*
 * vc_common.h — CPU ORACLE (test infrastructure, NOT product code).
 * Phase-ratio weighted rheology helpers shared by the multiphase (VC) Stokes restatements in 2D and 3D:
 *   src/phases/phases.jl:5-30 (fn_ratio), src/rheology/StressUpdate.jl:146-188 (plastic_params_phase),
 *   :384-452 (compute_yieldfunction_phase), :463-550 (compute_plastic_gradients_phase),
 *   src/rheology/Viscosity.jl:599-619 (compute_phase_viscosity), src/rheology/GeoParams.jl:1-15 (moduli),
 *   GeoParams.jl 0.7.19 (third party, not vendored): DruckerPrager F/∂Q∂τ/∂Q∂P/∂F∂P, PT_Density, second_invariant.
 */
#ifndef JR_ORACLE_VC_COMMON_H
#define JR_ORACLE_VC_COMMON_H
#include "jr_oracle.h"

/* ---- rheology table helpers (VC) ----------------------------------------------------------------------------------- */
static inline double second_invariant3(double xx, double yy, double xy) { return sqrt(0.5 * (xx * xx + yy * yy) + xy * xy); }
/* fn_ratio(fn, rheology, ratio)  phases.jl:5-16 */
static inline double ratio_G(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx)
{
    double x = 0.0;
    for (int p = 0; p < vc->nphase; p++) { const double r = ph[(size_t)p * stride + idx]; x += (r == 0.0) ? 0.0 : vc->phases[p].G * r; }
    return x;
}
static inline double ratio_Kb(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx)
{
    double x = 0.0;
    for (int p = 0; p < vc->nphase; p++) { const double r = ph[(size_t)p * stride + idx]; x += (r == 0.0) ? 0.0 : vc->phases[p].Kb * r; }
    return x;
}
/* fn_ratio(get_thermal_expansion, rheology, ratio)  rheology/GeoParams.jl:17: α of the density law, 0 for ConstantDensity */
static inline double ratio_alpha(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx)
{
    double x = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        x += (r == 0.0) ? 0.0 : (vc->phases[p].rho_kind == 0 ? 0.0 : vc->phases[p].alpha) * r;
    }
    return x;
}
/* _compute_P! with thermal stresses  PressureKernels.jl:197-206 */
static inline void P_point_dT(double *RP, double *P, double P0, double divV, double Q, double dT, double alpha, double eta, double K, double G,
                              double dt, double r, double theta_dtau)
{
    const double _Kdt = orc_inv(K * dt), _Gdt = orc_inv(G * dt), _dt = orc_inv(dt), Pc = *P;
    *RP = fma(-(Pc - P0), _Kdt, (-divV + (alpha * (dT * _dt)) + (Q * _dt)));
    const double psi = orc_inv(orc_inv(eta) + _Gdt) * r / theta_dtau;
    *P = ((fma(P0, _Kdt, (-divV + (alpha * (dT * _dt)) + (Q * _dt)))) * psi + Pc) / (1 + _Kdt * psi);
}
/* plastic_params_phase  StressUpdate.jl:153-176: is_pl if any phase with non-zero ratio is plastic; η_reg = Σ η_vp·ratio */
static inline void plastic_params(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx, int *is_pl, double *eta_reg)
{
    *is_pl = 0; *eta_reg = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        const int pl = (r != 0.0) && vc->phases[p].has_pl;
        if (pl) *is_pl = 1;
        *eta_reg += (pl ? vc->phases[p].eta_vp : 0.0) * r;
    }
}
/* compute_yieldfunction_phase  StressUpdate.jl:384-452: Σ r·F_phase (non-plastic phase: F = τII), zero ratios skipped */
/* soften_cohesion  StressUpdate.jl:305-332 → GeoParams LinearSoftening / NonLinearSoftening (third party, restated) */
static inline double soften_C(const orc_stokes_phase *q, double EII)
{
    if (q->soft_C_kind == 1) {
        if (EII >= q->soft_C[1]) return q->soft_C[3];
        if (EII <= q->soft_C[0]) return q->soft_C[2];
        return EII * q->soft_C[4] + q->soft_C[5];
    }
    if (q->soft_C_kind == 2) return q->soft_C[0] - 0.5 * q->soft_C[1] * erfc(-(EII - q->soft_C[2]) / q->soft_C[3]);
    return q->C;
}
static inline double yield_F_soft(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx, double P, double tII, double EII)
{
    double acc = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        double v = 0.0;
        if (r != 0.0) {
            const orc_stokes_phase *q = &vc->phases[p];
            const double Fp = q->has_pl ? (tII - q->cosphi * soften_C(q, EII) - q->sinphi * (P - 0.0)) - 2 * q->eta_vp * (0.0 * 0.5) : tII;
            v = r * Fp;
        }
        acc = p == 0 ? v : acc + v;
    }
    return acc;
}
static inline double yield_F(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx, double P, double tII)
{
    double acc = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        double v = 0.0;
        if (r != 0.0) {
            const orc_stokes_phase *q = &vc->phases[p];
            const double Fp = q->has_pl ? (tII - q->cosphi * q->C - q->sinphi * (P - 0.0)) - 2 * q->eta_vp * (0.0 * 0.5) : tII;
            v = r * Fp;
        }
        acc = p == 0 ? v : acc + v;
    }
    return acc;
}
/* compute_plastic_gradients_phase  StressUpdate.jl:463-550 (muladd → fma); t = trial stress (xx, yy, xy) */
static inline void plastic_grads(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx, const double t[3], double dQdt[3], double *dQdP,
                          double *dFdP)
{
    dQdt[0] = dQdt[1] = dQdt[2] = 0.0; *dQdP = 0.0; *dFdP = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        if (r == 0.0) continue;
        const orc_stokes_phase *q = &vc->phases[p];
        double g[3] = {0, 0, 0}, qp = 0.0, fp = 0.0;
        if (q->has_pl) {
            const double tII = second_invariant3(t[0], t[1], t[2]);
            g[0] = 0.5 * t[0] / tII; g[1] = 0.5 * t[1] / tII; g[2] = 0.5 * (t[2] / tII);
            qp = -q->sinpsi; fp = -q->sinphi;
        }
        for (int c = 0; c < 3; c++) dQdt[c] = fma(r, g[c], dQdt[c]);
        *dQdP = fma(r, qp, *dQdP);
        *dFdP = fma(r, fp, *dFdP);
    }
}
static inline double density(const orc_stokes_phase *p, double T, double P)
{
    if (p->rho_kind == 1) return p->rho0 * (1.0 - p->alpha * (T - p->T0) + p->beta * (P - p->P0));
    if (p->rho_kind == 2) return p->rho0 * (1.0 - p->alpha * (T - p->T0));
    return p->rho0;
}

static inline int density_is_constant(const orc_vc_inputs *vc)
{
    for (int p = 0; p < vc->nphase; p++) if (vc->phases[p].rho_kind != 0) return 0;
    return 1;
}

/* compute_viscosity_kernel! (τII form), centres and — 2D only — vertices (quirk Q19)  Viscosity.jl:282-323, 382-418 */
static inline double phase_viscosity(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx)
{
    /* compute_phase_viscosity  Viscosity.jl:599-619; per-phase composite: LinearViscous + elastic element at dt = Inf */
    for (int p = 0; p < vc->nphase; p++)
        if (ph[(size_t)p * stride + idx] > 0.999) return orc_inv(orc_inv(vc->phases[p].eta) + orc_inv(vc->phases[p].G * INFINITY));
    double e = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        if (r != 0.0) e += orc_inv(orc_inv(orc_inv(vc->phases[p].eta) + orc_inv(vc->phases[p].G * INFINITY))) * r;
    }
    return orc_inv(e);
}
static inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* second_invariant(xx, yy, zz, yz, xz, xy)  (GeoParams) */
static inline double second_invariant6(const double t[6])
{
    return sqrt(0.5 * (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) + t[3] * t[3] + t[4] * t[4] + t[5] * t[5]);
}
/* compute_plastic_gradients_phase, 3D Voigt order (xx, yy, zz, yz, xz, xy): shear slots halved (StressUpdate.jl:467-472) */
static inline void plastic_grads6(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx, const double t[6], double dQdt[6], double *dQdP,
                                  double *dFdP)
{
    for (int c = 0; c < 6; c++) dQdt[c] = 0.0;
    *dQdP = 0.0; *dFdP = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        if (r == 0.0) continue;
        const orc_stokes_phase *q = &vc->phases[p];
        double g[6] = {0, 0, 0, 0, 0, 0}, qp = 0.0, fp = 0.0;
        if (q->has_pl) {
            const double tII = second_invariant6(t);
            for (int c = 0; c < 3; c++) g[c] = 0.5 * t[c] / tII;
            for (int c = 3; c < 6; c++) g[c] = 0.5 * (t[c] / tII);
            qp = -q->sinpsi; fp = -q->sinphi;
        }
        for (int c = 0; c < 6; c++) dQdt[c] = fma(r, g[c], dQdt[c]);
        *dQdP = fma(r, qp, *dQdP);
        *dFdP = fma(r, fp, *dFdP);
    }
}
/* fn_ratio(compute_density, rheology, ratio, args)  phases.jl:18-30: a phase with ratio == 1 returns early */
static inline double ratio_density(const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx, double T, double P)
{
    double rho = 0.0;
    for (int p = 0; p < vc->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        if (r == 1.0) return density(&vc->phases[p], T, P) * r;
        rho += (r == 0.0) ? 0.0 : density(&vc->phases[p], T, P) * r;
    }
    return rho;
}
#endif
