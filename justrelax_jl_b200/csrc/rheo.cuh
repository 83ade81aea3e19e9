// rheo.cuh — the flat per-phase rheology table on the device and the point-wise GeoParams subset evaluated from it.
// Not part of the C ABI.  Mirrors (operation for operation) what the reference evaluates through GeoParams.jl dispatch:
//   fn_ratio                        src/phases/phases.jl:5-30
//   get_shear_modulus/get_bulk_modulus   src/rheology/GeoParams.jl:1-15 (Inf for NaN / 0)
//   plastic_params_phase            src/rheology/StressUpdate.jl:146-188
//   compute_yieldfunction_phase     src/rheology/StressUpdate.jl:384-452
//   compute_plastic_gradients_phase src/rheology/StressUpdate.jl:463-550
//   compute_phase_viscosity         src/rheology/Viscosity.jl:599-619 (local args dt = Inf, τII_old = 0: :510-522)
//   compute_density / fn_ratio      src/rheology/BuoyancyForces.jl:74-95
// Phase ratios are stored [phase][node] (node contiguous): ratio of phase p at node q is ph[p * stride + q].
#pragma once
#include "common.cuh"

#define JR_MAX_PHASES 8

struct jr_phase_tab {
    int n, g_scalar, rho_const, _pad;
    double g[3];
    double fs;  // dt * free_surface factor of compute_V!/compute_Res! (2D), 0 when off
    double eta[JR_MAX_PHASES], G[JR_MAX_PHASES], Kb[JR_MAX_PHASES];
    // composite viscosity of a phase at the local args of compute_viscosity (dt = Inf: the elastic element contributes 1 / (G · Inf)) and its
    // reciprocal — per-phase constants of the LinearViscous subset, formed once on the host with the same IEEE divisions the kernels used to
    // repeat per node: eta_c = 1 / (1 / η + 1 / (G · Inf)), ieta_c = 1 / eta_c
    double eta_c[JR_MAX_PHASES], ieta_c[JR_MAX_PHASES];
    double C[JR_MAX_PHASES], sinphi[JR_MAX_PHASES], cosphi[JR_MAX_PHASES], sinpsi[JR_MAX_PHASES], eta_vp[JR_MAX_PHASES];
    double rho0[JR_MAX_PHASES], alpha[JR_MAX_PHASES], beta[JR_MAX_PHASES], T0[JR_MAX_PHASES], P0[JR_MAX_PHASES];
    int has_pl[JR_MAX_PHASES], rho_kind[JR_MAX_PHASES];
    int soft_kind[JR_MAX_PHASES], any_soft, _pad2;   // cohesion softening (2D solves)
    double soft[JR_MAX_PHASES][6];
};

int jr_make_phase_tab(const jr_vc_inputs *vc, jr_phase_tab *out);

__device__ __forceinline__ double jr_ratio_G(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q)
{
    double x = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        x += (r == 0.0) ? 0.0 : pt.G[p] * r;
    }
    return x;
}
__device__ __forceinline__ double jr_ratio_Kb(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q)
{
    double x = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        x += (r == 0.0) ? 0.0 : pt.Kb[p] * r;
    }
    return x;
}
// fn_ratio(get_thermal_expansion, rheology, ratio)  src/rheology/GeoParams.jl:17 (α of the density law; 0 for ConstantDensity)
__device__ __forceinline__ double jr_ratio_alpha(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q)
{
    double x = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        x += (r == 0.0) ? 0.0 : (pt.rho_kind[p] == 0 ? 0.0 : pt.alpha[p]) * r;
    }
    return x;
}
__device__ __forceinline__ void jr_plastic_params(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, bool &is_pl,
                                                  double &eta_reg)
{
    is_pl = false;
    eta_reg = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        const bool pl = (r != 0.0) && pt.has_pl[p];
        if (pl) is_pl = true;
        eta_reg += (pl ? pt.eta_vp[p] : 0.0) * r;
    }
}
__device__ __forceinline__ double jr_yield_F(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, double P, double tII)
{
    double acc = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        double v = 0.0;
        if (r != 0.0) {
            const double Fp = pt.has_pl[p] ? (tII - pt.cosphi[p] * pt.C[p] - pt.sinphi[p] * (P - 0.0)) - 2 * pt.eta_vp[p] * (0.0 * 0.5) : tII;
            v = r * Fp;
        }
        acc = p == 0 ? v : acc + v;
    }
    return acc;
}
// soften_cohesion  StressUpdate.jl:305-332 → GeoParams LinearSoftening / NonLinearSoftening
__device__ __forceinline__ double jr_soften_C(const jr_phase_tab &pt, int p, double EII)
{
    if (pt.soft_kind[p] == 1) {
        if (EII >= pt.soft[p][1]) return pt.soft[p][3];
        if (EII <= pt.soft[p][0]) return pt.soft[p][2];
        return EII * pt.soft[p][4] + pt.soft[p][5];
    }
    if (pt.soft_kind[p] == 2) return pt.soft[p][0] - 0.5 * pt.soft[p][1] * erfc(-(EII - pt.soft[p][2]) / pt.soft[p][3]);
    return pt.C[p];
}
__device__ __forceinline__ double jr_yield_F_soft(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, double P, double tII, double EII)
{
    double acc = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        double v = 0.0;
        if (r != 0.0) {
            const double Fp = pt.has_pl[p] ? (tII - pt.cosphi[p] * jr_soften_C(pt, p, EII) - pt.sinphi[p] * (P - 0.0)) - 2 * pt.eta_vp[p] * (0.0 * 0.5) : tII;
            v = r * Fp;
        }
        acc = p == 0 ? v : acc + v;
    }
    return acc;
}
// NC = 3 (2D: xx, yy, xy) or 6 (3D: xx, yy, zz, yz, xz, xy); t = trial stress; shear slots are halved (StressUpdate.jl:467-472)
template <int NC>
__device__ __forceinline__ double jr_second_invariant(const double *t)
{
    if (NC == 3) return sqrt(0.5 * (t[0] * t[0] + t[1] * t[1]) + t[2] * t[2]);
    return sqrt(0.5 * (t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) + t[3] * t[3] + t[4] * t[4] + t[5] * t[5]);
}
template <int NC>
__device__ __forceinline__ void jr_plastic_grads(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, const double *t,
                                                 double *dQdt, double &dQdP, double &dFdP)
{
    constexpr int NN = NC == 3 ? 2 : 3;
#pragma unroll
    for (int c = 0; c < NC; c++) dQdt[c] = 0.0;
    dQdP = 0.0;
    dFdP = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        if (r == 0.0) continue;
        double g[NC], qp = 0.0, fp = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c++) g[c] = 0.0;
        if (pt.has_pl[p]) {
            const double tII = jr_second_invariant<NC>(t);
#pragma unroll
            for (int c = 0; c < NN; c++) g[c] = 0.5 * t[c] / tII;
#pragma unroll
            for (int c = NN; c < NC; c++) g[c] = 0.5 * (t[c] / tII);
            qp = -pt.sinpsi[p];
            fp = -pt.sinphi[p];
        }
#pragma unroll
        for (int c = 0; c < NC; c++) dQdt[c] = fma(r, g[c], dQdt[c]);
        dQdP = fma(r, qp, dQdP);
        dFdP = fma(r, fp, dFdP);
    }
}
// the same gradients in two parts, so that the divisions of ∂Q/∂τ are only paid where a node yields (and only for the components it uses):
// jr_plastic_dP = the pressure derivatives (needed everywhere: the volumetric term of λ), jr_plastic_dQ<NC, C> = component C of ∂Q/∂τ.
// Operation for operation the accumulations of jr_plastic_grads (same order over the phases, same fma), hence the same bits.
__device__ __forceinline__ void jr_plastic_dP(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, double &dQdP, double &dFdP)
{
    dQdP = 0.0;
    dFdP = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        if (r == 0.0) continue;
        dQdP = fma(r, pt.has_pl[p] ? -pt.sinpsi[p] : 0.0, dQdP);
        dFdP = fma(r, pt.has_pl[p] ? -pt.sinphi[p] : 0.0, dFdP);
    }
}
// jr_ratio_G + jr_ratio_Kb + jr_plastic_params + jr_plastic_dP in ONE sweep over the phases (one load of each ratio; every accumulation keeps
// its own order, so each result has the bits of the separate function)
__device__ __forceinline__ void jr_mix_sweep(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, double &G, double &Kb,
                                             bool &is_pl, double &eta_reg, double &dQdP, double &dFdP)
{
    G = 0.0; Kb = 0.0; eta_reg = 0.0; dQdP = 0.0; dFdP = 0.0;
    is_pl = false;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        const bool nz = r != 0.0, pl = nz && pt.has_pl[p];
        G += nz ? pt.G[p] * r : 0.0;
        Kb += nz ? pt.Kb[p] * r : 0.0;
        if (pl) is_pl = true;
        eta_reg += (pl ? pt.eta_vp[p] : 0.0) * r;
        if (nz) {
            dQdP = fma(r, pt.has_pl[p] ? -pt.sinpsi[p] : 0.0, dQdP);
            dFdP = fma(r, pt.has_pl[p] ? -pt.sinphi[p] : 0.0, dFdP);
        }
    }
}
template <int NC, int C>
__device__ __forceinline__ double jr_plastic_dQ(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, const double *t, double tII)
{
    constexpr int NN = NC == 3 ? 2 : 3;
    double acc = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        if (r == 0.0) continue;
        const double g = pt.has_pl[p] ? (C < NN ? 0.5 * t[C] / tII : 0.5 * (t[C] / tII)) : 0.0;
        acc = fma(r, g, acc);
    }
    return acc;
}
__device__ __forceinline__ double jr_density(const jr_phase_tab &pt, int p, double T, double P)
{
    if (pt.rho_kind[p] == 1) return pt.rho0[p] * (1.0 - pt.alpha[p] * (T - pt.T0[p]) + pt.beta[p] * (P - pt.P0[p]));
    if (pt.rho_kind[p] == 2) return pt.rho0[p] * (1.0 - pt.alpha[p] * (T - pt.T0[p]));
    return pt.rho0[p];
}
// fn_ratio(compute_density, rheology, ratio, args)
__device__ __forceinline__ double jr_ratio_density(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, double T, double P)
{
    double rho = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        if (r == 1.0) return jr_density(pt, p, T, P) * r;
        rho += (r == 0.0) ? 0.0 : jr_density(pt, p, T, P) * r;
    }
    return rho;
}
__device__ __forceinline__ double jr_phase_viscosity(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q)
{
    for (int p = 0; p < pt.n; p++)
        if (ph[(size_t)p * stride + q] > 0.999) return pt.eta_c[p];
    double e = 0.0;
    for (int p = 0; p < pt.n; p++) {
        const double r = ph[(size_t)p * stride + q];
        if (r != 0.0) e += pt.ieta_c[p] * r;
    }
    return jr_inv(e);
}
__device__ __forceinline__ double jr_clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
