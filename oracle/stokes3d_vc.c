/*
 * stokes3d_vc.c — CPU ORACLE (test infrastructure, NOT product code).
 * Restatement of the 3D multiphase visco-elasto-plastic Stokes PT loop, variant 3D-VC:
 *   src/stokes/Stokes3D.jl:447-668 (driver)
 *   src/stokes/PressureKernels.jl:64-102,186-195 (compute_P_kernel!, phase-ratio moduli)
 *   src/stokes/VelocityKernels.jl:3-6,59-104 (∇V, strain rate over `ni`: quirk Q20),:182-242 (compute_V!)
 *   src/stokes/StressKernels.jl:606-669 (clamped averages), :672-989 (update_stresses_center_vertex_ps! 3D)
 *   src/rheology/StressUpdate.jl:146-188,248-301,384-550 ; src/rheology/Viscosity.jl:454-522,599-619
 *   src/rheology/BuoyancyForces.jl:38-95,153-167 ; src/Utils.jl:409-461 (compute_maxloc!, every iteration: Q12)
 *   exit kernels: src/stress_rotation/stress_rotation_particles.jl:32-51 (compute_vorticity!),
 *   src/Interpolations.jl:313-323 (shear2center!), src/stokes/StressKernels.jl:394-438 (accumulate_tensor!/vol!)
 * The stress kernel is racy in the reference (quirk Q7); the oracle declares the Jacobi schedule canonical: every
 * value read from an array another thread of the same launch writes (τxx, τyy, τzz at neighbouring cells, the other
 * two edge shear arrays) is the value BEFORE the launch.
 * Parity status: no reference test pins numbers for 3D-VC (test_shearband3D_MPI.jl only runs it); the rheology helpers
 * are the ones pinned through the 2D shear-band goldens, the velocity/pressure kernels through the SolVi3D goldens,
 * and tests/test_oracle_stokes3d_vc.py cross-checks the converged 3D-VC solution against the 3D-VA oracle.
 */
#include "jr_oracle.h"
#include "mini.h"
#include "vc_common.h"
#include <stdlib.h>
#include <string.h>

#define F(name) (s->f[ORC_F_##name])
#define A3(p, n1, n2, i, j, k) ((p)[IX3(n1, n2, i, j, k)])

/* compute_ρg! 3D  BuoyancyForces.jl:38-60: fn_ratio(compute_density, …) .* g ; args.T (ni.+2) sampled at I+1 (Q17) */
void orc_rhog3d(const orc_fields *s, const orc_vc_inputs *vc)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const size_t nc = (size_t)nx * ny * nz;
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                const size_t c = IX3(nx, ny, i, j, k);
                const double T = F(T) ? A3(F(T), nx + 2, ny + 2, i + 1, j + 1, k + 1) : 0.0, P = F(Pargs) ? F(Pargs)[c] : 0.0;
                const double rho = ratio_density(vc, vc->ph_center, nc, c, T, P);
                if (vc->g_scalar) F(rhogz)[c] = rho * vc->g[2];
                else { F(rhogx)[c] = rho * vc->g[0]; F(rhogy)[c] = rho * vc->g[1]; F(rhogz)[c] = rho * vc->g[2]; }
            }
}

/* compute_viscosity_kernel! 3D (τII form, centres only: Viscosity.jl:306-308)  Viscosity.jl:454-504.
 * LinearViscous does not depend on the invariant, so AII (and its eps() guard) does not enter the subset. */
void orc_viscosity3d(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, double nu)
{
    const size_t nc = (size_t)s->n[0] * s->n[1] * s->n[2];
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < nc; c++) {
        const double ei = phase_viscosity(vc, vc->ph_center, nc, c);
        F(eta)[c] = clampd((1 - nu) * F(eta)[c] + nu * ei, o->visc_cutoff_lo, o->visc_cutoff_hi);
    }
}

/* _compute_P!  PressureKernels.jl:186-195 */
static inline void P_point(double *RP, double *P, double P0, double divV, double Q, double eta, double K, double G, double dt, double r, double th)
{
    const double _Kdt = orc_inv(K * dt), _Gdt = orc_inv(G * dt), _dt = orc_inv(dt), Pc = *P;
    *RP = fma(-(Pc - P0), _Kdt, (-divV + (Q * _dt)));
    const double psi = orc_inv(orc_inv(eta) + _Gdt) * r / th;
    *P = ((fma(P0, _Kdt, (-divV + (Q * _dt)))) * psi + Pc) / (1 + _Kdt * psi);
}

static inline double stress_inc(double t, double to, double eta, double e, double _Gdt, double dtr)
{
    return dtr * fma(2.0 * eta, e, fma(-(t - to) * eta, _Gdt, -t));
}

typedef struct { double *theta, *lam, *lamv_yz, *lamv_xz, *lamv_xy; } vc3_scratch;

/* one edge family of update_stresses_center_vertex_ps! (StressKernels.jl:716-778 yz, :781-849 xz, :852-921 xy) */
typedef struct {
    double eta, P, e[6], t[6], to[6];
} edge_in;

static inline void edge_update(const orc_stokes_opts *o, const orc_vc_inputs *vc, const double *ph, size_t stride, size_t idx, const edge_in *in, int slot,
                               double *lamv, double *tau, double *epl)
{
    const double dt = o->dt, th = o->theta_dtau, rel = o->lambda_relaxation;
    int is_pl; double eta_reg;
    plastic_params(vc, ph, stride, idx, &is_pl, &eta_reg);
    const double _Gdt = orc_inv(ratio_G(vc, ph, stride, idx) * dt), Kv = ratio_Kb(vc, ph, stride, idx);
    const double etav = in->eta, dtr = orc_inv(th + etav * _Gdt + 1.0);
    double d[6], trial[6], tt[6];
    for (int c = 0; c < 6; c++) d[c] = stress_inc(in->t[c], in->to[c], etav, in->e[c], _Gdt, dtr);
    for (int c = 0; c < 6; c++) { tt[c] = in->t[c] + d[c]; trial[c] = in->t[c] + d[c]; }
    const double tII = second_invariant6(tt);
    double dQ[6], dQdP, dFdP;
    plastic_grads6(vc, ph, stride, idx, trial, dQ, &dQdP, &dFdP);
    const double volume = isinf(Kv) ? 0.0 : Kv * dt * dFdP * dQdP;
    const double Fv = yield_F(vc, ph, stride, idx, in->P, tII);
    if (is_pl && tII != 0.0 && Fv > 0) {
        *lamv = (1.0 - rel) * *lamv + rel * (fmax(Fv, 0.0) / (etav * dtr + eta_reg + volume));
        const double e_pl = *lamv * dQ[slot];
        *tau += fma(-2.0, etav * e_pl * dtr, d[slot]);   /* @muladd dτ − 2.0 * ηv * ε_pl * dτ_rv */
        *epl = e_pl;
    } else {
        *tau += d[slot];
        *epl = 0.0;
    }
}

/* update_stresses_center_vertex_ps! 3D  StressKernels.jl:672-989, Jacobi schedule */
static void stress_vep3(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, vc3_scratch *w)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const size_t nc = (size_t)nx * ny * nz, nyz = (size_t)nx * (ny + 1) * (nz + 1), nxz = (size_t)(nx + 1) * ny * (nz + 1), nxy = (size_t)(nx + 1) * (ny + 1) * nz;
    const double dt = o->dt, th = o->theta_dtau, rel = o->lambda_relaxation;
    /* snapshots of everything the launch both reads at foreign indices and writes */
    double *txx0 = (double *)malloc(nc * 8), *tyy0 = (double *)malloc(nc * 8), *tzz0 = (double *)malloc(nc * 8);
    double *tyz0 = (double *)malloc(nyz * 8), *txz0 = (double *)malloc(nxz * 8), *txy0 = (double *)malloc(nxy * 8);
    memcpy(txx0, F(txx), nc * 8); memcpy(tyy0, F(tyy), nc * 8); memcpy(tzz0, F(tzz), nc * 8);
    memcpy(tyz0, F(tyz), nyz * 8); memcpy(txz0, F(txz), nxz * 8); memcpy(txy0, F(txy), nxy * 8);
    const double *theta = w->theta;
    /* C(A): cell-centred (nx,ny,nz); YZ/XZ/XY: edge arrays */
#define C(A, i, j, k) A3(A, nx, ny, i, j, k)
#define YZ(A, i, j, k) A3(A, nx, ny + 1, i, j, k)
#define XZ(A, i, j, k) A3(A, nx + 1, ny, i, j, k)
#define XY(A, i, j, k) A3(A, nx + 1, ny + 1, i, j, k)
    /* clamped averages  StressKernels.jl:620-669 (argument order = summation order) */
#define AV_YZ(A) (0.25 * (C(A, ic, j0, k0) + C(A, ic, jc, k0) + C(A, ic, j0, kc) + C(A, ic, jc, kc)))
#define AV_XZ(A) (0.25 * (C(A, i0, jc, k0) + C(A, ic, jc, k0) + C(A, i0, jc, kc) + C(A, ic, jc, kc)))
#define AV_XY(A) (0.25 * (C(A, i0, j0, kc) + C(A, ic, j0, kc) + C(A, i0, jc, kc) + C(A, ic, jc, kc)))
#define HARM_YZ(A) (4 / (1 / C(A, ic, j0, k0) + 1 / C(A, ic, jc, k0) + 1 / C(A, ic, j0, kc) + 1 / C(A, ic, jc, kc)))
#define HARM_XZ(A) (4 / (1 / C(A, i0, jc, k0) + 1 / C(A, ic, jc, k0) + 1 / C(A, i0, jc, kc) + 1 / C(A, ic, jc, kc)))
#define HARM_XY(A) (4 / (1 / C(A, i0, j0, kc) + 1 / C(A, ic, j0, kc) + 1 / C(A, i0, jc, kc) + 1 / C(A, ic, jc, kc)))
#define AV_YZ_Z(A) (0.25 * (XY(A, ic, jc, k0) + XY(A, i1, jc, k0) + XY(A, ic, jc, kc) + XY(A, i1, jc, kc)))   /* on xy arrays */
#define AV_YZ_Y(A) (0.25 * (XZ(A, ic, j0, kc) + XZ(A, i1, j0, kc) + XZ(A, ic, jc, kc) + XZ(A, i1, jc, kc)))   /* on xz arrays */
#define AV_XZ_Z(A) (0.25 * (XY(A, ic, jc, k0) + XY(A, ic, j1, k0) + XY(A, ic, jc, kc) + XY(A, ic, j1, kc)))   /* on xy arrays */
#define AV_XZ_X(A) (0.25 * (YZ(A, i0, jc, kc) + YZ(A, ic, jc, kc) + YZ(A, ic, j1, kc) + YZ(A, i0, j1, kc)))   /* on yz arrays */
#define AV_XY_Y(A) (0.25 * (XZ(A, ic, j0, kc) + XZ(A, ic, jc, kc) + XZ(A, ic, j0, k1) + XZ(A, ic, jc, k1)))   /* on xz arrays */
#define AV_XY_X(A) (0.25 * (YZ(A, i0, jc, kc) + YZ(A, ic, jc, kc) + YZ(A, i0, jc, k1) + YZ(A, ic, jc, k1)))   /* on yz arrays */
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz + 1; k++)
        for (int j = 1; j <= ny + 1; j++)
            for (int i = 1; i <= nx + 1; i++) {
                const int i0 = orc_clamp(i - 1, 1, nx), ic = orc_clamp(i, 1, nx), i1 = orc_clamp(i + 1, 1, nx);
                const int j0 = orc_clamp(j - 1, 1, ny), jc = orc_clamp(j, 1, ny), j1 = orc_clamp(j + 1, 1, ny);
                const int k0 = orc_clamp(k - 1, 1, nz), kc = orc_clamp(k, 1, nz), k1 = orc_clamp(k + 1, 1, nz);
                /* ---- yz edge ---- */
                if (i <= nx && j <= ny + 1 && k <= nz + 1) {
                    const size_t v = IX3(nx, ny + 1, i, j, k);
                    edge_in in;
                    in.eta = HARM_YZ(F(eta)); in.P = AV_YZ(theta);
                    in.e[0] = AV_YZ(F(exx)); in.e[1] = AV_YZ(F(eyy)); in.e[2] = AV_YZ(F(ezz));
                    in.e[3] = F(eyz)[v]; in.e[4] = AV_YZ_Y(F(exz)); in.e[5] = AV_YZ_Z(F(exy));
                    in.t[0] = AV_YZ(txx0); in.t[1] = AV_YZ(tyy0); in.t[2] = AV_YZ(tzz0);
                    in.t[3] = tyz0[v]; in.t[4] = AV_YZ_Y(txz0); in.t[5] = AV_YZ_Z(txy0);
                    in.to[0] = AV_YZ(F(txx_o)); in.to[1] = AV_YZ(F(tyy_o)); in.to[2] = AV_YZ(F(tzz_o));
                    in.to[3] = F(tyz_o)[v]; in.to[4] = AV_YZ_Y(F(txz_o)); in.to[5] = AV_YZ_Z(F(txy_o));
                    edge_update(o, vc, vc->ph_yz, nyz, v, &in, 3, &w->lamv_yz[v], &F(tyz)[v], &F(pyz)[v]);
                }
                /* ---- xz edge ---- */
                if (i <= nx + 1 && j <= ny && k <= nz + 1) {
                    const size_t v = IX3(nx + 1, ny, i, j, k);
                    edge_in in;
                    in.eta = HARM_XZ(F(eta)); in.P = AV_XZ(theta);
                    in.e[0] = AV_XZ(F(exx)); in.e[1] = AV_XZ(F(eyy)); in.e[2] = AV_XZ(F(ezz));
                    in.e[3] = AV_XZ_X(F(eyz)); in.e[4] = F(exz)[v]; in.e[5] = AV_XZ_Z(F(exy));
                    in.t[0] = AV_XZ(txx0); in.t[1] = AV_XZ(tyy0); in.t[2] = AV_XZ(tzz0);
                    in.t[3] = AV_XZ_X(tyz0); in.t[4] = txz0[v]; in.t[5] = AV_XZ_Z(txy0);
                    in.to[0] = AV_XZ(F(txx_o)); in.to[1] = AV_XZ(F(tyy_o)); in.to[2] = AV_XZ(F(tzz_o));
                    in.to[3] = AV_XZ_X(F(tyz_o)); in.to[4] = F(txz_o)[v]; in.to[5] = AV_XZ_Z(F(txy_o));
                    edge_update(o, vc, vc->ph_xz, nxz, v, &in, 4, &w->lamv_xz[v], &F(txz)[v], &F(pxz)[v]);
                }
                /* ---- xy edge ---- */
                if (i <= nx + 1 && j <= ny + 1 && k <= nz) {
                    const size_t v = IX3(nx + 1, ny + 1, i, j, k);
                    edge_in in;
                    in.eta = HARM_XY(F(eta)); in.P = AV_XY(theta);
                    in.e[0] = AV_XY(F(exx)); in.e[1] = AV_XY(F(eyy)); in.e[2] = AV_XY(F(ezz));
                    in.e[3] = AV_XY_X(F(eyz)); in.e[4] = AV_XY_Y(F(exz)); in.e[5] = F(exy)[v];
                    in.t[0] = AV_XY(txx0); in.t[1] = AV_XY(tyy0); in.t[2] = AV_XY(tzz0);
                    in.t[3] = AV_XY_X(tyz0); in.t[4] = AV_XY_Y(txz0); in.t[5] = txy0[v];
                    in.to[0] = AV_XY(F(txx_o)); in.to[1] = AV_XY(F(tyy_o)); in.to[2] = AV_XY(F(tzz_o));
                    in.to[3] = AV_XY_X(F(tyz_o)); in.to[4] = AV_XY_Y(F(txz_o)); in.to[5] = F(txy_o)[v];
                    edge_update(o, vc, vc->ph_xy, nxy, v, &in, 5, &w->lamv_xy[v], &F(txy)[v], &F(pxy)[v]);
                }
                /* ---- centre ----  StressKernels.jl:923-986 (no @muladd here: plain products and sums) */
                if (i <= nx && j <= ny && k <= nz) {
                    const size_t c = IX3(nx, ny, i, j, k);
                    const double _Gdt = orc_inv(ratio_G(vc, vc->ph_center, nc, c) * dt);
                    int is_pl; double eta_reg;
                    plastic_params(vc, vc->ph_center, nc, c, &is_pl, &eta_reg);
                    const double K = ratio_Kb(vc, vc->ph_center, nc, c), eta = F(eta)[c];
                    const double dtr = orc_inv(th + eta * _Gdt + 1.0);
                    const arr eyz = {F(eyz), nx, ny + 1, nz + 1}, exz = {F(exz), nx + 1, ny, nz + 1}, exy = {F(exy), nx + 1, ny + 1, nz};
                    /* cache_tensors  StressUpdate.jl:248-301: _av_yz/_av_xz/_av_xy (MiniKernels.jl) */
                    const double eij[6] = {F(exx)[c], F(eyy)[c], F(ezz)[c], av_yz3(eyz, i, j, k), av_xz3(exz, i, j, k), av_xy3(exy, i, j, k)};
                    double tij[6] = {txx0[c], tyy0[c], tzz0[c], F(tyz_c)[c], F(txz_c)[c], F(txy_c)[c]};
                    const double tijo[6] = {F(txx_o)[c], F(tyy_o)[c], F(tzz_o)[c], F(tyz_o_c)[c], F(txz_o_c)[c], F(txy_o_c)[c]};
                    double d[6], tt[6], trial[6];
                    for (int q = 0; q < 6; q++) d[q] = (-(tij[q] - tijo[q]) * eta * _Gdt - tij[q] + 2.0 * eta * eij[q]) * dtr;
                    for (int q = 0; q < 6; q++) { tt[q] = d[q] + tij[q]; trial[q] = tij[q] + d[q]; }
                    double tII = second_invariant6(tt);
                    double dQ[6], dQdP, dFdP;
                    const double Pr = theta[c];
                    plastic_grads6(vc, vc->ph_center, nc, c, trial, dQ, &dQdP, &dFdP);
                    const double volume = isinf(K) ? 0.0 : K * dt * dFdP * dQdP;
                    const double Fc = yield_F(vc, vc->ph_center, nc, c, Pr, tII);
                    if (is_pl && tII != 0.0 && Fc > 0) {
                        w->lam[c] = (1.0 - rel) * w->lam[c] + rel * (fmax(Fc, 0.0) / (eta * dtr + eta_reg + volume));
                        double epl[6];
                        for (int q = 0; q < 6; q++) {
                            epl[q] = w->lam[c] * dQ[q];
                            d[q] = d[q] - 2.0 * eta * epl[q] * dtr;
                            tij[q] = d[q] + tij[q];
                        }
                        F(e_vol_pl)[c] = -w->lam[c] * dQdP;
                        F(txx)[c] = tij[0]; F(tyy)[c] = tij[1]; F(tzz)[c] = tij[2]; F(tyz_c)[c] = tij[3]; F(txz_c)[c] = tij[4]; F(txy_c)[c] = tij[5];
                        F(pxx)[c] = epl[0]; F(pyy)[c] = epl[1]; F(pzz)[c] = epl[2];
                        tII = second_invariant6(tij);
                    } else {
                        F(e_vol_pl)[c] = 0.0;
                        F(txx)[c] = d[0] + tij[0]; F(tyy)[c] = d[1] + tij[1]; F(tzz)[c] = d[2] + tij[2];
                        F(tyz_c)[c] = d[3] + tij[3]; F(txz_c)[c] = d[4] + tij[4]; F(txy_c)[c] = d[5] + tij[5];
                        F(pxx)[c] = 0.0; F(pyy)[c] = 0.0; F(pzz)[c] = 0.0;
                    }
                    F(tII)[c] = tII;
                    F(eta_vep)[c] = tII * 0.5 * orc_inv(second_invariant6(eij));
                    F(P)[c] = Pr - (isinf(K) ? 0.0 : K * dt * w->lam[c] * dQdP);
                }
            }
    free(txx0); free(tyy0); free(tzz0); free(tyz0); free(txz0); free(txy0);
}

static void pre_VC3(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, vc3_scratch *w)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const size_t nc = (size_t)nx * ny * nz, nyz = (size_t)nx * (ny + 1) * (nz + 1), nxz = (size_t)(nx + 1) * ny * (nz + 1), nxy = (size_t)(nx + 1) * (ny + 1) * nz;
    memcpy(F(P0), F(P), nc * 8);                                             /* @copy stokes.P0 stokes.P   :493 */
    w->theta = (double *)malloc(nc * 8); memcpy(w->theta, F(P), nc * 8);     /* θ = deepcopy(stokes.P)     :494 */
    w->lam = (double *)calloc(nc, 8);                                        /* λ, λv_* = 0                :495-498 */
    w->lamv_yz = (double *)calloc(nyz, 8); w->lamv_xz = (double *)calloc(nxz, 8); w->lamv_xy = (double *)calloc(nxy, 8);
    memcpy(F(etatau), F(eta), nc * 8);                                       /* ητ = deepcopy(η)           :502 */
    orc_rhog3d(s, vc);                                                       /* compute_ρg!                :505 */
    /* compute_viscosity! (εII form, ν = 1)  :506 → Viscosity.jl:57-66: for the LinearViscous subset independent of εII */
    orc_viscosity3d(s, o, vc, 1.0);
}

void orc_vc3_P(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, double *theta)
{
    const size_t nc = (size_t)s->n[0] * s->n[1] * s->n[2];
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < nc; c++) { /* compute_P!(θ, P0, RP, ∇V, Q, ητ, rheology, phase_ratios, …)  :518-531 */
        const double K = ratio_Kb(vc, vc->ph_center, nc, c), G = ratio_G(vc, vc->ph_center, nc, c);
        if (F(dTargs)) /* args.ΔT given: compute_P_kernel!(…, ΔT, ::Nothing)  PressureKernels.jl:128-149 */
            P_point_dT(&F(RP)[c], &theta[c], F(P0)[c], F(divV)[c], F(Q)[c], F(dTargs)[o->dT_ghosted ? IX3(s->n[0] + 2, s->n[1] + 2, c % s->n[0] + 1, (c / s->n[0]) % s->n[1] + 1, c / ((size_t)s->n[0] * s->n[1]) + 1) : c], ratio_alpha(vc, vc->ph_center, nc, c), F(etatau)[c], K, G,
                       o->dt, o->r, o->theta_dtau);
        else
            P_point(&F(RP)[c], &theta[c], F(P0)[c], F(divV)[c], F(Q)[c], F(etatau)[c], K, G, o->dt, o->r, o->theta_dtau);
    }
}

/* the loop body in three pieces, so the multi-rank emulation of the tests can exchange halos where the reference does */
static void iter_VC3_a(const orc_fields *s)
{
    orc_maxloc3(F(etatau), F(eta), s->n[0], s->n[1], s->n[2], 1, 1, 1);      /* compute_maxloc!(ητ, η)  :514 ; update_halo!(ητ) :515 */
}
static void iter_VC3_b(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, vc3_scratch *w)
{
    orc_compute_divV3(s, o->_di);                                            /* :517 */
    orc_vc3_P(s, o, vc, w->theta);                                           /* :518-531 */
    orc_compute_strain_rate3(s, o->_di, 0);                                  /* over ni only (Q20)  :533-535 */
    if (!density_is_constant(vc)) orc_rhog3d(s, vc);                         /* update_ρg!  :538 */
    orc_viscosity3d(s, o, vc, o->viscosity_relaxation);                      /* update_viscosity_τII! BEFORE the stress kernel (Q13)  :541-548 */
    stress_vep3(s, o, vc, w);                                                /* :551-577 ; update_halo!(τyz, τxz, τxy) :578-580 */
}
static void iter_VC3_c(const orc_fields *s, const orc_stokes_opts *o)
{
    orc_compute_V3(s, o->eta_dtau, o->_di);                                  /* :583-592 */
    orc_velocity2displacement(s, o->dt);
    orc_flow_bcs3(s, o, 0);                                                  /* ; update_halo!(V…) :596 */
}

static void norms_VC3(const orc_fields *s, const orc_stokes_opts *o, double e[4])
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const double ng = (double)(o->n_g[0] - 1) * (o->n_g[1] - 1) * (o->n_g[2] - 1);   /* quirk Q4  :604-611 */
    e[0] = sqrt(orc_sumsq_interior(F(Rx), nx - 1, ny, nz, 1)) / ng;
    e[1] = sqrt(orc_sumsq_interior(F(Ry), nx, ny - 1, nz, 1)) / ng;
    e[2] = sqrt(orc_sumsq_interior(F(Rz), nx, ny, nz - 1, 1)) / ng;
    e[3] = sqrt(orc_sumsq_interior(F(RP), nx, ny, nz, 0)) / ((double)nx * ny * nz);    /* local length(RP) */
}

static void post_VC3(const orc_fields *s, const orc_stokes_opts *o, vc3_scratch *w, int finish)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    if (F(lam)) memcpy(F(lam), w->lam, (size_t)nx * ny * nz * 8);            /* expose the solver-local λ for parity checks */
    if (finish) {
        const double _dx = o->_di[0], _dy = o->_di[1], _dz = o->_di[2];
        const arr Vx = {F(Vx), nx + 1, ny + 2, nz + 2}, Vy = {F(Vy), nx + 2, ny + 1, nz + 2}, Vz = {F(Vz), nx + 2, ny + 2, nz + 1};
        /* compute_vorticity!(ωyz, ωxz, ωxy, V…, _di) over ni.+1  :641-643 ; stress_rotation_particles.jl:32-51 (plain _d_*a at I) */
        for (int k = 1; k <= nz + 1; k++)
            for (int j = 1; j <= ny + 1; j++)
                for (int i = 1; i <= nx + 1; i++) {
                    if (F(wyz) && i <= nx && j <= ny + 1 && k <= nz + 1)
                        A3(F(wyz), nx, ny + 1, i, j, k) = 0.5 * (d_ya3(Vz, _dy, i, j, k) - d_za3(Vy, _dz, i, j, k));
                    if (F(wxz) && i <= nx + 1 && j <= ny && k <= nz + 1)
                        A3(F(wxz), nx + 1, ny, i, j, k) = 0.5 * (d_za3(Vx, _dz, i, j, k) - d_xa3(Vz, _dx, i, j, k));
                    if (F(wxy) && i <= nx + 1 && j <= ny + 1 && k <= nz)
                        A3(F(wxy), nx + 1, ny + 1, i, j, k) = 0.5 * (d_xa3(Vy, _dx, i, j, k) - d_ya3(Vx, _dy, i, j, k));
                }
        /* shear2center!(ε), (ε_pl), (Δε)  :646-648 ; Interpolations.jl:313-323 */
        double *cen[3][3] = {{F(eyz_c), F(exz_c), F(exy_c)}, {F(pyz_c), F(pxz_c), F(pxy_c)}, {F(dyz_c), F(dxz_c), F(dxy_c)}};
        double *she[3][3] = {{F(eyz), F(exz), F(exy)}, {F(pyz), F(pxz), F(pxy)}, {F(dyz), F(dxz), F(dxy)}};
        for (int q = 0; q < 3; q++) {
            if (!cen[q][0] || !she[q][0] || !cen[q][1] || !she[q][1] || !cen[q][2] || !she[q][2]) continue;
            const double *yz = she[q][0], *xz = she[q][1], *xy = she[q][2];
            for (int k = 1; k <= nz; k++)
                for (int j = 1; j <= ny; j++)
                    for (int i = 1; i <= nx; i++) {
                        const size_t c = IX3(nx, ny, i, j, k);
                        cen[q][0][c] = 0.25 * (YZ(yz, i, j, k) + YZ(yz, i, j + 1, k) + YZ(yz, i, j, k + 1) + YZ(yz, i, j + 1, k + 1));
                        cen[q][1][c] = 0.25 * (XZ(xz, i, j, k) + XZ(xz, i + 1, j, k) + XZ(xz, i, j, k + 1) + XZ(xz, i + 1, j, k + 1));
                        cen[q][2][c] = 0.25 * (XY(xy, i, j, k) + XY(xy, i + 1, j, k) + XY(xy, i, j + 1, k) + XY(xy, i + 1, j + 1, k));
                    }
        }
        /* accumulate_tensor!(EII_pl, ε_pl, dt), accumulate_vol!  :651-652 ; second_invariant_staggered (GeoParams): the shear
         * terms are the means of the squares of the four gathered edge values (_gather_yz/_xz/_xy, MiniKernels.jl:196-204) */
        for (int k = 1; k <= nz; k++)
            for (int j = 1; j <= ny; j++)
                for (int i = 1; i <= nx; i++) {
                    const size_t c = IX3(nx, ny, i, j, k);
                    const double xx = F(pxx)[c], yy = F(pyy)[c], zz = F(pzz)[c];
                    const double a1 = YZ(F(pyz), i, j, k), a2 = YZ(F(pyz), i, j + 1, k), a3 = YZ(F(pyz), i, j, k + 1), a4 = YZ(F(pyz), i, j + 1, k + 1);
                    const double b1 = XZ(F(pxz), i, j, k), b2 = XZ(F(pxz), i + 1, j, k), b3 = XZ(F(pxz), i, j, k + 1), b4 = XZ(F(pxz), i + 1, j, k + 1);
                    const double c1 = XY(F(pxy), i, j, k), c2 = XY(F(pxy), i + 1, j, k), c3 = XY(F(pxy), i, j + 1, k), c4 = XY(F(pxy), i + 1, j + 1, k);
                    const double yz2 = (((a1 * a1 + a2 * a2) + a3 * a3) + a4 * a4) / 4, xz2 = (((b1 * b1 + b2 * b2) + b3 * b3) + b4 * b4) / 4,
                                 xy2 = (((c1 * c1 + c2 * c2) + c3 * c3) + c4 * c4) / 4;
                    F(EII_pl)[c] += sqrt(0.5 * (xx * xx + yy * yy + zz * zz) + yz2 + xz2 + xy2) * o->dt;
                    F(EVol_pl)[c] += o->dt * F(e_vol_pl)[c];
                }
        orc_multi_copy_tau3(s);                                              /* :654-655 */
    }
    free(w->theta); free(w->lam); free(w->lamv_yz); free(w->lamv_xz); free(w->lamv_xy);
}

int orc_iterate3d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, int64_t niter, int finish)
{
    vc3_scratch w;
    pre_VC3(s, o, vc, &w);
    for (int64_t it = 0; it < niter; it++) { iter_VC3_a(s); iter_VC3_b(s, o, vc, &w); iter_VC3_c(s, o); }
    post_VC3(s, o, &w, finish);
    return 0;
}

/* stepwise entry points for the multi-rank emulation (the caller owns the scratch through an opaque handle) */
void *orc_vc3_begin(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc)
{
    vc3_scratch *w = (vc3_scratch *)malloc(sizeof(vc3_scratch));
    pre_VC3(s, o, vc, w);
    return w;
}
void orc_vc3_step(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, void *h, int piece)
{
    if (piece == 0) iter_VC3_a(s);
    else if (piece == 1) iter_VC3_b(s, o, vc, (vc3_scratch *)h);
    else iter_VC3_c(s, o);
}
void orc_vc3_end(const orc_fields *s, const orc_stokes_opts *o, void *h, int finish)
{
    post_VC3(s, o, (vc3_scratch *)h, finish);
    free(h);
}

/* _solve! 3D-VC  Stokes3D.jl:447-668 */
int orc_solve3d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, orc_stokes_result *res)
{
    vc3_scratch w;
    double err_it1 = 1.0, err = INFINITY;
    int64_t iter = 0, cont = 0;
    int status = 0;
    pre_VC3(s, o, vc, &w);
    while (iter < 2 || (((err / err_it1) > o->eps_rel && err > o->eps_abs) && iter <= o->iterMax)) {
        iter_VC3_a(s); iter_VC3_b(s, o, vc, &w); iter_VC3_c(s, o);
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            double e[4];
            norms_VC3(s, o, e);
            res->norm_Rx[cont] = e[0]; res->norm_Ry[cont] = e[1]; res->norm_Rz[cont] = e[2]; res->norm_divV[cont] = e[3];
            err = fmax(fmax(e[0], e[1]), fmax(e[2], e[3]));
            if (isnan(e[0]) || isnan(e[1]) || isnan(e[2]) || isnan(e[3])) err = NAN;
            res->err_evo1[cont] = err; res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), fmax(res->norm_Rz[0], res->norm_divV[0]));
            if (isnan(err)) { status = 1; break; }
        }
    }
    post_VC3(s, o, &w, 1);
    res->iter = iter; res->nhist = cont; res->err = err;
    return status;
}
