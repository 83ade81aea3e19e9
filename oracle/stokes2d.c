/*
 * stokes2d.c — CPU ORACLE (test infrastructure, NOT product code): the 2D Stokes PT loops.
 *
 *   2D-V2  visco-elastic, arrays G, K            src/stokes/Stokes2D.jl:181-325      (config 2, SolCx)
 *   2D-VC  multiphase visco-elasto-plastic        src/stokes/Stokes2D.jl:577-866      (config 3, shear band)
 * kernels: compute_∇V! VelocityKernels.jl:3-6; compute_strain_rate! :10-44; compute_V! :108-131, 134-180; compute_Res!
 *   :246-307; compute_P! PressureKernels.jl:10-15, 64-105, 186-195; compute_τ! StressKernels.jl:63-91;
 *   update_stresses_center_vertex_ps! StressKernels.jl:992-1144 (+ clamped_indices / av_clamped / harm_clamped :1304-1319);
 *   plastic_params_phase, compute_yieldfunction_phase, compute_plastic_gradients_phase, cache_tensors
 *   src/rheology/StressUpdate.jl:146-188, 190-230, 384-452, 463-550; compute_viscosity_kernel! Viscosity.jl:382-418,
 *   local args :510-522, compute_phase_viscosity :599-619; compute_ρg! BuoyancyForces.jl:74-95;
 *   free_slip!/no_slip!/periodic_boundary! (2D) free_slip.jl:1-13, no_slip.jl:1-19, periodic.jl:15-35;
 *   exit: compute_vorticity! stress_rotation_particles.jl:17-30, shear2center! Interpolations.jl:306-311,
 *   accumulate_tensor!/accumulate_vol! StressKernels.jl:379-438, multi_copy!.
 *
 * GeoParams.jl (third party, compat 0.7.19, NOT vendored) — restated from its published definitions, pinned only
 * through the reference's integration goldens (test/test_shearband2D.jl:197-201):
 *   second_invariant(xx,yy,xy)        = √(½(xx²+yy²) + xy²)
 *   second_invariant_staggered(xx,yy,(4 shear)) = √(½(xx²+yy²) + Σ shear²/4)
 *   DruckerPrager[_regularised]:  F = τII − C cosϕ − P sinϕ (λ = 0 in the call), ∂Q∂τ = (½τxx/τII, ½τyy/τII, τxy/τII)
 *   [shear slot halved by StressUpdate.jl:470], ∂Q∂P = −sinΨ, ∂F∂P = −sinϕ
 *   compute_viscosity_τII(CompositeRheology): harmonic composition of the non-plastic elements with dt = Inf,
 *   τII_old = 0 (Viscosity.jl:517-520) → inv(inv(η_lin) + inv(G·Inf))
 *
 * @muladd (MuladdMacro.jl, third party) placement restated from its documented rule — the LAST product of a sum is
 * fused with the rest, an n-ary product splits into its first factor and the product of the others:
 *   (1−relλ)·λ + relλ·x → fma(relλ, x, (1−relλ)·λ);   dτ − 2·η·ε_pl·dτ_r → fma(−2, η·ε_pl·dτ_r, dτ)      (ulp-level, unpinned)
 *
 * Schedule of the racy stress kernel (quirk Q7): JACOBI — the vertex part reads the centre τxx, τyy as they were BEFORE the
 * launch (that is what a race-free parallel execution must define; the fixed point is unaffected).
 */
#include "jr_oracle.h"
#include "mini.h"
#include <stdlib.h>
#include <string.h>

#define F(name) (s->f[ORC_F_##name])
#define A2(p, n1, i, j) ((p)[IX2(n1, i, j)])

static arr mk2(const double *p, int n1, int n2) { arr a = {p, n1, n2, 1}; return a; }

/* ---- shared kernels ------------------------------------------------------------------------------------------ */
static void divV2(const orc_fields *s, const double _di[3], const double *Vx, const double *Vy, double *out)
{
    const int nx = s->n[0], ny = s->n[1];
    const arr ax = mk2(Vx, nx + 1, ny + 2), ay = mk2(Vy, nx + 2, ny + 1);
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) A2(out, nx, i, j) = d_xi2(ax, _di[0], i, j) + d_yi2(ay, _di[1], i, j);
}

static inline void compute_P_point(double *RP, double *P, double P0, double divV, double Q, double eta, double K, double G, double dt,
                                   double r, double theta_dtau)
{
    double _Kdt = orc_inv(K * dt), _Gdt = orc_inv(G * dt), _dt = orc_inv(dt), Pc = *P;
    *RP = fma(-(Pc - P0), _Kdt, (-divV + (Q * _dt)));
    double psi = orc_inv(orc_inv(eta) + _Gdt) * r / theta_dtau;
    *P = ((fma(P0, _Kdt, (-divV + (Q * _dt)))) * psi + Pc) / (1 + _Kdt * psi);
}

/* compute_strain_rate! 2D  VelocityKernels.jl:10-44 — on (V, ∇V) → ε, or on (U, ∇U) → Δε (strain-increment form, Stokes2D.jl:681-689) */
static void strain_rate2_of(const orc_fields *s, const double _di[3], const double *Ax, const double *Ay, const double *div, double *exx, double *eyy,
                            double *exy)
{
    const int nx = s->n[0], ny = s->n[1];
    const arr ax = mk2((double *)Ax, nx + 1, ny + 2), ay = mk2((double *)Ay, nx + 2, ny + 1);
    for (int j = 1; j <= ny + 1; j++)
        for (int i = 1; i <= nx + 1; i++) {
            if (i <= nx && j <= ny) {
                const double dV = A2(div, nx, i, j) * orc_inv(3.0);
                A2(exx, nx, i, j) = d_xi2(ax, _di[0], i, j) - dV;
                A2(eyy, nx, i, j) = d_yi2(ay, _di[1], i, j) - dV;
            }
            A2(exy, nx + 1, i, j) = 0.5 * (_di[1] * (AT2(ax, i, j + 1) - AT2(ax, i, j)) + _di[0] * (AT2(ay, i + 1, j) - AT2(ay, i, j)));
        }
}
static void strain_rate2(const orc_fields *s, const double _di[3]) { strain_rate2_of(s, _di, F(Vx), F(Vy), F(divV), F(exx), F(eyy), F(exy)); }

static inline double dtau_r(double th, double eta, double _Gdt) { return orc_inv(th + fma(eta, _Gdt, 1.0)); }
static inline double stress_inc(double t, double to, double eta, double e, double _Gdt, double dtr)
{
    return dtr * fma(2.0 * eta, e, fma(-(t - to) * eta, _Gdt, -t));
}

/* compute_τ! 2D visco-elastic  StressKernels.jl:63-91 */
static void tau2_VE(const orc_fields *s, double dt, double th)
{
    const int nx = s->n[0], ny = s->n[1];
    const arr eta = mk2(F(eta), nx, ny), G = mk2(F(G), nx, ny);
    for (int j = 1; j <= ny + 1; j++)
        for (int i = 1; i <= nx + 1; i++) {
            if (i <= nx && j <= ny) {
                const double _Gdt = orc_inv(AT2(G, i, j) * dt), e = AT2(eta, i, j), dtr = dtau_r(th, e, _Gdt);
                A2(F(txx), nx, i, j) += stress_inc(A2(F(txx), nx, i, j), A2(F(txx_o), nx, i, j), e, A2(F(exx), nx, i, j), _Gdt, dtr);
                A2(F(tyy), nx, i, j) += stress_inc(A2(F(tyy), nx, i, j), A2(F(tyy_o), nx, i, j), e, A2(F(eyy), nx, i, j), _Gdt, dtr);
            }
            const double e = av_ai_clamped2(eta, i, j), _Gdt = orc_inv(av_ai_clamped2(G, i, j) * dt), dtr = dtau_r(th, e, _Gdt);
            A2(F(txy), nx + 1, i, j) += stress_inc(A2(F(txy), nx + 1, i, j), A2(F(txy_o), nx + 1, i, j), e, A2(F(exy), nx + 1, i, j), _Gdt, dtr);
        }
}

/* compute_V!  VelocityKernels.jl:108-131 (fs < 0) or the free-surface form :134-180 with dt·free_surface = fs */
static void V2(const orc_fields *s, const orc_stokes_opts *o, int fs_form, double fs)
{
    const int nx = s->n[0], ny = s->n[1];
    const arr P = mk2(F(P), nx, ny), txx = mk2(F(txx), nx, ny), tyy = mk2(F(tyy), nx, ny), txy = mk2(F(txy), nx + 1, ny + 1);
    const arr fx = mk2(F(rhogx), nx, ny), fy = mk2(F(rhogy), nx, ny), ett = mk2(F(etatau), nx, ny);
    const double _dx = o->_di[0], _dy = o->_di[1], edt = o->eta_dtau;
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            if (i <= nx - 1)
                A2(F(Vx), nx + 1, i + 1, j + 1) += (-d_xa2(P, _dx, i, j) + d_xa2(txx, _dx, i, j) + d_yi2(txy, _dy, i, j) - av_xa2(fx, i, j)) * edt / av_xa2(ett, i, j);
            if (j <= ny - 1) {
                if (!fs_form)
                    A2(F(Vy), nx + 2, i + 1, j + 1) += (-d_ya2(P, _dy, i, j) + d_ya2(tyy, _dy, i, j) + d_xi2(txy, _dx, i, j) - av_ya2(fy, i, j)) * edt / av_ya2(ett, i, j);
                else {
                    const double Vy = A2(F(Vy), nx + 2, i + 1, j + 1);
                    const int jN = j + 1 < ny ? j + 1 : ny;
                    const double drg = (AT2(fy, i, jN) - AT2(fy, i, j)) * _dy;
                    const double corr = Vy * drg * 1.0 * fs;
                    A2(F(Vy), nx + 2, i + 1, j + 1) += (-d_ya2(P, _dy, i, j) + d_ya2(tyy, _dy, i, j) + d_xi2(txy, _dx, i, j) - av_ya2(fy, i, j) + corr) * edt / av_ya2(ett, i, j);
                }
            }
        }
}

/* compute_Res!  VelocityKernels.jl:246-307 */
static void Res2(const orc_fields *s, const orc_stokes_opts *o, int fs_form, double fs)
{
    const int nx = s->n[0], ny = s->n[1];
    const arr P = mk2(F(P), nx, ny), txx = mk2(F(txx), nx, ny), tyy = mk2(F(tyy), nx, ny), txy = mk2(F(txy), nx + 1, ny + 1);
    const arr fx = mk2(F(rhogx), nx, ny), fy = mk2(F(rhogy), nx, ny);
    const double _dx = o->_di[0], _dy = o->_di[1];
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            if (i <= nx - 1) A2(F(Rx), nx - 1, i, j) = d_xa2(txx, _dx, i, j) + d_yi2(txy, _dy, i, j) - d_xa2(P, _dx, i, j) - av_xa2(fx, i, j);
            if (j <= ny - 1) {
                double R = d_ya2(tyy, _dy, i, j) + d_xi2(txy, _dx, i, j) - d_ya2(P, _dy, i, j) - av_ya2(fy, i, j);
                if (fs_form) {
                    const double Vy = A2(F(Vy), nx + 2, i + 1, j + 1);
                    const int jN = j + 1 < ny ? j + 1 : ny;
                    const double drg = (AT2(fy, i, jN) - AT2(fy, i, j)) * _dy;
                    R = R + (Vy * drg) * 1.0 * fs;
                }
                A2(F(Ry), nx, i, j) = R;
            }
        }
}

static void v2u2(const orc_fields *s, double dt)
{
    const size_t nVx = (size_t)(s->n[0] + 1) * (s->n[1] + 2), nVy = (size_t)(s->n[0] + 2) * (s->n[1] + 1);
    if (F(Ux)) for (size_t q = 0; q < nVx; q++) F(Ux)[q] = F(Vx)[q] * dt;
    if (F(Uy)) for (size_t q = 0; q < nVy; q++) F(Uy)[q] = F(Vy)[q] * dt;
}

/* flow_bcs! 2D: no_slip! → free_slip! → periodic_boundary!  BoundaryConditions.jl:86-99.
 * flags: left,right,front,back,top,bot → 2D uses left(0), right(1), top(4), bot(5) */
void orc_flow_bcs2(const orc_fields *s, const orc_stokes_opts *o, int displacement)
{
    const int nx = s->n[0], ny = s->n[1];
    double *Ax = displacement ? F(Ux) : F(Vx), *Ay = displacement ? F(Uy) : F(Vy);
    const int n1x = nx + 1, n2x = ny + 2, n1y = nx + 2, n2y = ny + 1;
    const int32_t *ns = o->no_slip, *fs = o->free_slip, *pe = o->periodic;
    if (ns[0] | ns[1] | ns[4] | ns[5]) { /* no_slip.jl:1-19: sequential broadcasts */
        if (ns[0]) { for (int j = 1; j <= n2x; j++) A2(Ax, n1x, 1, j) = 0; for (int j = 1; j <= n2y; j++) A2(Ay, n1y, 1, j) = -A2(Ay, n1y, 2, j); }
        if (ns[1]) { for (int j = 1; j <= n2x; j++) A2(Ax, n1x, n1x, j) = 0; for (int j = 1; j <= n2y; j++) A2(Ay, n1y, n1y, j) = -A2(Ay, n1y, n1y - 1, j); }
        if (ns[5]) { for (int i = 1; i <= n1x; i++) A2(Ax, n1x, i, 1) = -A2(Ax, n1x, i, 2); for (int i = 1; i <= n1y; i++) A2(Ay, n1y, i, 1) = 0; }
        if (ns[4]) { for (int i = 1; i <= n1x; i++) A2(Ax, n1x, i, n2x) = -A2(Ax, n1x, i, n2x - 1); for (int i = 1; i <= n1y; i++) A2(Ay, n1y, i, n2y) = 0; }
    }
    if (fs[0] | fs[1] | fs[4] | fs[5]) { /* free_slip.jl:1-13 */
        for (int i = 1; i <= n1x; i++) {
            if (fs[5]) A2(Ax, n1x, i, 1) = A2(Ax, n1x, i, 2);
            if (fs[4]) A2(Ax, n1x, i, n2x) = A2(Ax, n1x, i, n2x - 1);
        }
        for (int j = 1; j <= n2y; j++) {
            if (fs[0]) A2(Ay, n1y, 1, j) = A2(Ay, n1y, 2, j);
            if (fs[1]) A2(Ay, n1y, n1y, j) = A2(Ay, n1y, n1y - 1, j);
        }
    }
    if (pe[0] | pe[1] | pe[4] | pe[5]) { /* periodic.jl:15-35, single-thread order i ascending */
        const int n = (n1y > n2x ? n1y : n2x) > (n1x > n2y ? n1x : n2y) ? (n1y > n2x ? n1y : n2x) : (n1x > n2y ? n1x : n2y);
        for (int i = 1; i <= n; i++) {
            if (i <= n2x && pe[0]) A2(Ax, n1x, 1, i) = A2(Ax, n1x, n1x, i);
            if (i <= n2y) {
                if (pe[0]) A2(Ay, n1y, 1, i) = A2(Ay, n1y, n1y - 1, i);
                if (pe[1]) A2(Ay, n1y, n1y, i) = A2(Ay, n1y, 2, i);
            }
            if (i <= n1x) {
                if (pe[5]) A2(Ax, n1x, i, 1) = A2(Ax, n1x, i, n2x - 1);
                if (pe[4]) A2(Ax, n1x, i, n2x) = A2(Ax, n1x, i, 2);
            }
            if (i <= n1y && pe[5]) A2(Ay, n1y, i, 1) = A2(Ay, n1y, i, n2y);
        }
    }
}

static void norms2(const orc_fields *s, const orc_stokes_opts *o, double out[3])
{
    const int nx = s->n[0], ny = s->n[1];
    const double gx = o->n_g[0], gy = o->n_g[1];
    out[0] = sqrt(orc_sumsq_interior(F(Rx), nx - 1, ny, 1, 1)) / sqrt((gx - 2) * (gy - 1));
    out[1] = sqrt(orc_sumsq_interior(F(Ry), nx, ny - 1, 1, 1)) / sqrt((gx - 1) * (gy - 2));
    out[2] = sqrt(orc_sumsq_interior(F(RP), nx, ny, 1, 0)) / sqrt(gx * gy);
}

static void multi_copy2(const orc_fields *s)
{
    const size_t nc = (size_t)s->n[0] * s->n[1], nv = (size_t)(s->n[0] + 1) * (s->n[1] + 1);
    memcpy(F(txx_o), F(txx), nc * 8); memcpy(F(tyy_o), F(tyy), nc * 8); memcpy(F(txy_o), F(txy), nv * 8);
    if (F(txy_c) && F(txy_o_c)) memcpy(F(txy_o_c), F(txy_c), nc * 8);
}

/* ---- 2D-V2 ---------------------------------------------------------------------------------------------------- */
static void pre_V2(const orc_fields *s) { orc_maxloc3(F(etatau), F(eta), s->n[0], s->n[1], 1, 1, 1, 0); }
static void iter_V2(const orc_fields *s, const orc_stokes_opts *o)
{
    const int nx = s->n[0], ny = s->n[1];
    divV2(s, o->_di, F(Vx), F(Vy), F(divV));
    for (size_t q = 0; q < (size_t)nx * ny; q++) /* compute_P! with ητ (quirk Q5)  Stokes2D.jl:231-233 */
        compute_P_point(&F(RP)[q], &F(P)[q], F(P0)[q], F(divV)[q], F(Q)[q], F(etatau)[q], F(K)[q], F(G)[q], o->dt, o->r, o->theta_dtau);
    strain_rate2(s, o->_di);
    tau2_VE(s, o->dt, o->theta_dtau);
    V2(s, o, 0, 0.0);
    v2u2(s, o->dt);
    orc_flow_bcs2(s, o, 0);
}
int orc_iterate2d_V2(const orc_fields *s, const orc_stokes_opts *o, int64_t niter)
{
    pre_V2(s);
    for (int64_t it = 0; it < niter; it++) iter_V2(s, o);
    Res2(s, o, 0, 0.0);
    return 0;
}
int orc_solve2d_V2(const orc_fields *s, const orc_stokes_opts *o, orc_stokes_result *res)
{
    double err_it1 = 1.0, err = 1.0;
    int64_t iter = 0, cont = 0;
    pre_V2(s);
    while (iter < 2 || (((err / err_it1) > o->eps_rel && err > o->eps_abs) && iter <= o->iterMax)) {
        iter_V2(s, o);
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            double e[3];
            Res2(s, o, 0, 0.0);
            norms2(s, o, e);
            res->norm_Rx[cont] = e[0]; res->norm_Ry[cont] = e[1]; res->norm_divV[cont] = e[2];
            err = fmax(fmax(e[0], e[1]), e[2]);
            if (isnan(e[0]) || isnan(e[1]) || isnan(e[2])) err = NAN;
            res->err_evo1[cont] = err; res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), res->norm_divV[0]);
        }
    }
    multi_copy2(s);
    res->iter = iter; res->nhist = cont; res->err = err;
    return 0;
}

#include "vc_common.h"

/* compute_ρg!  BuoyancyForces.jl:74-95: fn_ratio(compute_density, …, args) .* (g[1], g[3]); args.T sampled at I+1 (Q17) */
void orc_rhog2d(const orc_fields *s, const orc_vc_inputs *vc)
{
    const int nx = s->n[0], ny = s->n[1];
    const size_t nc = (size_t)nx * ny;
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            const size_t c = IX2(nx, i, j);
            const double T = F(T) ? A2(F(T), nx + 2, i + 1, j + 1) : 0.0, P = F(Pargs) ? F(Pargs)[c] : 0.0;
            double rho = 0.0;
            for (int p = 0; p < vc->nphase; p++) {
                const double r = vc->ph_center[(size_t)p * nc + c];
                if (r == 1.0) { rho = density(&vc->phases[p], T, P) * r; break; }
                rho += (r == 0.0) ? 0.0 : density(&vc->phases[p], T, P) * r;
            }
            if (vc->g_scalar) F(rhogy)[c] = rho * vc->g[2];
            else { F(rhogx)[c] = rho * vc->g[0]; F(rhogy)[c] = rho * vc->g[2]; }
        }
}
void orc_viscosity2d(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, double nu)
{
    const int nx = s->n[0], ny = s->n[1];
    const size_t nc = (size_t)nx * ny, nv = (size_t)(nx + 1) * (ny + 1);
    /* LinearViscous does not depend on the invariant: AII (and its eps() zero-guard) does not enter the subset */
    for (size_t c = 0; c < nc; c++) {
        const double ei = phase_viscosity(vc, vc->ph_center, nc, c);
        F(eta)[c] = clampd((1 - nu) * F(eta)[c] + nu * ei, o->visc_cutoff_lo, o->visc_cutoff_hi);
    }
    if (F(etav) && vc->ph_vertex)
        for (size_t v = 0; v < nv; v++) {
            const double ei = phase_viscosity(vc, vc->ph_vertex, nv, v);
            F(etav)[v] = clampd((1 - nu) * F(etav)[v] + nu * ei, o->visc_cutoff_lo, o->visc_cutoff_hi);
        }
}

/* compute_stress_increment, strain-increment form  StressKernels.jl:19-22 */
static inline double stress_inc_d(double t, double to, double eta, double de, double _G, double dtr, double dt)
{
    return dtr * fma(2.0 * eta, de, fma(-(t - to) * eta, _G, -t * dt));
}

/* update_stresses_center_vertex_ps! 2D  StressKernels.jl:992-1144 (ε form) and :1147-1302 (Δε form, inc != 0), Jacobi schedule.
 * Δε form: dτ_r = 1/(θ_dτ·dt + η/G + dt), increments from Δε with 1/G and −τ·dt, plastic terms carry the extra dt. */
static void stress_vep2(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, const double *theta, double *lam, double *lamv)
{
    const int inc = o->strain_increment;
    const int nx = s->n[0], ny = s->n[1];
    const size_t nc = (size_t)nx * ny, nv = (size_t)(nx + 1) * (ny + 1);
    const double dt = o->dt, th = o->theta_dtau, rel = o->lambda_relaxation;
    double *txx0 = (double *)malloc(nc * 8), *tyy0 = (double *)malloc(nc * 8);
    memcpy(txx0, F(txx), nc * 8); memcpy(tyy0, F(tyy), nc * 8);
#define AVC(p) (0.25 * (A2(p, nx, i0, j0) + A2(p, nx, ic, jc) + A2(p, nx, i0, jc) + A2(p, nx, ic, j0)))
    for (int j = 1; j <= ny + 1; j++)
        for (int i = 1; i <= nx + 1; i++) {
            const int i0 = orc_clamp(i - 1, 1, nx), ic = orc_clamp(i, 1, nx), j0 = orc_clamp(j - 1, 1, ny), jc = orc_clamp(j, 1, ny);
            const size_t v = IX2(nx + 1, i, j);
            /* ---- vertex ---- */
            {
                const double Pv = AVC(theta), exxv = inc ? AVC(F(dxx)) : AVC(F(exx)), eyyv = inc ? AVC(F(dyy)) : AVC(F(eyy)), txxv = AVC(txx0), tyyv = AVC(tyy0);
                const double txxov = AVC(F(txx_o)), tyyov = AVC(F(tyy_o)), EIIv = AVC(F(EII_pl));

                int is_pl; double eta_reg;
                plastic_params(vc, vc->ph_vertex, nv, v, &is_pl, &eta_reg);
                const double _Gdt = orc_inv(ratio_G(vc, vc->ph_vertex, nv, v) * dt), Kv = ratio_Kb(vc, vc->ph_vertex, nv, v);
                const double _G = orc_inv(ratio_G(vc, vc->ph_vertex, nv, v));
                const double etav = 4 / (1 / A2(F(eta), nx, i0, j0) + 1 / A2(F(eta), nx, ic, jc) + 1 / A2(F(eta), nx, i0, jc) + 1 / A2(F(eta), nx, ic, j0));
                const double dtr = inc ? orc_inv(th * dt + etav * _G + dt) : orc_inv(th + etav * _Gdt + 1.0);
                const double txyv = F(txy)[v];
                const double dxx = inc ? stress_inc_d(txxv, txxov, etav, exxv, _G, dtr, dt) : stress_inc(txxv, txxov, etav, exxv, _Gdt, dtr);
                const double dyy = inc ? stress_inc_d(tyyv, tyyov, etav, eyyv, _G, dtr, dt) : stress_inc(tyyv, tyyov, etav, eyyv, _Gdt, dtr);
                const double dxy = inc ? stress_inc_d(txyv, F(txy_o)[v], etav, F(dxy)[v], _G, dtr, dt) : stress_inc(txyv, F(txy_o)[v], etav, F(exy)[v], _Gdt, dtr);
                const double trial[3] = {txxv + dxx, tyyv + dyy, txyv + dxy};
                const double tII = second_invariant3(dxx + txxv, dyy + tyyv, dxy + txyv);
                double dQ[3], dQdP, dFdP;
                plastic_grads(vc, vc->ph_vertex, nv, v, trial, dQ, &dQdP, &dFdP);
                const double volume = isinf(Kv) ? 0.0 : Kv * dt * dFdP * dQdP;
                const double Fv = yield_F_soft(vc, vc->ph_vertex, nv, v, Pv, tII, EIIv);
                if (is_pl && tII != 0.0 && Fv > 0) {
                    lamv[v] = fma(rel, fmax(Fv, 0.0) / ((inc ? etav * dtr * dt : etav * dtr) + eta_reg + volume), (1.0 - rel) * lamv[v]);
                    const double epl = lamv[v] * dQ[2];
                    F(txy)[v] += inc ? fma(-2.0, etav * dt * epl * dtr, dxy) : fma(-2.0, etav * epl * dtr, dxy);
                    F(pxy)[v] = epl;
                } else {
                    F(txy)[v] += dxy;
                    F(pxy)[v] = 0.0;
                }
            }
            /* ---- centre ---- */
            if (i <= nx && j <= ny) {
                const size_t c = IX2(nx, i, j);
                const double _Gdt = orc_inv(ratio_G(vc, vc->ph_center, nc, c) * dt);
                int is_pl; double eta_reg;
                plastic_params(vc, vc->ph_center, nc, c, &is_pl, &eta_reg);
                const double K = ratio_Kb(vc, vc->ph_center, nc, c), eta = F(eta)[c];
                const double _G = orc_inv(ratio_G(vc, vc->ph_center, nc, c));
                const double dtr = inc ? 1.0 / (th * dt + eta * _G + dt) : 1.0 / (th + eta * _Gdt + 1.0);
                const arr exyv = mk2(F(exy), nx + 1, ny + 1);
                const double eij[3] = {F(exx)[c], F(eyy)[c],
                                       (((AT2(exyv, i, j) + AT2(exyv, i + 1, j)) + AT2(exyv, i, j + 1)) + AT2(exyv, i + 1, j + 1)) / 4};
                double tij[3] = {F(txx)[c], F(tyy)[c], F(txy_c)[c]};
                const double tijo[3] = {F(txx_o)[c], F(tyy_o)[c], F(txy_o_c)[c]};
                double dt_[3];
                if (inc) {
                    const arr dxyv = mk2(F(dxy), nx + 1, ny + 1);
                    const double dij[3] = {F(dxx)[c], F(dyy)[c], (((AT2(dxyv, i, j) + AT2(dxyv, i + 1, j)) + AT2(dxyv, i, j + 1)) + AT2(dxyv, i + 1, j + 1)) / 4};
                    for (int q = 0; q < 3; q++) dt_[q] = stress_inc_d(tij[q], tijo[q], eta, dij[q], _G, dtr, dt);
                } else
                    for (int q = 0; q < 3; q++) dt_[q] = stress_inc(tij[q], tijo[q], eta, eij[q], _Gdt, dtr);
                double tII = second_invariant3(dt_[0] + tij[0], dt_[1] + tij[1], dt_[2] + tij[2]);
                const double trial[3] = {tij[0] + dt_[0], tij[1] + dt_[1], tij[2] + dt_[2]};
                double dQ[3], dQdP, dFdP;
                const double Pr = theta[c];
                plastic_grads(vc, vc->ph_center, nc, c, trial, dQ, &dQdP, &dFdP);
                const double volume = isinf(K) ? 0.0 : K * dt * dFdP * dQdP;
                const double Fc = yield_F_soft(vc, vc->ph_center, nc, c, Pr, tII, F(EII_pl)[c]);
                if (is_pl && tII != 0.0 && Fc > 0) {
                    lam[c] = fma(rel, fmax(Fc, 0.0) / ((inc ? eta * dtr * dt : eta * dtr) + eta_reg + volume), (1.0 - rel) * lam[c]);
                    double epl[3];
                    for (int q = 0; q < 3; q++) {
                        epl[q] = lam[c] * dQ[q];
                        dt_[q] = inc ? fma(-2.0, eta * dt * epl[q] * dtr, dt_[q]) : fma(-2.0, eta * epl[q] * dtr, dt_[q]);
                        tij[q] = dt_[q] + tij[q];
                    }
                    F(e_vol_pl)[c] = -lam[c] * dQdP;
                    F(txx)[c] = tij[0]; F(tyy)[c] = tij[1]; F(txy_c)[c] = tij[2];
                    F(pxx)[c] = epl[0]; F(pyy)[c] = epl[1];
                    tII = second_invariant3(tij[0], tij[1], tij[2]);
                } else {
                    F(e_vol_pl)[c] = 0.0;
                    F(txx)[c] = dt_[0] + tij[0]; F(tyy)[c] = dt_[1] + tij[1]; F(txy_c)[c] = dt_[2] + tij[2];
                    F(pxx)[c] = 0.0; F(pyy)[c] = 0.0;
                }
                F(tII)[c] = tII;
                F(eta_vep)[c] = tII * 0.5 * orc_inv(second_invariant3(eij[0], eij[1], eij[2]));
                F(P)[c] = Pr - (isinf(K) ? 0.0 : K * dt * lam[c] * dQdP);
            }
        }
#undef AVC
    free(txx0); free(tyy0);
}

/* second_invariant_staggered(xx, yy, gather(xy))  — tensor_invariant!  StressKernels.jl:470-480 */
static inline double inv_stag2(double xx, double yy, const double *xy, int n1, int i, int j)
{
    const double a = A2(xy, n1, i, j), b = A2(xy, n1, i + 1, j), c = A2(xy, n1, i, j + 1), d = A2(xy, n1, i + 1, j + 1);
    return sqrt(0.5 * (xx * xx + yy * yy) + (((a * a + b * b) + c * c) + d * d) / 4);
}
void orc_tensor_invariant2d(double *II, const double *xx, const double *yy, const double *xy, int nx, int ny)
{
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) A2(II, nx, i, j) = inv_stag2(A2(xx, nx, i, j), A2(yy, nx, i, j), xy, nx + 1, i, j);
}

static void shear2center2(double *c, const double *v, int nx, int ny)
{
    if (!c || !v) return;
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++)
            A2(c, nx, i, j) = 0.25 * (A2(v, nx + 1, i, j) + A2(v, nx + 1, i + 1, j) + A2(v, nx + 1, i, j + 1) + A2(v, nx + 1, i + 1, j + 1));
}

typedef struct { double *theta, *lam, *lamv; } vc_scratch;

static void pre_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, vc_scratch *w)
{
    const int nx = s->n[0], ny = s->n[1];
    const size_t nc = (size_t)nx * ny, nv = (size_t)(nx + 1) * (ny + 1);
    memcpy(F(P0), F(P), nc * 8);                                   /* @copy stokes.P0 stokes.P            :609 */
    orc_maxloc3(F(etatau), F(eta), nx, ny, 1, 1, 1, 0);            /* ητ = deepcopy(η); compute_maxloc!   :611-614 */
    w->theta = (double *)malloc(nc * 8); memcpy(w->theta, F(P), nc * 8);   /* θ = deepcopy(stokes.P)       :635 */
    w->lam = (double *)calloc(nc, 8); w->lamv = (double *)calloc(nv, 8);   /* λ, λv = 0                    :636-637 */
    memset(F(pxx), 0, nc * 8); memset(F(pyy), 0, nc * 8);          /* @tensor_center(ε_pl) .= 0           :641-643 */
    if (F(pxy_c)) memset(F(pxy_c), 0, nc * 8);
    orc_rhog2d(s, vc);                                             /* compute_ρg!                         :646 */
    if (o->displacement_bcs) {                                     /* displacement2velocity!(stokes, dt, flow_bcs) :647 ; types/displacement.jl:33-70 */
        const size_t nVx = (size_t)(nx + 1) * (ny + 2), nVy = (size_t)(nx + 2) * (ny + 1);
        const double _dt = orc_inv(o->dt);
        for (size_t q = 0; q < nVx; q++) F(Vx)[q] = F(Ux)[q] * _dt;
        for (size_t q = 0; q < nVy; q++) F(Vy)[q] = F(Uy)[q] * _dt;
    }
}

static void iter_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, vc_scratch *w)
{
    const int nx = s->n[0], ny = s->n[1];
    const size_t nc = (size_t)nx * ny;
    orc_maxloc3(F(etatau), F(eta), nx, ny, 1, 1, 1, 0);
    divV2(s, o->_di, F(Vx), F(Vy), F(divV));
    for (size_t c = 0; c < nc; c++) { /* compute_P!(θ, P0, RP, ∇V, Q, ητ, rheology, phase_ratios, …)  :664-677 */
        const double K = ratio_Kb(vc, vc->ph_center, nc, c), G = ratio_G(vc, vc->ph_center, nc, c);
        if (F(dTargs)) /* args.ΔT given: compute_P_kernel!(…, ΔT, ::Nothing)  PressureKernels.jl:128-149 */
            P_point_dT(&F(RP)[c], &w->theta[c], F(P0)[c], F(divV)[c], F(Q)[c], F(dTargs)[o->dT_ghosted ? IX2(nx + 2, c % nx + 1, c / nx + 1) : c], ratio_alpha(vc, vc->ph_center, nc, c), F(etatau)[c], K, G,
                       o->dt, o->r, o->theta_dtau);
        else
            compute_P_point(&F(RP)[c], &w->theta[c], F(P0)[c], F(divV)[c], F(Q)[c], F(etatau)[c], K, G, o->dt, o->r, o->theta_dtau);
    }
    if (!density_is_constant(vc)) orc_rhog2d(s, vc);               /* update_ρg!                           :679 */
    if (o->strain_increment) {                                     /* :660-662, 681-695 */
        const size_t nv = (size_t)(nx + 1) * (ny + 1);
        const double _dt = orc_inv(o->dt);
        divV2(s, o->_di, F(Ux), F(Uy), F(divU));                   /* compute_∇V!(stokes.∇U, @displacement(stokes), …) */
        strain_rate2_of(s, o->_di, F(Ux), F(Uy), F(divU), F(dxx), F(dyy), F(dxy));   /* Δε from U */
        for (size_t c = 0; c < nc; c++) { F(exx)[c] = F(dxx)[c] * _dt; F(eyy)[c] = F(dyy)[c] * _dt; }   /* compute_strain_rate_from_increment! */
        for (size_t v = 0; v < nv; v++) F(exy)[v] = F(dxy)[v] * _dt;
    } else
        strain_rate2(s, o->_di);
    stress_vep2(s, o, vc, w->theta, w->lam, w->lamv);
    orc_viscosity2d(s, o, vc, o->viscosity_relaxation);            /* update_viscosity_τII! AFTER the stress kernel (quirk Q13) */
    V2(s, o, 1, vc->free_surface);
    v2u2(s, o->dt);
    orc_flow_bcs2(s, o, o->displacement_bcs);                      /* flow_bcs!(stokes, flow_bcs): on U for DisplacementBoundaryConditions */
}

static void post_VC(const orc_fields *s, const orc_stokes_opts *o, vc_scratch *w)
{
    const int nx = s->n[0], ny = s->n[1];
    const size_t nc = (size_t)nx * ny;
    const arr ax = mk2(F(Vx), nx + 1, ny + 2), ay = mk2(F(Vy), nx + 2, ny + 1);
    if (F(wxy))
        for (int j = 1; j <= ny + 1; j++)
            for (int i = 1; i <= nx + 1; i++) A2(F(wxy), nx + 1, i, j) = 0.5 * (d_xa2(ay, o->_di[0], i, j) - d_ya2(ax, o->_di[1], i, j));
    shear2center2(F(exy_c), F(exy), nx, ny);
    shear2center2(F(pxy_c), F(pxy), nx, ny);
    shear2center2(F(dxy_c), F(dxy), nx, ny);
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) { /* accumulate_tensor!, accumulate_vol! */
            const size_t c = IX2(nx, i, j);
            F(EII_pl)[c] += inv_stag2(F(pxx)[c], F(pyy)[c], F(pxy), nx + 1, i, j) * o->dt;
            F(EVol_pl)[c] += o->dt * F(e_vol_pl)[c];
        }
    multi_copy2(s);
    (void)nc;
    free(w->theta); free(w->lam); free(w->lamv);
}

int orc_iterate2d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, int64_t niter, int finish)
{
    vc_scratch w;
    pre_VC(s, o, vc, &w);
    for (int64_t it = 0; it < niter; it++) iter_VC(s, o, vc, &w);
    Res2(s, o, 1, vc->free_surface);
    if (F(lam)) memcpy(F(lam), w.lam, (size_t)s->n[0] * s->n[1] * 8);      /* expose the solver-local λ, λv for parity checks */
    if (F(lamv)) memcpy(F(lamv), w.lamv, (size_t)(s->n[0] + 1) * (s->n[1] + 1) * 8);
    if (finish) post_VC(s, o, &w);
    else { free(w.theta); free(w.lam); free(w.lamv); }
    return 0;
}

int orc_solve2d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, orc_stokes_result *res)
{
    vc_scratch w;
    double err_it1 = 1.0, err = 1.0;
    int64_t iter = 0, cont = 0;
    int status = 0;
    pre_VC(s, o, vc, &w);
    while (iter <= o->iterMax) {
        if (o->iterMin < iter && ((err / err_it1) < o->eps_rel || err < o->eps_abs)) break;
        iter_VC(s, o, vc, &w);
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            double e[3];
            Res2(s, o, 1, vc->free_surface);
            norms2(s, o, e);
            res->norm_Rx[cont] = e[0]; res->norm_Ry[cont] = e[1]; res->norm_divV[cont] = e[2];
            err = fmax(fmax(e[0], e[1]), e[2]);
            if (isnan(e[0]) || isnan(e[1]) || isnan(e[2])) err = NAN;
            res->err_evo1[cont] = err; res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), res->norm_divV[0]);
            if (isnan(err)) { status = 1; break; }
        }
    }
    if (F(lam)) memcpy(F(lam), w.lam, (size_t)s->n[0] * s->n[1] * 8);
    if (F(lamv)) memcpy(F(lamv), w.lamv, (size_t)(s->n[0] + 1) * (s->n[1] + 1) * 8);
    post_VC(s, o, &w);
    res->iter = iter; res->nhist = cont; res->err = err;
    return status;
}
