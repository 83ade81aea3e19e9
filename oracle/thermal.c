/*
 * thermal.c — CPU ORACLE (test infrastructure) for `heatdiffusion_PT!`, 2D and 3D.
 *
 * Restates, kernel by kernel and in the reference's operation order,
 *   _heatdiffusion_PT!           src/thermal_diffusion/DiffusionPT_solver.jl:34-149 (arrays K, ρCp) and :181-305 (rheology)
 *   compute_flux!                src/thermal_diffusion/DiffusionPT_kernels.jl:6-61, 63-158 (3D), :327-364, 366-517 (2D)
 *   update_T!                    :160-199, 201-248 (3D), :519-551, 553-601 (2D)
 *   check_res!                   :250-323 (3D), :603-668 (2D);  update_ΔT! :670-673;  adiabatic_heating :720-731
 *   compute_pt_thermal_arrays!   src/thermal_diffusion/DiffusionPT_coefficients.jl:105-151
 *   compute_ρCp, compute_radioactive_heating, fn_ratio   DiffusionPT_GeoParams.jl:104-150, src/phases/phases.jl:5-30
 *   thermal_bcs!                 src/boundaryconditions/BoundaryConditions.jl:39-54, constant_value.jl:1-32,
 *                                free_slip.jl:72-103, periodic.jl:1-13,37-54
 *   Dirichlet mask               src/boundaryconditions/Dirichlet.jl:72-99, src/mask/mask.jl:47-59
 * GeoParams (third party, compat 0.7.19, not vendored) laws restated from their published definitions:
 *   ConstantDensity ρ = ρ0; PT_Density ρ = ρ0 (1 − α (T − T0) + β (P − P0)); T_Density ρ = ρ0 (1 − α (T − T0));
 *   ConstantHeatCapacity Cp; ConstantConductivity k; ConstantRadioactiveHeat H_r.
 *
 * Indices: arrays are column-major; T, Told, ΔT (and the Dirichlet mask/value) carry one ghost layer (n+2 per dim).
 * Boundary flags are per dimension and side (lo/hi): 2D  x: left/right, y: bot/top;  3D  x: left/right, y: front/back,
 * z: bot/top — the caller passes the reference's named faces in the order left,right,front,back,top,bot.
 * The boundary kernels run in the order a single thread executes the reference kernel (i fastest, then j), which fixes
 * the values of ghost edges/corners (they depend on thread order in the reference; nothing on the path reads them).
 */
#include "jr_oracle.h"
#include "thermal.h"
#include <string.h>

#define PI_ 3.141592653589793 /* Julia's Float64(π) */

/* ---- geometry helpers ------------------------------------------------------------------------------------- */
typedef struct { int nd, nx, ny, nz, gx, gy, gz; } dims_t; /* g* = ghosted extents of T */
static dims_t mkdims(const orc_thermal_fields *f)
{
    dims_t d;
    d.nd = f->ndim; d.nx = f->n[0]; d.ny = f->n[1]; d.nz = f->ndim == 3 ? f->n[2] : 1;
    d.gx = d.nx + 2; d.gy = d.ny + 2; d.gz = f->ndim == 3 ? d.nz + 2 : 1;
    return d;
}
/* 0-based linear indices */
#define TI(d, i, j, k) ((size_t)(k) * (d).gy * (d).gx + (size_t)(j) * (d).gx + (size_t)(i))          /* ghosted T index */
#define CI(d, i, j, k) ((size_t)(k) * (d).ny * (d).nx + (size_t)(j) * (d).nx + (size_t)(i))          /* cell index     */
static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* per-dimension boundary flags from the named faces (left,right,front,back,top,bot) */
static void face_map(int nd, int lo[3], int hi[3])
{
    lo[0] = 0; hi[0] = 1;
    if (nd == 2) { lo[1] = 5; hi[1] = 4; lo[2] = hi[2] = -1; }
    else { lo[1] = 2; hi[1] = 3; lo[2] = 5; hi[2] = 4; }
}

/* ---- GeoParams subset --------------------------------------------------------------------------------------- */
static inline double ph_density(const orc_thermal_phase *p, double T, double P)
{
    switch (p->rho_kind) {
    case 1: return p->rho0 * (1.0 - p->alpha * (T - p->T0) + p->beta * (P - p->P0));
    case 2: return p->rho0 * (1.0 - p->alpha * (T - p->T0));
    default: return p->rho0;
    }
}
static inline double ph_rhoCp(const orc_thermal_phase *p, double T, double P) { return p->Cp * ph_density(p, T, P); }

/* fn_ratio(fn, rheology, ratio, args)  phases.jl:18-30 — a ratio equal to one returns that phase's value alone */
static double ratio_rhoCp(const orc_thermal_opts *o, const double *ph, size_t stride, size_t idx, double T, double P)
{
    if (!ph) return ph_rhoCp(&o->phases[0], T, P);
    double x = 0.0;
    for (int p = 0; p < o->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        if (r == 1.0) return ph_rhoCp(&o->phases[p], T, P) * r;
        x += (r == 0.0) ? 0.0 : ph_rhoCp(&o->phases[p], T, P) * r;
    }
    return x;
}
/* compute_conductivity(phase, args): ConstantConductivity k, or TP_Conductivity k = (a + b / (T + c)) (1 + d P)
 * (GeoParams.jl, not vendored: restated from its docstring; used by miniapps/convection/Particles3D/Layered_rheology.jl:45-57) */
static inline double ph_cond(const orc_thermal_phase *p, double T, double P)
{
    return p->k_kind == 1 ? (p->k_a + p->k_b / (T + p->k_c)) * (1.0 + p->k_d * P) : p->k;
}
static double ratio_K(const orc_thermal_opts *o, const double *ph, size_t stride, size_t idx, double T, double P)
{
    if (!ph) return ph_cond(&o->phases[0], T, P);
    double x = 0.0;
    for (int p = 0; p < o->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        if (r == 1.0) return ph_cond(&o->phases[p], T, P) * r;
        x += (r == 0.0) ? 0.0 : ph_cond(&o->phases[p], T, P) * r;
    }
    return x;
}
/* fn_ratio(fn, rheology, ratio)  phases.jl:5-16 (no early return) */
static double ratio_Hr(const orc_thermal_opts *o, const double *ph, size_t stride, size_t idx)
{
    if (!ph) return o->phases[0].has_Hr ? o->phases[0].Hr : 0.0;
    double x = 0.0;
    for (int p = 0; p < o->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        x += (r == 0.0) ? 0.0 : o->phases[p].Hr * r;
    }
    return x;
}
static double ratio_alpha(const orc_thermal_opts *o, const double *ph, size_t stride, size_t idx)
{
    if (!ph) return o->phases[0].alpha;
    double x = 0.0;
    for (int p = 0; p < o->nphase; p++) {
        const double r = ph[(size_t)p * stride + idx];
        x += (r == 0.0) ? 0.0 : o->phases[p].alpha * r;
    }
    return x;
}

/* ---- compute_pt_thermal_arrays!  DiffusionPT_coefficients.jl:105-151 ---------------------------------------- */
void orc_thermal_pt_arrays(const orc_thermal_fields *f, const orc_thermal_opts *o)
{
    const dims_t d = mkdims(f);
    const size_t nc = (size_t)d.nx * d.ny * d.nz;
    const double _dt = 1.0 / o->dt, L = o->max_lxyz;
    for (int k = 0; k < d.nz; k++)
        for (int j = 0; j < d.ny; j++)
            for (int i = 0; i < d.nx; i++) {
                const size_t c = CI(d, i, j, k);
                const double T = f->T[TI(d, i + 1, j + 1, d.nd == 3 ? k + 1 : 0)], P = f->P ? f->P[c] : 0.0;
                const double rhoCp = ratio_rhoCp(o, f->phase_c, nc, c, T, P);
                const double _K = 1.0 / ratio_K(o, f->phase_c, nc, c, T, P);   /* args_ij = (T[I+1], P[I])  DiffusionPT_coefficients.jl:125 */
                const double _Re = 1.0 / (PI_ + sqrt(PI_ * PI_ + rhoCp * (L * L) * _K * _dt));
                f->theta_r_dtau[c] = L / o->Vpdtau * _Re;
                f->dtau_rho[c] = o->Vpdtau * L * _K * _Re;
            }
}

/* ---- compute_flux! ------------------------------------------------------------------------------------------ */
/* one direction: dim = 0,1,2.  face array extents: n + e_dim */
static void flux_dim(const orc_thermal_fields *f, const orc_thermal_opts *o, int dim)
{
    const dims_t d = mkdims(f);
    int lo[3], hi[3];
    face_map(d.nd, lo, hi);
    double *q = dim == 0 ? f->qTx : dim == 1 ? f->qTy : f->qTz, *q2 = dim == 0 ? f->qTx2 : dim == 1 ? f->qTy2 : f->qTz2;
    const double *phf = dim == 0 ? f->phase_x : dim == 1 ? f->phase_y : f->phase_z;
    const int ncell[3] = {d.nx, d.ny, d.nz};
    int e[3] = {d.nx, d.ny, d.nz};
    e[dim] += 1;
    /* extents of the face phase-ratio array (PhaseRatios.Vx etc.: n + e_dim); indexed with clamped CENTRE indices (Q9) */
    const size_t pstride = (size_t)e[0] * e[1] * e[2];
    const size_t nc = (size_t)d.nx * d.ny * d.nz;
    const int g3 = d.nd == 3;
    for (int k = 0; k < e[2]; k++)
        for (int j = 0; j < e[1]; j++)
            for (int i = 0; i < e[0]; i++) {
                const int I[3] = {i, j, k};
                const size_t qi = ((size_t)k * e[1] + j) * e[0] + i;
                if (I[dim] == 0 && o->cf_active[lo[dim]]) { q[qi] = o->cf_value[lo[dim]]; continue; }
                if (I[dim] == e[dim] - 1 && o->cf_active[hi[dim]]) { q[qi] = o->cf_value[hi[dim]]; continue; }
                int L[3] = {i, j, k}, R[3] = {i, j, k};
                L[dim] = clampi(I[dim] - 1, 0, ncell[dim] - 1);
                R[dim] = clampi(I[dim], 0, ncell[dim] - 1);
                const size_t cL = CI(d, L[0], L[1], L[2]), cR = CI(d, R[0], R[1], R[2]);
                /* T at the two cells adjacent to the face, ghosted indices: low side = I, high side = I + e_dim */
                int tl[3] = {i + 1, j + 1, g3 ? k + 1 : 0}, th[3] = {i + 1, j + 1, g3 ? k + 1 : 0};
                tl[dim] = I[dim]; th[dim] = I[dim] + 1;
                const double Tl = f->T[TI(d, tl[0], tl[1], tl[2])], Th = f->T[TI(d, th[0], th[1], th[2])];
                double K;
                if (o->form == 0) K = (f->K[cL] + f->K[cR]) * 0.5;
                else {
                    /* args: T = mean of the two nodes adjacent to the face, the other args (P) of the clamped cell on either side
                     * DiffusionPT_kernels.jl:93-100 (3D), 391-402 (2D) */
                    const size_t pL = ((size_t)L[2] * e[1] + L[1]) * e[0] + L[0], pR = ((size_t)R[2] * e[1] + R[1]) * e[0] + R[0];
                    const double Tf = (Tl + Th) * 0.5, PL = f->P ? f->P[cL] : 0.0, PR = f->P ? f->P[cR] : 0.0;
                    K = (ratio_K(o, phf, pstride, pL, Tf, PL) + ratio_K(o, phf, pstride, pR, Tf, PR)) * 0.5;
                }
                const double th_ = (f->theta_r_dtau[cL] + f->theta_r_dtau[cR]) * 0.5;
                const double qx = -K * (Th - Tl) * o->_di[dim];
                q2[qi] = qx;
                q[qi] = (q[qi] * th_ + qx) / (1.0 + th_);
                (void)nc;
            }
}
void orc_thermal_flux(const orc_thermal_fields *f, const orc_thermal_opts *o)
{
    /* the reference kernel handles x, y, z of node I in one thread; the three parts touch disjoint arrays */
    for (int dim = 0; dim < f->ndim; dim++) flux_dim(f, o, dim);
}

/* ---- update_T! ---------------------------------------------------------------------------------------------- */
static inline double div_q(const orc_thermal_fields *f, const orc_thermal_opts *o, const dims_t d, int i, int j, int k, int second)
{
    const double *qx = second ? f->qTx2 : f->qTx, *qy = second ? f->qTy2 : f->qTy, *qz = second ? f->qTz2 : f->qTz;
    const size_t ix = ((size_t)k * d.ny + j) * (d.nx + 1) + i, iy = ((size_t)k * (d.ny + 1) + j) * d.nx + i;
    double s = (qx[ix + 1] - qx[ix]) * o->_di[0] + (qy[iy + d.nx] - qy[iy]) * o->_di[1];
    if (d.nd == 3) {
        const size_t iz = ((size_t)k * d.ny + j) * d.nx + i;
        s = s + (qz[iz + (size_t)d.nx * d.ny] - qz[iz]) * o->_di[2];
    }
    return s;
}

void orc_thermal_update_T(const orc_thermal_fields *f, const orc_thermal_opts *o)
{
    const dims_t d = mkdims(f);
    const size_t nc = (size_t)d.nx * d.ny * d.nz;
    const double _dt = 1.0 / o->dt;
    for (int k = 0; k < d.nz; k++)
        for (int j = 0; j < d.ny; j++)
            for (int i = 0; i < d.nx; i++) {
                const size_t c = CI(d, i, j, k), t = TI(d, i + 1, j + 1, d.nd == 3 ? k + 1 : 0);
                if (f->dir_mask && f->dir_mask[t] != 0.0) {
                    /* apply_mask!: A = inv(m)·A + m·B  (mask.jl:51-52) */
                    const double m = f->dir_mask[t], B = f->dir_value ? f->dir_value[t] : o->dir_const;
                    f->T[t] = (1 - m) * f->T[t] + m * B;
                    continue;
                }
                const double Tc = f->T[t];
                if (o->form == 0) {
                    f->T[t] = (f->dtau_rho[c] * (-(div_q(f, o, d, i, j, k, 0)) + f->Told[t] * f->rhoCp[c] * _dt + f->H[c] + f->shear_heating[c]) + Tc) /
                              (1.0 + f->dtau_rho[c] * f->rhoCp[c] * _dt);
                } else {
                    const double P = f->P ? f->P[c] : 0.0;
                    const double rhoCp = ratio_rhoCp(o, f->phase_c, nc, c, Tc, P);
                    f->T[t] = (f->dtau_rho[c] * (-(div_q(f, o, d, i, j, k, 0)) + f->Told[t] * rhoCp * _dt + ratio_Hr(o, f->phase_c, nc, c) + f->H[c] +
                                                 f->shear_heating[c] + f->adiabatic[c] * Tc) + Tc) /
                              (1.0 + f->dtau_rho[c] * rhoCp * _dt);
                }
            }
}

/* ---- check_res! --------------------------------------------------------------------------------------------- */
void orc_thermal_check_res(const orc_thermal_fields *f, const orc_thermal_opts *o)
{
    const dims_t d = mkdims(f);
    const size_t nc = (size_t)d.nx * d.ny * d.nz;
    const double _dt = 1.0 / o->dt;
    for (int k = 0; k < d.nz; k++)
        for (int j = 0; j < d.ny; j++)
            for (int i = 0; i < d.nx; i++) {
                const size_t c = CI(d, i, j, k), t = TI(d, i + 1, j + 1, d.nd == 3 ? k + 1 : 0);
                if (f->dir_mask && f->dir_mask[t] != 0.0) { f->ResT[c] = 0.0; continue; }
                const double Tc = f->T[t];
                if (o->form == 0)
                    f->ResT[c] = -f->rhoCp[c] * (Tc - f->Told[t]) * _dt - div_q(f, o, d, i, j, k, 1) + f->H[c] + f->shear_heating[c];
                else {
                    const double rhoCp = ratio_rhoCp(o, f->phase_c, nc, c, Tc, f->P ? f->P[c] : 0.0);
                    f->ResT[c] = -rhoCp * (Tc - f->Told[t]) * _dt - div_q(f, o, d, i, j, k, 1) + ratio_Hr(o, f->phase_c, nc, c) + f->H[c] +
                                 f->shear_heating[c] + f->adiabatic[c] * Tc;
                }
            }
}

/* ---- thermal_bcs!  (constant_value → no_flux → periodic; each a literal single-thread sweep of the reference kernel) */
typedef enum { BC_CV, BC_NOFLUX, BC_PERIODIC } bckind;
static inline void bc_set(double *T, const dims_t d, bckind kind, int active, double val, int i, int j, int k, int dim, int hi)
{
    if (!active) return;
    int g[3] = {d.gx, d.gy, d.gz}, c[3] = {i, j, k}, s[3] = {i, j, k};
    c[dim] = hi ? g[dim] - 1 : 0;
    if (kind == BC_PERIODIC) s[dim] = hi ? 1 : g[dim] - 2;
    else s[dim] = hi ? g[dim] - 2 : 1;
    const double src = T[TI(d, s[0], s[1], s[2])];
    T[TI(d, c[0], c[1], c[2])] = kind == BC_CV ? 2 * val - src : src;
}
static void bc_sweep(const orc_thermal_fields *f, const orc_thermal_opts *o, bckind kind)
{
    const dims_t d = mkdims(f);
    int lo[3], hi[3];
    face_map(d.nd, lo, hi);
    const int32_t *act = kind == BC_CV ? o->cv_active : kind == BC_NOFLUX ? o->no_flux : o->periodic;
    const double *val = o->cv_value;
    double *T = f->T;
    if (d.nd == 2) {
        const int n = d.gx > d.gy ? d.gx : d.gy;
        for (int i = 0; i < n; i++) {
            if (i < d.gx) { /* bot, top */
                bc_set(T, d, kind, act[lo[1]], val[lo[1]], i, 0, 0, 1, 0);
                bc_set(T, d, kind, act[hi[1]], val[hi[1]], i, 0, 0, 1, 1);
            }
            if (i < d.gy) { /* left, right */
                bc_set(T, d, kind, act[lo[0]], val[lo[0]], 0, i, 0, 0, 0);
                bc_set(T, d, kind, act[hi[0]], val[hi[0]], 0, i, 0, 0, 1);
            }
        }
    } else {
        int n = d.gx > d.gy ? d.gx : d.gy;
        n = n > d.gz ? n : d.gz;
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) {
                if (i < d.gx && j < d.gy) { /* bot (k=1), top (k=end) */
                    bc_set(T, d, kind, act[lo[2]], val[lo[2]], i, j, 0, 2, 0);
                    bc_set(T, d, kind, act[hi[2]], val[hi[2]], i, j, 0, 2, 1);
                }
                if (i < d.gy && j < d.gz) { /* left, right: T[1, i, j] */
                    bc_set(T, d, kind, act[lo[0]], val[lo[0]], 0, i, j, 0, 0);
                    bc_set(T, d, kind, act[hi[0]], val[hi[0]], 0, i, j, 0, 1);
                }
                if (i < d.gx && j < d.gz) { /* front, back: T[i, 1, j] */
                    bc_set(T, d, kind, act[lo[1]], val[lo[1]], i, 0, j, 1, 0);
                    bc_set(T, d, kind, act[hi[1]], val[hi[1]], i, 0, j, 1, 1);
                }
            }
    }
}
static int any6(const int32_t *b) { return b[0] | b[1] | b[2] | b[3] | b[4] | b[5]; }
void orc_thermal_bcs(const orc_thermal_fields *f, const orc_thermal_opts *o)
{
    /* do_bc(bc) = any(!=(false), values(bc))  BoundaryConditions.jl:19 */
    if (any6(o->cv_active)) bc_sweep(f, o, BC_CV);
    if (any6(o->no_flux)) bc_sweep(f, o, BC_NOFLUX);
    if (any6(o->periodic)) bc_sweep(f, o, BC_PERIODIC);
}

/* adiabatic_heating  DiffusionPT_kernels.jl:720-731:  A = (P − P0)·α·_dt */
void orc_thermal_adiabatic(const orc_thermal_fields *f, const orc_thermal_opts *o, const double *P, const double *P0)
{
    const dims_t d = mkdims(f);
    const size_t nc = (size_t)d.nx * d.ny * d.nz;
    const double _dt = 1.0 / o->dt;
    for (size_t c = 0; c < nc; c++) f->adiabatic[c] = (P[c] - P0[c]) * ratio_alpha(o, f->phase_c, nc, c) * _dt;
}

void orc_thermal_iterate_once(const orc_thermal_fields *f, const orc_thermal_opts *o)
{
    if (o->form == 1 && f->phase_c) orc_thermal_pt_arrays(f, o); /* update_pt_thermal_arrays!  solver.jl:233-234 */
    orc_thermal_flux(f, o);
    orc_thermal_update_T(f, o);
    orc_thermal_bcs(f, o);
    /* update_halo!(thermal.T): the caller exchanges between ranks */
}

static size_t ghosted_len(const dims_t d) { return (size_t)d.gx * d.gy * d.gz; }

/* _heatdiffusion_PT!  (single rank) */
int orc_heatdiffusion_PT(const orc_thermal_fields *f, const orc_thermal_opts *o, const double *stokes_P, const double *stokes_P0,
                         orc_thermal_result *res)
{
    const dims_t d = mkdims(f);
    const size_t nc = (size_t)d.nx * d.ny * d.nz, ng = ghosted_len(d);
    const double _sq_len_RT = 1.0 / sqrt((double)nc);
    memcpy(f->Told, f->T, ng * sizeof(double));
    if (o->form == 1 && stokes_P && stokes_P0) orc_thermal_adiabatic(f, o, stokes_P, stokes_P0);
    int64_t iter = 0, cont = 0;
    double err = 2 * o->eps;
    while (err > o->eps && iter < o->iterMax) {
        orc_thermal_iterate_once(f, o);
        iter += 1;
        if (iter % o->nout == 0) {
            orc_thermal_check_res(f, o);
            double s = 0.0;
            for (size_t c = 0; c < nc; c++) s += f->ResT[c] * f->ResT[c];
            err = sqrt(s) * _sq_len_RT;
            if (res && cont < res->cap) { res->norm_ResT[cont] = err; res->iter_count[cont] = iter; }
            cont += 1;
        }
    }
    for (size_t t = 0; t < ng; t++) f->dT[t] = f->T[t] - f->Told[t]; /* update_ΔT! */
    if (res) { res->iter = iter; res->nhist = cont; res->err = err; }
    return 0;
}
