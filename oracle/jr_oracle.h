/*
 * jr_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C Float64 restatement of the JustRelax.jl pseudo-transient hot path
 * (reference tree /root/reference, v0.7.1).  One C function per reference
 * `@parallel` kernel, same kernel split, same temporaries, same arithmetic
 * order (fma where the reference writes fma/muladd; compiled with
 * -ffp-contract=off so nothing else is contracted).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this library.  The product path
 * (justrelax_jl_b200 + libjrb200.so) never links or calls it.
 *
 * Parity status: the Julia reference cannot be executed in this environment
 * (no julia binary, no network).  The oracle is pinned against every
 * known-answer value the reference's own test-suite holds for this path
 * (tests/test_oracle_*.py list them with file:line); GeoParams.jl formulas
 * (third-party, compat 0.7.19, not vendored) are restated from its published
 * definitions and pinned only through the reference's integration goldens.
 *
 * Layout: dense column-major (x fastest) exactly like Julia Arrays; all
 * index macros below are 1-based so each line can be compared with the
 * reference source it cites.
 */
#ifndef JR_ORACLE_H
#define JR_ORACLE_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field slots of a StokesArrays object (src/types/stokes.jl:161-183 and
 * src/types/constructors/stokes.jl:10-302).  Same order as
 * include/jrb200.h (checked by tests/test_abi.py). */
#define JR_STOKES_FIELDS(X)                                                    \
    X(P) X(P0) X(divV) X(Q)                                                    \
    X(Vx) X(Vy) X(Vz) X(Ux) X(Uy) X(Uz)                                        \
    X(txx) X(tyy) X(tzz) X(tyz) X(txz) X(txy) X(tyz_c) X(txz_c) X(txy_c) X(tII) \
    X(txx_o) X(tyy_o) X(tzz_o) X(tyz_o) X(txz_o) X(txy_o)                      \
    X(tyz_o_c) X(txz_o_c) X(txy_o_c) X(tII_o)                                  \
    X(exx) X(eyy) X(ezz) X(eyz) X(exz) X(exy) X(eyz_c) X(exz_c) X(exy_c) X(eII) \
    X(pxx) X(pyy) X(pzz) X(pyz) X(pxz) X(pxy) X(pyz_c) X(pxz_c) X(pxy_c) X(pII) \
    X(dxx) X(dyy) X(dzz) X(dyz) X(dxz) X(dxy) X(dyz_c) X(dxz_c) X(dxy_c) X(dII) \
    X(EII_pl) X(EVol_pl) X(e_vol_pl)                                           \
    X(eta) X(etav) X(eta_vep) X(etatau)                                        \
    X(Rx) X(Ry) X(Rz) X(RP)                                                    \
    X(wyz) X(wxz) X(wxy)                                                       \
    X(divU) X(lam) X(lamv) X(dPpsi)                                            \
    X(rhogx) X(rhogy) X(rhogz)                                                 \
    X(K) X(G) X(T) X(Pargs)                                                    \
    X(txx_v) X(tyy_v) X(txx_o_v) X(tyy_o_v)                                    \
    X(dTargs)

enum orc_field {
#define X(n) ORC_F_##n,
    JR_STOKES_FIELDS(X)
#undef X
    ORC_F_COUNT
};

typedef struct {
    int32_t ndim;          /* 2 or 3 */
    int32_t n[3];          /* cells nx, ny, nz (nz = 1 in 2D) */
    double *f[ORC_F_COUNT];
} orc_fields;

/* PTStokesCoeffs (src/types/stokes.jl:203-229) + grid + loop control */
typedef struct {
    double r, theta_dtau, eta_dtau, eps_rel, eps_abs;
    double _di[3];         /* inverse spacing (uniform grid) */
    double dt;
    int64_t iterMax, nout;
    int32_t n_g[3];        /* nx_g(), ny_g(), nz_g() (IGG global sizes) */
    /* VelocityBoundaryConditions flags, order: left,right,front,back,top,bot
     * (3D) or left,right,top,bot (2D uses slots 0,1,4,5). */
    int32_t free_slip[6], no_slip[6], periodic[6];
    /* VC options */
    double viscosity_relaxation, lambda_relaxation, visc_cutoff_lo, visc_cutoff_hi;
    int64_t iterMin;
    int32_t strain_rate_ni_only;  /* Q20: 3D-VC launches strain rate over ni */
    int32_t strain_increment;     /* 2D-VC kwarg strain_increment: Δε form (Stokes2D.jl:659-730, StressKernels.jl:1147-1302) */
    int32_t displacement_bcs;     /* flow_bcs isa DisplacementBoundaryConditions: V = U/dt before the loop, BCs applied to U */
    int32_t dT_ghosted;           /* args.ΔT is (ni.+2) and indexed ΔT[I...] without offset, as compute_P_kernel! does (quirk) */
} orc_stokes_opts;

typedef struct {
    int64_t iter;
    int64_t nhist;
    double err;
    /* history arrays are provided by the caller, capacity iterMax/nout + 2 */
    double *err_evo1; int64_t *err_evo2;
    double *norm_Rx, *norm_Ry, *norm_Rz, *norm_divV;
} orc_stokes_result;

/* flat per-phase rheology row of the Stokes VC variants (lowered GeoParams MaterialParams):
 * LinearViscous η, ConstantElasticity (G, Kb; Inf when absent / NaN / 0: rheology/GeoParams.jl:1-15),
 * DruckerPrager[_regularised] (first plastic element wins, StressUpdate.jl:131-144), density law, expansivity. */
typedef struct {
    double eta;
    double G, Kb;
    int32_t has_pl, rho_kind;          /* rho_kind: 0 ConstantDensity, 1 PT_Density, 2 T_Density */
    double C, sinphi, cosphi, sinpsi, eta_vp;
    double rho0, alpha, beta, T0, P0;
    /* cohesion softening with the accumulated plastic strain EII (StressUpdate.jl:305-332): 0 none; 1 LinearSoftening
     * {lo, hi, max, min, slope, ordinate}; 2 NonLinearSoftening {ξ₀, Δ, μ, σ}: C(EII) = ξ₀ − ½ Δ erfc(−(EII − μ)/σ).  2D solves only. */
    int32_t soft_C_kind, _pad;
    double soft_C[6];
} orc_stokes_phase;

/* extra inputs of the multiphase (VC) solves: rheology table, gravity of phase 1 (BuoyancyForces.jl:25,56),
 * phase ratios [phase][node] at centres and vertices (2D) / xy, yz, xz edges (3D) */
typedef struct {
    int32_t nphase, g_scalar;          /* g_scalar: compute_gravity returned a Number → only the last ρg component is filled */
    const orc_stokes_phase *phases;
    double g[3];
    const double *ph_center, *ph_vertex, *ph_xy, *ph_yz, *ph_xz;
    double free_surface;               /* dt * free_surface factor of compute_V!/compute_Res! (2D), 0 when off */
} orc_vc_inputs;

const char *orc_field_name(int i);
int orc_field_count(void);

/* ---- 3D kernels (one per reference @parallel launch) -------------------- */
void orc_compute_divV3(const orc_fields *s, const double _di[3]);
void orc_compute_P_VA(const orc_fields *s, const double *eta, double dt, double r, double theta_dtau);
void orc_compute_strain_rate3(const orc_fields *s, const double _di[3], int range_plus1);
void orc_compute_tau3_VE(const orc_fields *s, double dt, double theta_dtau);
void orc_compute_V3(const orc_fields *s, double eta_dtau, const double _di[3]);
void orc_velocity2displacement(const orc_fields *s, double dt);
void orc_free_slip3(double *Ax, double *Ay, double *Az, const int32_t n[3], const int32_t bc[6]);
void orc_no_slip3(double *Ax, double *Ay, double *Az, const int32_t n[3], const int32_t bc[6]);
void orc_periodic3(double *Ax, double *Ay, double *Az, const int32_t n[3], const int32_t bc[6]);
void orc_flow_bcs3(const orc_fields *s, const orc_stokes_opts *o, int displacement);
void orc_maxloc3(double *B, const double *A, int nx, int ny, int nz, int wx, int wy, int wz);
void orc_multi_copy_tau3(const orc_fields *s);
int  orc_solve3d_VA(const orc_fields *s, const orc_stokes_opts *o, orc_stokes_result *res);
/* run exactly `niter` PT iterations (no convergence test) – used for fixed-iteration parity */
int  orc_iterate3d_VA(const orc_fields *s, const orc_stokes_opts *o, int64_t niter);
/* loop pieces for multi-rank emulation: pre-loop maxloc, one PT iteration (no halo: the caller exchanges) */
void orc_pre3d_VA(const orc_fields *s);
void orc_iterate3d_VA_once(const orc_fields *s, const orc_stokes_opts *o);

/* ---- 2D (oracle/stokes2d.c) ------------------------------------------------ */
int  orc_solve2d_V2(const orc_fields *s, const orc_stokes_opts *o, orc_stokes_result *res);
int  orc_iterate2d_V2(const orc_fields *s, const orc_stokes_opts *o, int64_t niter);
int  orc_solve2d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, orc_stokes_result *res);
/* niter iterations of the VC loop incl. its pre-loop initialisation, then (if finish) the exit kernels */
int  orc_iterate2d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, int64_t niter, int finish);
void orc_flow_bcs2(const orc_fields *s, const orc_stokes_opts *o, int displacement);
void orc_viscosity2d(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, double nu);
void orc_rhog2d(const orc_fields *s, const orc_vc_inputs *vc);
void orc_tensor_invariant2d(double *II, const double *xx, const double *yy, const double *xy, int nx, int ny);

/* ---- 3D multiphase VEP (oracle/stokes3d_vc.c) ---------------------------------- */
int  orc_solve3d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, orc_stokes_result *res);
/* niter iterations of the 3D-VC loop incl. its pre-loop initialisation, then (if finish) the exit kernels */
int  orc_iterate3d_VC(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, int64_t niter, int finish);
void orc_rhog3d(const orc_fields *s, const orc_vc_inputs *vc);
void orc_viscosity3d(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, double nu);
/* stepwise form for multi-rank emulation: piece 0 = maxloc (then halo ητ), 1 = ∇V…stress (then halo τ shear), 2 = V + BCs (then halo V) */
void *orc_vc3_begin(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc);
void orc_vc3_step(const orc_fields *s, const orc_stokes_opts *o, const orc_vc_inputs *vc, void *h, int piece);
void orc_vc3_end(const orc_fields *s, const orc_stokes_opts *o, void *h, int finish);

/* ---- grid-based phase ratios (oracle/phase_ratios.c): update_phase_ratios_{2,3}D!  src/phases/PhaseRatios.jl:21-78 ---- */
void orc_phase_ratios_from_arrays(int nd, const int32_t n[3], int N, const double *const *ph, const double *const *xc, const double *const *xv,
                                  double *center, double *vertex, double *Vx, double *Vy, double *Vz, double *xy, double *yz, double *xz);

/* ---- per-time-step kernels around the loops (oracle/aux.c) ---- */
void orc_velocity2vertex(int nd, const int32_t n[3], const int32_t e[3], double *Xv, double *Yv, double *Zv, const double *Vx, const double *Vy, const double *Vz);
void orc_velocity2center(int nd, const int32_t n[3], const int32_t e[3], double *Xc, double *Yc, double *Zc, const double *Vx, const double *Vy, const double *Vz);
void orc_lithostatic_pressure(int nd, const int32_t n[3], double *P, const double *rhog, double dz, const double *dzv, const double *above);
void orc_shear_heating(const orc_fields *s, const orc_vc_inputs *vc, const double *chi, double dt, double *out);

/* norms (src/Utils.jl:698-701), interior slice 2:end-1 in every dim when interior!=0 */
double orc_sumsq_interior(const double *A, int n1, int n2, int n3, int interior);

/* ---- mini-kernel known-answer entry points (src/MiniKernels.jl) --------- */
double orc_mini3(const char *name, const double *A, int n1, int n2, int n3, double _d, int i, int j, int k);
double orc_mini2(const char *name, const double *A, int n1, int n2, double _d, int i, int j);

#ifdef __cplusplus
}
#endif

/* ---- 1-based accessors --------------------------------------------------- */
#define IX3(n1, n2, i, j, k) ((size_t)((k) - 1) * (size_t)(n2) * (size_t)(n1) + (size_t)((j) - 1) * (size_t)(n1) + (size_t)((i) - 1))
#define IX2(n1, i, j) ((size_t)((j) - 1) * (size_t)(n1) + (size_t)((i) - 1))

static inline int orc_clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline double orc_inv(double x) { return 1.0 / x; }

#endif
