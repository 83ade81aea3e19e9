/*
 * jrb200.h — C ABI of libjrb200.so, the B200 (sm_100a) backend for the
 * pseudo-transient (PT) hot path of PTsolvers/JustRelax.jl.
 *
 * This is the drop-in boundary: the entry points below are what a
 * `JustRelaxB200Ext` Julia package extension binds with `ccall` in place of
 * the ParallelStencil/CUDA.jl kernels that ext/JustRelaxCUDAExt.jl +
 * src/ext/CUDA/{2D,3D}.jl provide today (INTEGRATION.md shows the Julia side).
 * Plain C: POD structs, raw device pointers, sizes; no torch / C++ types.
 *
 * Conventions
 *  - Every array is Float64, dense, column-major (x fastest) — the memory
 *    layout of the Julia arrays in StokesArrays / ThermalArrays
 *    (src/types/constructors/stokes.jl, src/types/constructors/heat_diffusion.jl).
 *  - All pointers in jr_fields / jr_thermal_fields are DEVICE pointers on the
 *    context's device unless a function name says `_host`.
 *  - Functions return 0 (JR_OK) or a negative jr_status; jr_last_error() gives
 *    the message.  There is no CPU fallback anywhere in this library.
 *  - One context per GPU / host thread; calls on a context are blocking unless
 *    stated otherwise and are not re-entrant.
 */
#ifndef JRB200_H
#define JRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JRB200_ABI_VERSION 3

typedef enum {
    JR_OK = 0,
    JR_ERR_CUDA = -1,          /* CUDA runtime error (message has the cudaError string)          */
    JR_ERR_SHAPE = -2,         /* inconsistent sizes / null required field                        */
    JR_ERR_NAN = -3,           /* residual became NaN — reference: error("NaN(s)") Stokes3D.jl:162 */
    JR_ERR_UNSUPPORTED = -4,   /* rheology / BC / option outside the supported subset             */
    JR_ERR_NCCL = -5,
    JR_ERR_ARG = -6
} jr_status;

/* ------------------------------------------------------------------------- *
 * Field slots of a StokesArrays object — replaces the struct-of-CuArrays the
 * reference passes to its kernels (src/types/stokes.jl:161-183; shapes
 * src/types/constructors/stokes.jl:10-302).  Unused slots are NULL.
 * Naming: t=τ, e=ε, p=ε_pl, d=Δε, w=ω, `_o` = τ_o, `_c` = shear @ centres,
 * `_v` = normal @ vertices (2D), lam=λ, lamv=λv, etatau=ητ, divV=∇V, divU=∇U,
 * rhog*=ρg tuple, K/G = bulk/shear modulus arrays (VA/V2 variants),
 * T/Pargs = args.T (ghosted, ni.+2) and args.P (ni); dTargs = args.ΔT (ni): when
 * non-NULL the VC solves use the thermal-stress form of compute_P!
 * (src/stokes/PressureKernels.jl:128-149,197-206).
 * ------------------------------------------------------------------------- */
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

typedef enum {
#define X(n) JR_F_##n,
    JR_STOKES_FIELDS(X)
#undef X
    JR_F_COUNT
} jr_field;

typedef struct {
    int32_t ndim;            /* 2 or 3                                  */
    int32_t n[3];            /* local cells nx, ny, nz (nz = 1 in 2D)   */
    double *f[JR_F_COUNT];   /* device pointers                         */
} jr_fields;

/* PTStokesCoeffs (src/types/stokes.jl:203-229), grid spacing (uniform
 * Geometry, src/grid/Cartesian.jl:42-58), VelocityBoundaryConditions flags
 * (src/boundaryconditions/types.jl:110-157) and the `kwargs` NamedTuple of
 * solve! (src/stokes/Stokes3D.jl:35-40, 458-465; Stokes2D.jl:588-598). */
typedef struct {
    double r, theta_dtau, eta_dtau, eps_rel, eps_abs;
    double _di[3];
    double dt;
    int64_t iterMax, nout;
    int32_t n_g[3];                 /* nx_g(), ny_g(), nz_g()                         */
    int32_t free_slip[6], no_slip[6], periodic[6]; /* left,right,front,back,top,bot   */
    double viscosity_relaxation, lambda_relaxation, visc_cutoff_lo, visc_cutoff_hi;
    int64_t iterMin;
    int32_t strain_rate_ni_only;
    int32_t strain_increment;       /* 2D-VC kwarg strain_increment = true: Δε form (Stokes2D.jl:659-730, StressKernels.jl:1147-1302) */
    int32_t displacement_bcs;       /* flow_bcs isa DisplacementBoundaryConditions (src/types/displacement.jl:62-70): V = U/dt before the
                                     * loop, flow_bcs! applied to U instead of V (2D-VC) */
    int32_t dT_ghosted;             /* args.ΔT has the extents ni.+2 (thermal.ΔT, as the reference's scripts pass it: test_thermalstresses.jl:321,
                                     * Blob3D.jl:280) and is indexed ΔT[I...] WITHOUT an offset, exactly like compute_P_kernel! does
                                     * (PressureKernels.jl:143-146: quirk — the low corner of the ghosted array); 0: extents ni */
} jr_stokes_opts;

typedef struct {
    int64_t iter;
    int64_t nhist;
    double err;
    double *err_evo1; int64_t *err_evo2;         /* HOST arrays, capacity iterMax/nout + 2 */
    double *norm_Rx, *norm_Ry, *norm_Rz, *norm_divV;
    double time_s;               /* device time of the PT loop (CUDA events), = `time`     */
    int64_t kernel_launches;     /* kernels launched by this call                          */
} jr_stokes_result;

typedef struct jr_context jr_context;

/* flags for jr_context_set_flags */
#define JR_FLAG_UNFUSED   1u  /* run the reference-structured one-kernel-per-@parallel path */
#define JR_FLAG_DIAG_EVERY_ITER 2u /* fused path: write ∇V, ε, R, U every iteration (default: only when read) */

const char *jr_last_error(void);
int jr_abi_version(void);
int jr_field_count(void);
const char *jr_field_name(int i);

/* device / context management.  `stream` is a cudaStream_t (0 = create an own stream). */
int jr_context_create(int device, void *stream, jr_context **out);
int jr_context_destroy(jr_context *ctx);
int jr_context_set_flags(jr_context *ctx, uint32_t flags);
int jr_context_synchronize(jr_context *ctx);

/* memory helpers — what PTArray(B200Backend) / Array(::B200Array) bind
 * (ext/JustRelaxCUDAExt.jl:7-10, src/types/type_conversions.jl:20-63). */
int jr_malloc(jr_context *ctx, size_t bytes, void **dptr);
int jr_free(jr_context *ctx, void *dptr);
int jr_memcpy_h2d(jr_context *ctx, void *dst, const void *src_host, size_t bytes);
int jr_memcpy_d2h(jr_context *ctx, void *dst_host, const void *src, size_t bytes);
int jr_memcpy_d2d(jr_context *ctx, void *dst, const void *src, size_t bytes);
int jr_fill_f64(jr_context *ctx, double *dptr, double value, size_t count);

/* --- 3D Stokes, visco-elastic variant with K,G arrays ----------------------
 * replaces JR3D.solve!(::CUDABackendTrait, stokes, pt_stokes, grid, flow_bcs, ρg, K, G, dt, igg; kwargs)
 * (src/ext/CUDA/3D.jl:375-377 → src/stokes/Stokes3D.jl:25-186). */
int jr_stokes3d_solve_VA(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, jr_stokes_result *res);
/* exactly `niter` PT iterations, no convergence test (benchmark / fixed-iteration parity). */
int jr_stokes3d_iterate_VA(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int64_t niter,
                           jr_stokes_result *res);

/* iteration session: what the body of the reference's `while` loop (src/stokes/Stokes3D.jl:76-122) is to a host that
 * keeps its own convergence logic.  begin = the pre-loop work of _solve! (:43-57: ητ = maxloc(η) + halo) and the entry
 * into the library's TMA box layout; step = `niter` PT iterations (res->time_s = device time of exactly these
 * iterations, CUDA events on the context's stream); with observe_last the last iteration also writes everything the
 * reference's iteration leaves in the user's arrays (∇V, ε, R, RP, U and the state V, P, τ), otherwise the dense arrays
 * are stale until the next observable iteration or end; end = leave the layout (the dense state is made current).
 * One open session per context; fields/opts are copied at begin, the arrays must stay allocated until end. */
int jr_stokes3d_VA_begin(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o);
int jr_stokes3d_VA_step(jr_context *ctx, int64_t niter, int observe_last, jr_stokes_result *res);
int jr_stokes3d_VA_end(jr_context *ctx);

/* facts about the fused plan the last VA solve on this context used:
 * info = {tile rows BY, z-chunks, 1 if the body-force arrays were constant and not streamed, finite dt,
 *         box pitch PX, PY, PZ, lock-step slack}.  No reference counterpart (reporting only). */
int jr_stokes3d_VA_plan_info(jr_context *ctx, int32_t info[8]);

/* --- multiphase visco-elasto-plastic (VC) inputs ---------------------------------------------------------------
 * One row of the flat Stokes rheology table, lowered from rheology::NTuple{N,MaterialParams} once per solve
 * (SURVEY.md Appendix C): LinearViscous η; ConstantElasticity G, Kb (Inf when absent / NaN / 0:
 * src/rheology/GeoParams.jl:1-15); the FIRST DruckerPrager[_regularised] element (src/rheology/StressUpdate.jl:131-144);
 * Constant/PT_/T_Density.  Anything else must be rejected at lowering time (JR_ERR_UNSUPPORTED, no fallback). */
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
} jr_stokes_phase;

/* rheology table (HOST pointer, nphase <= 8 rows), gravity of phase 1 (src/rheology/BuoyancyForces.jl:25,56),
 * JustPIC PhaseRatios flattened to [phase][node] DEVICE arrays: centres + vertices (2D) / centres + xy, yz, xz edges (3D) */
typedef struct {
    int32_t nphase, g_scalar;          /* g_scalar: compute_gravity returned a Number → only the last ρg component is filled */
    const jr_stokes_phase *phases;
    double g[3];
    const double *ph_center, *ph_vertex, *ph_xy, *ph_yz, *ph_xz;
    double free_surface;               /* dt * free_surface of compute_V!/compute_Res! (2D, VelocityKernels.jl:134-180); 0 = off */
} jr_vc_inputs;

/* --- 2D Stokes ------------------------------------------------------------------------------------------------
 * replaces JR2D.solve!(::CUDABackendTrait, stokes, pt_stokes, di, flow_bcs, ρg, G, K, dt, igg; kwargs)
 * (src/ext/CUDA/2D.jl → src/stokes/Stokes2D.jl:181-325, variant 2D-V2) and the multiphase VEP
 * solve!(stokes, pt_stokes, di, flow_bcs, ρg, phase_ratios, rheology, args, dt, igg; kwargs)
 * (src/stokes/Stokes2D.jl:577-866, variant 2D-VC).  One fused sm_100a kernel per PT iteration. */
int jr_stokes2d_solve_V2(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, jr_stokes_result *res);
int jr_stokes2d_iterate_V2(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int64_t niter, jr_stokes_result *res);
int jr_stokes2d_solve_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, jr_stokes_result *res);
/* pre-loop initialisation + exactly niter iterations (+ the exit kernels when finish != 0); λ, λv are exposed in the lam/lamv slots */
int jr_stokes2d_iterate_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, int64_t niter, int finish,
                           jr_stokes_result *res);
/* flow_bcs!(stokes, bcs) 2D  src/boundaryconditions/BoundaryConditions.jl:65-100 (flags use slots left,right,top,bot = 0,1,4,5) */
int jr_flow_bcs2d(jr_context *ctx, double *Ax, double *Ay, const int32_t n[3], const int32_t free_slip[6], const int32_t no_slip[6],
                  const int32_t periodic[6]);
/* compute_viscosity!(stokes, phase_ratios, args, rheology, cutoff; relaxation = nu)  src/rheology/Viscosity.jl:67-106,282-323 */
int jr_compute_viscosity2d(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, double nu);
/* compute_ρg!(ρg, phase_ratios, rheology, args)  src/rheology/BuoyancyForces.jl:74-95 */
int jr_compute_rhog2d(jr_context *ctx, const jr_fields *s, const jr_vc_inputs *vc);
/* tensor_invariant!(II, xx, yy, xy_vertex)  src/stokes/StressKernels.jl:470-480 */
int jr_tensor_invariant2d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *xy, const int32_t n[3]);

/* --- 3D multiphase visco-elasto-plastic Stokes (variant 3D-VC) -------------------------------------------------------
 * replaces JR3D.solve!(::CUDABackendTrait, stokes, pt_stokes, grid|di, flow_bcs, ρg, phase_ratios, rheology, args, dt, igg; kwargs)
 * (src/ext/CUDA/3D.jl:375-377 → src/stokes/Stokes3D.jl:447-668).  Three sm_100a kernels per PT iteration, cut where the
 * reference exchanges halos (ητ | τyz, τxz, τxy | V); with a communicator attached the same kernels run with the peer-memory
 * halo updates in between.  args.T → slot T (ni.+2), args.P → slot Pargs (may alias P). */
int jr_stokes3d_solve_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, jr_stokes_result *res);
/* pre-loop initialisation + exactly niter iterations (+ the exit kernels when finish != 0); λ is exposed in the lam slot */
int jr_stokes3d_iterate_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, int64_t niter, int finish,
                           jr_stokes_result *res);
/* compute_viscosity!(stokes, phase_ratios, args, rheology, cutoff; relaxation = nu) 3D  src/ext/CUDA/3D.jl:231-263 → src/rheology/Viscosity.jl:282-323,454-504 */
int jr_compute_viscosity3d(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, double nu);
/* compute_ρg!(ρg, phase_ratios, rheology, args) 3D  src/ext/CUDA/3D.jl:287-299 → src/rheology/BuoyancyForces.jl:38-60 */
int jr_compute_rhog3d(jr_context *ctx, const jr_fields *s, const jr_vc_inputs *vc);
/* tensor_invariant!(A::SymmetricTensor) 3D  src/ext/CUDA/3D.jl:266-272 → src/stokes/StressKernels.jl:442-468,482-500 */
int jr_tensor_invariant3d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *zz, const double *yz, const double *xz,
                          const double *xy, const int32_t n[3]);
/* shear2center!(A::SymmetricTensor) 3D  src/ext/CUDA/3D.jl:319-327 → src/Interpolations.jl:313-323 */
int jr_shear2center3d(jr_context *ctx, double *yz_c, double *xz_c, double *xy_c, const double *yz, const double *xz, const double *xy, const int32_t n[3]);

/* update_phase_ratios_{2,3}D!(phase_ratios, phase_arrays, xci, xvi) — grid-based phases (config 5)  src/ext/CUDA/3D.jl:519-539 →
 * src/phases/PhaseRatios.jl:21-78.  phase_arrays: nphase DEVICE arrays (nx, ny[, nz]); xci/xvi: HOST coordinate vectors per dimension;
 * outputs: DEVICE ratio arrays laid out [phase][node] at centres, vertices, velocity faces and (3D) edge midpoints; NULL = skip. */
int jr_phase_ratios_from_arrays(jr_context *ctx, int32_t ndim, const int32_t n[3], int32_t nphase, const double *const *phase_arrays,
                                const double *const *xci_host, const double *const *xvi_host, double *center, double *vertex, double *Vx,
                                double *Vy, double *Vz, double *xy, double *yz, double *xz);

/* --- stand-alone kernels the reference exposes outside the loops ----------- */
/* accumulate_tensor!(II, A::SymmetricTensor, dt): II += second_invariant_staggered(A) * dt  src/ext/CUDA/3D.jl:274-279 → src/stokes/StressKernels.jl:364-408 */
int jr_accumulate_tensor2d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *xy, const int32_t n[3], double dt);
int jr_accumulate_tensor3d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *zz, const double *yz, const double *xz,
                           const double *xy, const int32_t n[3], double dt);
/* accumulate_vol!(EVol_pl, ε_vol_pl, dt): EVol_pl += dt * ε_vol_pl  src/ext/CUDA/3D.jl:281-284 → src/stokes/StressKernels.jl:422-438 */
int jr_accumulate_vol(jr_context *ctx, double *EVol_pl, const double *e_vol_pl, size_t count, double dt);
/* maximum(abs.(A)) (allreduce != 0: maximum_mpi over the communicator) — the reduction of compute_dt  src/ext/CUDA/3D.jl:388-390 → src/Utils.jl:492-519 */
int jr_absmax(jr_context *ctx, const double *A, size_t count, int allreduce, double *out_host);
/* flow_bcs!(stokes, bcs)  src/ext/CUDA/3D.jl:195-218 → BoundaryConditions.jl:65-100 */
int jr_flow_bcs3d(jr_context *ctx, double *Ax, double *Ay, double *Az, const int32_t n[3],
                  const int32_t free_slip[6], const int32_t no_slip[6], const int32_t periodic[6]);
/* compute_maxloc!(B, A; window)  src/Utils.jl:409-461 */
int jr_maxloc3d(jr_context *ctx, double *B, const double *A, const int32_t n[3], const int32_t window[3]);
/* velocity2displacement!/displacement2velocity!  src/ext/CUDA/3D.jl:358-372 */
int jr_scale_copy(jr_context *ctx, double *dst, const double *src, double factor, size_t count);
/* Σ A[2:end-1,…]^2 (interior != 0) or Σ A^2 — the local part of norm_mpi, src/Utils.jl:698-701 */
int jr_sumsq(jr_context *ctx, const double *A, const int32_t n[3], int interior, double *out_host);

/* --- per-time-step kernels between the Stokes and the thermal loops (a coupled time step stays on the GPU) ----------------
 * velocity2vertex!(Vx_v, Vy_v[, Vz_v], Vx, Vy[, Vz]) / velocity2center!(Vx_c, Vy_c[, Vz_c], Vx, Vy[, Vz])
 * (src/ext/CUDA/3D.jl:344-356 → src/Interpolations.jl:212-289): n = local cells, out_ext = size of the output arrays (the
 * reference launches over size(Vx_v) / size(Vx_c)). */
int jr_velocity2vertex(jr_context *ctx, int32_t ndim, const int32_t n[3], const int32_t out_ext[3], double *Vx_v, double *Vy_v, double *Vz_v,
                       const double *Vx, const double *Vy, const double *Vz);
int jr_velocity2center(jr_context *ctx, int32_t ndim, const int32_t n[3], const int32_t out_ext[3], double *Vx_c, double *Vy_c, double *Vz_c,
                       const double *Vx, const double *Vy, const double *Vz);
/* compute_lithostatic_pressure!(P, ρg, dz[, igg]) (src/Utils.jl:541-617): P[j] = Σ_{k>j} ρg[k] dz[k] + ρg[j] dz[j] / 2 along the last
 * dimension; n = size(P); dz_cells = DEVICE vector of cell heights or NULL (then the scalar dz); across_ranks != 0 = the four-argument
 * method: the weight of the cells held by the ranks stacked above is gathered over peer memory (replaces MPI.Allgather on the vertical
 * sub-communicator), ncell_vertical = the local cell count given to init_global_grid in the vertical direction (IGG's nxyz[N]).
 * Without across_ranks the call fails (like the reference) when the vertical direction is split across ranks. */
int jr_lithostatic_pressure(jr_context *ctx, int32_t ndim, const int32_t n[3], double *P, const double *rhog, double dz, const double *dz_cells,
                            int across_ranks, int32_t ncell_vertical);
/* compute_shear_heating!(thermal, stokes, [phase_ratios,] rheology, dt) (src/thermal_diffusion/ShearHeating.jl:14-72):
 * shear_heating = max(0, Χ τij (εij − εij_el)), εij_el = ½(τij − τij_o)/(G dt); Χ per phase in chi_host[nphase] (GeoParams
 * ConstantShearheating); vc->ph_center NULL = the single-MaterialParams method. */
int jr_compute_shear_heating(jr_context *ctx, const jr_fields *s, const jr_vc_inputs *vc, const double *chi_host, double dt, double *shear_heating);

/* --- thermal diffusion: heatdiffusion_PT! ------------------------------------------------------------------------
 * replaces JR{2,3}D.heatdiffusion_PT!(::CUDABackendTrait, thermal, args...; kwargs) (src/ext/CUDA/3D.jl:383-385 →
 * src/thermal_diffusion/DiffusionPT_solver.jl:34-149 [K, ρCp arrays] and :181-305 [rheology]).
 * jr_thermal_fields = ThermalArrays (src/types/heat_diffusion.jl:1-16) + the arrays of PTThermalCoeffs (:30-44) +
 * the per-solve inputs.  T, Told, dT and the Dirichlet mask/value carry one ghost layer ((n+2)^d); q* are face
 * arrays; everything else is (n)^d.  Unused pointers are NULL. */
typedef struct {
    int32_t ndim;
    int32_t n[3];
    double *T, *Told, *dT;
    double *qTx, *qTy, *qTz, *qTx2, *qTy2, *qTz2;
    double *H, *shear_heating, *adiabatic, *ResT;
    double *theta_r_dtau, *dtau_rho;               /* pt_thermal.θr_dτ, pt_thermal.dτ_ρ                       */
    double *K, *rhoCp;                             /* array form                                              */
    double *P;                                     /* args.P (rheology form); args.T is T itself               */
    double *dir_mask, *dir_value;                  /* Dirichlet mask / value on the ghosted grid, or NULL      */
    double *phase_c, *phase_x, *phase_y, *phase_z; /* phase ratios [phase][node]: centre, Vx, Vy, Vz; or NULL  */
} jr_thermal_fields;

/* one row of the flat thermal rheology table (GeoParams subset: Constant/PT_/T_Density, ConstantHeatCapacity,
 * ConstantConductivity / TP_Conductivity, ConstantRadioactiveHeat) — lowered from rheology::NTuple{N,MaterialParams} once per solve.
 * TP_Conductivity (miniapps/convection/Particles3D/Layered_rheology.jl:45-57, miniapps/benchmarks/stokes2D/shear_heating/
 * Shearheating_rheology.jl:9-17): k(T, P) = (k_a + k_b / (T + k_c)) · (1 + k_d · P), evaluated where the reference evaluates
 * compute_conductivity: at the faces with T = the mean of the two adjacent nodes and P of the (clamped) cell on either side
 * (DiffusionPT_kernels.jl:93-100, 391-402), at the centres for the PT coefficients (DiffusionPT_coefficients.jl:122-135) */
typedef struct {
    int32_t rho_kind; /* 0 ConstantDensity, 1 PT_Density, 2 T_Density */
    int32_t has_Hr;
    double rho0, alpha, beta, T0, P0, Cp, k, Hr;
    int32_t k_kind;   /* 0 ConstantConductivity (k), 1 TP_Conductivity (k_a … k_d) */
    int32_t _pad;
    double k_a, k_b, k_c, k_d;
} jr_thermal_phase;

typedef struct {
    double _di[3], dt, eps;                        /* grid._di, dt, pt_thermal.ϵ                               */
    int64_t iterMax, nout;                         /* kwargs                                                   */
    double max_lxyz, Vpdtau;                       /* pt_thermal.max_lxyz, pt_thermal.Vpdτ                     */
    int32_t form;                                  /* 0: K, ρCp arrays; 1: rheology table                      */
    int32_t nphase;
    const jr_thermal_phase *phases;                /* HOST pointer, nphase rows                                */
    double dir_const;                              /* ConstantDirichletBoundaryCondition value (dir_value NULL) */
    /* TemperatureBoundaryConditions (src/boundaryconditions/types.jl:65-99), faces left,right,front,back,top,bot;
     * cv_active/cf_active: the entry is not `false` / is a number */
    int32_t no_flux[6], cv_active[6], cf_active[6], periodic[6];
    double cv_value[6], cf_value[6];
} jr_thermal_opts;

typedef struct {
    int64_t iter, nhist, cap;
    double err;
    double *norm_ResT; int64_t *iter_count;        /* HOST arrays of capacity cap                              */
    double time_s; int64_t kernel_launches;
} jr_thermal_result;

int jr_heatdiffusion_PT(jr_context *ctx, const jr_thermal_fields *f, const jr_thermal_opts *o, const double *stokes_P,
                        const double *stokes_P0, jr_thermal_result *res);
/* exactly niter PT iterations (no convergence test; Told is NOT reset) — fixed-iteration parity / benchmark */
int jr_thermal_iterate(jr_context *ctx, const jr_thermal_fields *f, const jr_thermal_opts *o, int64_t niter, jr_thermal_result *res);
/* thermal_bcs!(thermal, bcs)  src/ext/CUDA/3D.jl:220-226 → BoundaryConditions.jl:39-54 */
int jr_thermal_bcs(jr_context *ctx, double *T, int32_t ndim, const int32_t n[3], const jr_thermal_opts *o);
/* update_thermal_coeffs! / compute_pt_thermal_arrays!  src/ext/CUDA/3D.jl:110-175 → DiffusionPT_coefficients.jl:105-208 */
int jr_thermal_pt_arrays(jr_context *ctx, const jr_thermal_fields *f, const jr_thermal_opts *o);

/* --- multi-GPU: ImplicitGlobalGrid-compatible decomposition over CUDA-IPC peer memory (NVLink) -----------------
 * Replaces IGG's update_halo! (call sites src/stokes/Stokes3D.jl:57,120,515,578-580,596; Stokes2D.jl:655,757,784;
 * src/thermal_diffusion/DiffusionPT_solver.jl:110,261) and the MPI.Allreduce of norm_mpi/maximum_mpi
 * (src/Utils.jl:688-730).  One process per GPU; `dims`/`coords` are IGG's Cartesian topology (src/grid/Grid.jl:18-24).
 * Bootstrap needs ONE host collective, supplied by the caller: an all-gather of `bytes_per_rank` bytes per rank
 * (MPI.Allgather on igg.comm_cart in Julia; torch.distributed in the Python twin).  Return 0 on success. */
typedef int (*jr_allgather_fn)(const void *sendbuf, void *recvbuf, size_t bytes_per_rank, void *user);
typedef struct jr_comm jr_comm;
int jr_comm_create(jr_context *ctx, int rank, int nranks, const int32_t dims[3], const int32_t coords[3],
                   jr_allgather_fn allgather, void *user, jr_comm **out);
/* the same with IGG's periodx / periody / periodz (init_global_grid keywords; test/test_periodic_boundary_conditions_MPI.jl:12-19,
 * src/grid/Utils.jl:29-83): the grid of ranks wraps around in a periodic dimension — the first rank's low ghost planes come from the
 * last rank and vice versa; with a single rank in that dimension the rank exchanges with itself.  periods == NULL: none. */
int jr_comm_create_periodic(jr_context *ctx, int rank, int nranks, const int32_t dims[3], const int32_t coords[3], const int32_t periods[3],
                            jr_allgather_fn allgather, void *user, jr_comm **out);
int jr_comm_destroy(jr_comm *comm);
/* solves / halo updates / reductions on `ctx` go through `comm` from now on (NULL detaches) */
int jr_context_set_comm(jr_context *ctx, jr_comm *comm);
int jr_comm_barrier(jr_context *ctx);
/* update_halo!(A1, A2, ...): dense arrays, extents[3*q..3*q+2] = size(Aq), ncell = local (nx,ny,nz); blocking */
int jr_update_halo3d(jr_context *ctx, int narrays, double *const *arrays, const int32_t *extents, const int32_t ncell[3]);
/* MPI.Allreduce of n <= 16 host doubles, op 0 = sum, 1 = max, 2 = min; rank-order deterministic; blocking */
int jr_allreduce_f64(jr_context *ctx, double *vals_host, int n, int op);
/* host-only index arithmetic of the exchange (no GPU needed): after update_halo!, element idx (0-based) of an array
 * of extents `ext` on rank `coords` holds the pre-exchange value of element src_idx on rank src_coords.
 * Returns 1 if the element is overwritten by the exchange, 0 if not, < 0 on error. */
int jr_halo_source(const int32_t dims[3], const int32_t coords[3], const int32_t ext[3], const int32_t ncell[3],
                   const int32_t idx[3], int32_t src_coords[3], int32_t src_idx[3]);

int jr_halo_source_periodic(const int32_t dims[3], const int32_t periods[3], const int32_t coords[3], const int32_t ext[3],
                            const int32_t ncell[3], const int32_t idx[3], int32_t src_coords[3], int32_t src_idx[3]);

#ifdef __cplusplus
}
#endif
#endif /* JRB200_H */
