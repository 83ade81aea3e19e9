// stokes3d_vc.cu — the 3D multiphase visco-elasto-plastic Stokes PT loop of libjrb200 (sm_100a), variant 3D-VC
// (src/stokes/Stokes3D.jl:447-668; config 5 of BASELINE.json: 3D convection with grid-based phases).
//
// The reference launches per iteration: compute_maxloc! → update_halo!(ητ) → compute_∇V! → compute_P! → compute_strain_rate! →
// update_ρg! → update_viscosity_τII! → update_stresses_center_vertex_ps! → update_halo!(τyz, τxz, τxy) → compute_V! →
// velocity2displacement! → flow_bcs! → update_halo!(V)   (≈ 100 + 4N array passes).  Here one iteration is three kernels,
// cut exactly where the reference exchanges halos, so the multi-GPU path runs the same kernels:
//   k_vc3_prep    per cell:  ητ = maxloc(η) | ∇V | θ ← compute_P!(ητ, K, G from phase ratios) | ε (centres and the three edge
//                            families, launched over `ni` only: quirk Q20) | ρg(T, P) | η ← relaxed phase viscosity (Q10, Q13)
//   k_vc3_stress  per node:  update_stresses_center_vertex_ps! (yz, xz, xy edges and the centre) as a race-free Jacobi step:
//                            τ is ping-ponged (set in → set out), the schedule the oracle declares canonical for the racy
//                            reference kernel (quirk Q7)
//   k_vc3_vel     per cell:  compute_V! (+ residuals on sampled iterations); then the in-place flow_bcs! kernels of bc.cu
// Diagnostics nobody reads inside the loop (∇V, RP, ε_pl, τII, η_vep, ε_vol_pl, R, U) are only stored on iterations whose result
// can be observed (every `nout`, and the last).  Arithmetic is operation for operation the reference's (fma only where it writes
// fma/muladd; -fmad=false).
#include "rheo.cuh"
#include "comm.cuh"
#include "tma.cuh"

#define F(name) (s->f[JR_F_##name])

struct V3 {
    int nx, ny, nz;
    double _dx, _dy, _dz, dt, r, th, edt, rel, nu, cut_lo, cut_hi;
    double *Vx, *Vy, *Vz, *theta, *P;
    const double *P0, *Q, *eta_i;
    double *eta_o, *etatau;
    double *exx, *eyy, *ezz, *eyz, *exz, *exy;
    const double *txx_i, *tyy_i, *tzz_i, *tyz_i, *txz_i, *txy_i;
    double *txx_o, *tyy_o, *tzz_o, *tyz_o, *txz_o, *txy_o;
    double *tyzc, *txzc, *txyc;
    const double *oxx, *oyy, *ozz, *oyz, *oxz, *oxy, *oyzc, *oxzc, *oxyc;   // τ_o
    double *lam, *lamyz, *lamxz, *lamxy;
    double *rgx, *rgy, *rgz;
    const double *T, *Pargs, *dTargs, *ph_c, *ph_xy, *ph_yz, *ph_xz;
    double *divV, *RP, *pxx, *pyy, *pzz, *pyz, *pxz, *pxy, *tII, *eta_vep, *e_vol_pl, *Rx, *Ry, *Rz;
    int dT_ghosted;   // args.ΔT is (ni.+2), indexed ΔT[i, j, k] without offset (the reference's compute_P_kernel!)
    int pf_next;   // L2 prefetch of the next plane's operands (JRB200_VC3_PREFETCH, default on)
    int bo_x, bo_y, bo_z;             // block offsets of a stress launch that covers a slab of the node lattice only (the rim of the z-marching region)
    int zm_i, zm_j, zm_k, zm_chunk;   // nodes i ≤ zm_i, j ≤ zm_j, k ≤ zm_k belong to the z-marching stress kernel (0: none); planes per CTA
    int xfull_c, xfull_n;   // block columns with full 32-wide tiles for the cell (nx) and node (nx+1) lattices; a further block
                            // column, if launched, packs the few remainder columns densely (nx = 257: 1 cell / 2 node columns)
};
// thread → 1-based (i, j): block columns < xfull are ordinary 32 × 8 tiles, block column xfull packs the n0 − 32·xfull remainder columns
__device__ __forceinline__ void vc3_map_ij(int n0, int xfull, int &i, int &j)
{
    if ((int)blockIdx.x < xfull) {
        i = blockIdx.x * 32 + threadIdx.x + 1; j = blockIdx.y * 8 + threadIdx.y + 1;
    } else {
        const int rem = n0 - xfull * 32, lin = blockIdx.y * 256 + threadIdx.y * 32 + threadIdx.x, jj = lin / rem;
        i = xfull * 32 + (lin - jj * rem) + 1; j = jj + 1;
    }
}
__device__ __forceinline__ void vc3_map_ij_b(int n0, int xfull, int bx, int by, int &i, int &j)
{
    if (bx < xfull) {
        i = bx * 32 + threadIdx.x + 1; j = by * 8 + threadIdx.y + 1;
    } else {
        const int rem = n0 - xfull * 32, lin = by * 256 + threadIdx.y * 32 + threadIdx.x, jj = lin / rem;
        i = xfull * 32 + (lin - jj * rem) + 1; j = jj + 1;
    }
}
__device__ __forceinline__ void jr_prefetch_l2(const double *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#define CC(A, i, j, k) (A)[IX3(nx, ny, i, j, k)]
#define YZ(A, i, j, k) (A)[IX3(nx, ny + 1, i, j, k)]
#define XZ(A, i, j, k) (A)[IX3(nx + 1, ny, i, j, k)]
#define XY(A, i, j, k) (A)[IX3(nx + 1, ny + 1, i, j, k)]
#define VX(i, j, k) a.Vx[IX3(nx + 1, ny + 2, i, j, k)]
#define VY(i, j, k) a.Vy[IX3(nx + 2, ny + 1, i, j, k)]
#define VZ(i, j, k) a.Vz[IX3(nx + 2, ny + 2, i, j, k)]

// ---------------------------------------------------------------------------------------------------------------------------------
// MAXLOC: compute ητ here (single rank); otherwise ητ was computed by k_maxloc3 and halo-exchanged before this launch.
// NP = compile-time bound on the number of phases: the centre ratios are loaded once and reused for K, G, ρ and η.
#ifndef JR_PREP_MINB
#define JR_PREP_MINB 4
#endif
template <bool DIAG, bool MAXLOC, int NP, bool DTF = false>
__global__ void __launch_bounds__(256, JR_PREP_MINB) k_vc3_prep(const __grid_constant__ V3 a, const __grid_constant__ jr_phase_tab pt)
{
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    int i, j;
    vc3_map_ij(nx, a.xfull_c, i, j);
    const int k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const size_t nc = (size_t)nx * ny * nz, c = IX3(nx, ny, i, j, k);
    if (a.pf_next && k + 2 <= nz) {
        // L2 prefetch of this thread's operands two planes up (the CTAs of that plane start about one wave later)
        const size_t c2 = c + 2 * (size_t)nx * ny;
        jr_prefetch_l2(a.eta_i + c2); jr_prefetch_l2(a.theta + c2); jr_prefetch_l2(a.P0 + c2); jr_prefetch_l2(a.Q + c2);
        jr_prefetch_l2(a.Vx + IX3(nx + 1, ny + 2, i, j + 1, k + 3)); jr_prefetch_l2(a.Vy + IX3(nx + 2, ny + 1, i + 1, j, k + 3));
        jr_prefetch_l2(a.Vz + IX3(nx + 2, ny + 2, i + 1, j + 1, k + 3));
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (p < pt.n) jr_prefetch_l2(a.ph_c + (size_t)p * nc + c2);
        if (!pt.rho_const) {
            if (a.T) jr_prefetch_l2(a.T + IX3(nx + 2, ny + 2, i + 1, j + 1, k + 3));
            if (a.Pargs) jr_prefetch_l2(a.Pargs + c2);
        }
    }
    const double eta = __ldg(a.eta_i + c);
    // every operand is requested before the first store (V, P0, Q, T, P are read-only here: ld.global.nc lets the compiler hoist them;
    // θ is read-modify-write by this thread only)
    const double th_in = a.theta[c], P0c = __ldg(a.P0 + c), Qc = __ldg(a.Q + c);
    const double Targ = (!pt.rho_const && a.T) ? __ldg(a.T + IX3(nx + 2, ny + 2, i + 1, j + 1, k + 1)) : 0.0;
    const double Parg = (!pt.rho_const && a.Pargs) ? __ldg(a.Pargs + c) : 0.0;
    double ett;
    if (MAXLOC) {  // compute_maxloc!(ητ, η)  Stokes3D.jl:514 ; Utils.jl:409-461 (window clamped to the array)
        const int sy = nx, sz = nx * ny;
        const int oi[3] = {jr_clamp(i - 1, 1, nx) - i, 0, jr_clamp(i + 1, 1, nx) - i};
        const int oj[3] = {(jr_clamp(j - 1, 1, ny) - j) * sy, 0, (jr_clamp(j + 1, 1, ny) - j) * sy};
        const int ok[3] = {(jr_clamp(k - 1, 1, nz) - k) * sz, 0, (jr_clamp(k + 1, 1, nz) - k) * sz};
        double x = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < 3; kk++)
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int ii = 0; ii < 3; ii++) {
                    const double e = __ldg(a.eta_i + c + (oi[ii] + oj[jj] + ok[kk]));
                    if (e > x) x = e;
                }
        ett = x;
        a.etatau[c] = ett;
    } else
        ett = a.etatau[c];
    // compute_∇V!  VelocityKernels.jl:3-6   (Vx (nx+1, ny+2, nz+2), Vy (nx+2, ny+1, nz+2), Vz (nx+2, ny+2, nz+1))
    const int xsy = nx + 1, xsz = (nx + 1) * (ny + 2), ysy = nx + 2, ysz = (nx + 2) * (ny + 1), zsy = nx + 2, zsz = (nx + 2) * (ny + 2);
    const double *__restrict__ pVx = a.Vx + IX3(nx + 1, ny + 2, i, j + 1, k + 1);
    const double *__restrict__ pVy = a.Vy + IX3(nx + 2, ny + 1, i + 1, j, k + 1);
    const double *__restrict__ pVz = a.Vz + IX3(nx + 2, ny + 2, i + 1, j + 1, k);
    const double vx0 = __ldg(pVx), vy0 = __ldg(pVy), vz0 = __ldg(pVz);
    const double dVx = (-vx0 + __ldg(pVx + 1)) * a._dx;
    const double dVy = (-vy0 + __ldg(pVy + ysy)) * a._dy;
    const double dVz = (-vz0 + __ldg(pVz + zsz)) * a._dz;
    const double vyB = __ldg(pVy - ysz), vzS = __ldg(pVz - zsy), vxB = __ldg(pVx - xsz), vzW = __ldg(pVz - 1), vxS = __ldg(pVx - xsy), vyW = __ldg(pVy - 1);
    const double divV = dVx + dVy + dVz;
    // phase ratios at the centre, loaded once
    double r[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) r[p] = p < pt.n ? __ldg(a.ph_c + (size_t)p * nc + c) : 0.0;
    // compute_P!(θ, P0, RP, ∇V, Q, ητ, rheology, phase_ratios, …)  Stokes3D.jl:518-531 ; PressureKernels.jl:87-102,186-195
    double Kc = 0.0, Gc = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++)
        if (p < pt.n) { Kc += (r[p] == 0.0) ? 0.0 : pt.Kb[p] * r[p]; Gc += (r[p] == 0.0) ? 0.0 : pt.G[p] * r[p]; }
    double RP, th = th_in;
    if (DTF) {  // args.ΔT given (instantiated separately): thermal-stress form, α = fn_ratio(get_thermal_expansion, …)  PressureKernels.jl:128-149
        double al = 0.0;
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (p < pt.n) al += (r[p] == 0.0) ? 0.0 : (pt.rho_kind[p] == 0 ? 0.0 : pt.alpha[p]) * r[p];
        jr_compute_P_point_dT(RP, th, P0c, divV, Qc, __ldg(a.dTargs + (a.dT_ghosted ? IX3(nx + 2, ny + 2, i, j, k) : c)), al, ett, Kc, Gc, a.dt, a.r, a.th);
    } else
        jr_compute_P_point(RP, th, P0c, divV, Qc, ett, Kc, Gc, a.dt, a.r, a.th);
    a.theta[c] = th;
    // compute_strain_rate! over ni (quirk Q20)  Stokes3D.jl:533-535 ; VelocityKernels.jl:59-104
    const double d3 = divV * jr_inv(3.0);
    a.exx[c] = dVx - d3;
    a.eyy[c] = dVy - d3;
    a.ezz[c] = dVz - d3;
    YZ(a.eyz, i, j, k) = 0.5 * (a._dz * (vy0 - vyB) + a._dy * (vz0 - vzS));
    XZ(a.exz, i, j, k) = 0.5 * (a._dz * (vx0 - vxB) + a._dx * (vz0 - vzW));
    XY(a.exy, i, j, k) = 0.5 * (a._dy * (vx0 - vxS) + a._dx * (vy0 - vyW));
    // update_ρg!  Stokes3D.jl:538 ; BuoyancyForces.jl:38-60 (args.T sampled at I+1, quirk Q17); fn_ratio with args: a ratio == 1 returns that phase
    if (!pt.rho_const) {
        const double Tc = Targ, Pc = Parg;
        double rho = 0.0;
        bool done = false;
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (p < pt.n && !done) {
                if (r[p] == 1.0) { rho = jr_density(pt, p, Tc, Pc) * r[p]; done = true; }
                else rho += (r[p] == 0.0) ? 0.0 : jr_density(pt, p, Tc, Pc) * r[p];
            }
        if (!pt.g_scalar) { a.rgx[c] = rho * pt.g[0]; a.rgy[c] = rho * pt.g[1]; }
        a.rgz[c] = rho * pt.g[2];
    }
    // update_viscosity_τII! BEFORE the stress kernel (quirk Q13)  Stokes3D.jl:541-548 ; Viscosity.jl:454-504,599-619
    double eph = 0.0;
    bool single = false;
#pragma unroll
    for (int p = 0; p < NP; p++)
        if (p < pt.n && !single && r[p] > 0.999) { eph = pt.eta_c[p]; single = true; }
    if (!single) {
        double e = 0.0;
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (p < pt.n && r[p] != 0.0) e += pt.ieta_c[p] * r[p];
        eph = jr_inv(e);
    }
    a.eta_o[c] = jr_clampd((1 - a.nu) * eta + a.nu * eph, a.cut_lo, a.cut_hi);
    if (DIAG) { a.divV[c] = divV; a.RP[c] = RP; }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// phase mixture at one staggered node: the ratios are loaded ONCE into registers and every phase-weighted quantity the stress kernel
// needs is formed from them with the reference's operation order (fn_ratio phases.jl:5-16, plastic_params_phase StressUpdate.jl:153-176,
// compute_yieldfunction_phase :384-452, compute_plastic_gradients_phase :463-550).  NP = compile-time bound on the number of phases.
template <int NP>
struct Mix {
    double r[NP];
    double G, Kb, eta_reg;
    bool is_pl;
};
template <int NP>
__device__ __forceinline__ void mix_load(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, Mix<NP> &m)
{
    m.G = 0.0; m.Kb = 0.0; m.eta_reg = 0.0; m.is_pl = false;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (p < pt.n) {
            const double r = __ldg(ph + (size_t)p * stride + q);
            m.r[p] = r;
            m.G += (r == 0.0) ? 0.0 : pt.G[p] * r;
            m.Kb += (r == 0.0) ? 0.0 : pt.Kb[p] * r;
            const bool pl = (r != 0.0) && pt.has_pl[p];
            if (pl) m.is_pl = true;
            m.eta_reg += (pl ? pt.eta_vp[p] : 0.0) * r;
        } else
            m.r[p] = 0.0;
    }
}
template <int NP>
__device__ __forceinline__ double mix_yield_F(const jr_phase_tab &pt, const Mix<NP> &m, double P, double tII)
{
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (p < pt.n) {
            const double r = m.r[p];
            double v = 0.0;
            if (r != 0.0) {
                const double Fp = pt.has_pl[p] ? (tII - pt.cosphi[p] * pt.C[p] - pt.sinphi[p] * (P - 0.0)) - 2 * pt.eta_vp[p] * (0.0 * 0.5) : tII;
                v = r * Fp;
            }
            acc = p == 0 ? v : acc + v;
        }
    }
    return acc;
}
template <int NP>
__device__ __forceinline__ void mix_dP(const jr_phase_tab &pt, const Mix<NP> &m, double &dQdP, double &dFdP)
{
    dQdP = 0.0; dFdP = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (p < pt.n) {
            const double r = m.r[p];
            if (r != 0.0) {
                dQdP = fma(r, pt.has_pl[p] ? -pt.sinpsi[p] : 0.0, dQdP);
                dFdP = fma(r, pt.has_pl[p] ? -pt.sinphi[p] : 0.0, dFdP);
            }
        }
    }
}
// one component of ∂Q/∂τ (tensor convention, shear slots halved) at the trial stress t with second invariant tII — evaluated lazily,
// only where the node yields (the value is the one compute_plastic_gradients_phase returns: same operations, same order)
template <int NP, bool SHEAR>
__device__ __forceinline__ double mix_dQdt(const jr_phase_tab &pt, const Mix<NP> &m, double t, double tII)
{
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (p < pt.n) {
            const double r = m.r[p];
            if (r != 0.0) {
                const double g = pt.has_pl[p] ? (SHEAR ? 0.5 * (t / tII) : 0.5 * t / tII) : 0.0;
                acc = fma(r, g, acc);
            }
        }
    }
    return acc;
}

// one edge family of update_stresses_center_vertex_ps!  StressKernels.jl:716-778 (yz), :781-849 (xz), :852-921 (xy).
// t/to/e: the six Voigt components (xx, yy, zz, yz, xz, xy) interpolated to the edge; SLOT = the component that lives on this edge.
template <int SLOT, bool DIAG, int NP>
__device__ __forceinline__ void vc3_edge_mix(const V3 &a, const jr_phase_tab &pt, const Mix<NP> &m, size_t v, double etav, double Pv, const double (&t)[6],
                                             const double (&to)[6], const double (&e)[6], double *__restrict__ lamv, double *__restrict__ tau_out,
                                             double *__restrict__ epl)
{
    const double _Gdt = jr_inv(m.G * a.dt);
    const double dtr = jr_inv(a.th + etav * _Gdt + 1.0);
    double trial[6], dS = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        const double d = jr_stress_increment(t[q], to[q], etav, e[q], _Gdt, dtr);
        trial[q] = t[q] + d;
        if (q == SLOT) dS = d;
    }
    const double tII = jr_second_invariant<6>(trial);
    double tn = t[SLOT] + dS, ep = 0.0;
    if (m.is_pl && tII != 0.0) {
        const double Fv = mix_yield_F<NP>(pt, m, Pv, tII);
        if (Fv > 0) {
            double dQdP, dFdP;
            mix_dP<NP>(pt, m, dQdP, dFdP);
            const double volume = isinf(m.Kb) ? 0.0 : m.Kb * a.dt * dFdP * dQdP;
            const double l = (1.0 - a.rel) * lamv[v] + a.rel * (fmax(Fv, 0.0) / (etav * dtr + m.eta_reg + volume));
            lamv[v] = l;
            ep = l * mix_dQdt<NP, true>(pt, m, trial[SLOT], tII);
            tn = t[SLOT] + fma(-2.0, etav * ep * dtr, dS);
        }
    }
    tau_out[v] = tn;
    if (DIAG) epl[v] = ep;
}
template <int SLOT, bool DIAG, int NP>
__device__ __forceinline__ void vc3_edge(const V3 &a, const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t v, double etav, double Pv,
                                         const double (&t)[6], const double (&to)[6], const double (&e)[6], double *__restrict__ lamv,
                                         double *__restrict__ tau_out, double *__restrict__ epl)
{
    Mix<NP> m;
    mix_load<NP>(pt, ph, stride, v, m);
    vc3_edge_mix<SLOT, DIAG, NP>(a, pt, m, v, etav, Pv, t, to, e, lamv, tau_out, epl);
}

// the cell-centre part of update_stresses_center_vertex_ps!  StressKernels.jl:923-986 (plain products and sums: no @muladd there);
// eij / tij / tijo: strain rate, stress, old stress at the centre in Voigt order (shear: 4-edge averages / centre copies)
template <bool DIAG, int NP>
__device__ __forceinline__ void vc3_centre(const V3 &a, const jr_phase_tab &pt, const Mix<NP> &m, size_t c, double et, double Pr, const double (&eij)[6],
                                           double (&tij)[6], const double (&tijo)[6], double lam_in)
{
    const double _Gdt = jr_inv(m.G * a.dt), K = m.Kb;
    const double dtr = jr_inv(a.th + et * _Gdt + 1.0);
    double d[6], trial[6];
#pragma unroll
    for (int q = 0; q < 6; q++) {
        d[q] = (-(tij[q] - tijo[q]) * et * _Gdt - tij[q] + 2.0 * et * eij[q]) * dtr;
        trial[q] = tij[q] + d[q];
    }
    double tII = jr_second_invariant<6>(trial);
    double dQdP, dFdP;
    mix_dP<NP>(pt, m, dQdP, dFdP);
    double lam = lam_in, evol = 0.0, epl[3] = {0.0, 0.0, 0.0};   // lam_in = a.lam[c], loaded by the caller (early in the staged kernel)
    double Fc = 0.0;
    if (m.is_pl && tII != 0.0) Fc = mix_yield_F<NP>(pt, m, Pr, tII);
    if (m.is_pl && tII != 0.0 && Fc > 0) {
        const double volume = isinf(K) ? 0.0 : K * a.dt * dFdP * dQdP;
        lam = (1.0 - a.rel) * lam + a.rel * (fmax(Fc, 0.0) / (et * dtr + m.eta_reg + volume));
        a.lam[c] = lam;
        const double tIIt = tII;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const double e = lam * (q < 3 ? mix_dQdt<NP, false>(pt, m, trial[q], tIIt) : mix_dQdt<NP, true>(pt, m, trial[q], tIIt));
            if (q < 3) epl[q] = e;
            d[q] = d[q] - 2.0 * et * e * dtr;
            tij[q] = d[q] + tij[q];
        }
        evol = -lam * dQdP;
        tII = jr_second_invariant<6>(tij);
    } else {
#pragma unroll
        for (int q = 0; q < 6; q++) tij[q] = d[q] + tij[q];
    }
    a.txx_o[c] = tij[0]; a.tyy_o[c] = tij[1]; a.tzz_o[c] = tij[2];
    a.tyzc[c] = tij[3]; a.txzc[c] = tij[4]; a.txyc[c] = tij[5];
    a.P[c] = Pr - (isinf(K) ? 0.0 : K * a.dt * lam * dQdP);
    if (DIAG) {
        a.pxx[c] = epl[0]; a.pyy[c] = epl[1]; a.pzz[c] = epl[2];
        a.e_vol_pl[c] = evol;
        a.tII[c] = tII;
        a.eta_vep[c] = tII * 0.5 * jr_inv(jr_second_invariant<6>(eij));
    }
}

// Node (i,j,k) of the (nx+1, ny+1, nz+1) lattice updates its yz, xz, xy edges and its cell centre.  All neighbour addresses are a family
// base index (cell / yz / xz / xy array shapes) plus small precomputed offsets (clamped: 0 or ± one stride).
template <bool DIAG, int NP>
__device__ __forceinline__ void vc3_stress_body(const V3 &a, const jr_phase_tab &pt)
{
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    int i, j;
    vc3_map_ij_b(nx + 1, a.xfull_n, blockIdx.x + a.bo_x, blockIdx.y + a.bo_y, i, j);
    const int k = blockIdx.z + a.bo_z + 1;
    if (i > nx + 1 || j > ny + 1 || k > nz + 1) return;
    if (i <= a.zm_i && j <= a.zm_j && k <= a.zm_k) return;   // updated by k_vc3_stress_zm
    const size_t nc = (size_t)nx * ny * nz, nyz = (size_t)nx * (ny + 1) * (nz + 1), nxz = (size_t)(nx + 1) * ny * (nz + 1),
                 nxy = (size_t)(nx + 1) * (ny + 1) * nz;
    const int i0 = jr_clamp(i - 1, 1, nx), ic = jr_clamp(i, 1, nx), i1 = jr_clamp(i + 1, 1, nx);
    const int j0 = jr_clamp(j - 1, 1, ny), jc = jr_clamp(j, 1, ny), j1 = jr_clamp(j + 1, 1, ny);
    const int k0 = jr_clamp(k - 1, 1, nz), kc = jr_clamp(k, 1, nz), k1 = jr_clamp(k + 1, 1, nz);
    // family bases at (ic, jc, kc) and offsets
    const int csy = nx, csz = nx * ny;                    // cell arrays (nx, ny, nz)
    const int ysy = nx, ysz = nx * (ny + 1);              // yz arrays   (nx, ny+1, nz+1)
    const int zsy = nx + 1, zsz = (nx + 1) * ny;          // xz arrays   (nx+1, ny, nz+1)
    const int xsy = nx + 1, xsz = (nx + 1) * (ny + 1);    // xy arrays   (nx+1, ny+1, nz)
    const size_t cb = IX3(nx, ny, ic, jc, kc), yb = IX3(nx, ny + 1, ic, jc, kc), zb = IX3(nx + 1, ny, ic, jc, kc), xb = IX3(nx + 1, ny + 1, ic, jc, kc);
    const int di0 = i0 - ic, di1 = i1 - ic, dj0 = j0 - jc, dj1 = j1 - jc, dk0 = k0 - kc, dk1 = k1 - kc;
    const int ci0 = di0, cj0 = dj0 * csy, ck0 = dk0 * csz;
    const double *__restrict__ eta = a.eta_o;  // the relaxed viscosity of this iteration
#define LC(A, o) __ldg((A) + cb + (o))
#define LY(A, o) __ldg((A) + yb + (o))
#define LZ(A, o) __ldg((A) + zb + (o))
#define LX(A, o) __ldg((A) + xb + (o))
    // clamped averages  StressKernels.jl:620-669 (argument order = summation order)
#define AV_YZ(A) (0.25 * (LC(A, cj0 + ck0) + LC(A, ck0) + LC(A, cj0) + LC(A, 0)))
#define AV_XZ(A) (0.25 * (LC(A, ci0 + ck0) + LC(A, ck0) + LC(A, ci0) + LC(A, 0)))
#define AV_XY(A) (0.25 * (LC(A, ci0 + cj0) + LC(A, cj0) + LC(A, ci0) + LC(A, 0)))
#define HARM_YZ(A) (4 / (1 / LC(A, cj0 + ck0) + 1 / LC(A, ck0) + 1 / LC(A, cj0) + 1 / LC(A, 0)))
#define HARM_XZ(A) (4 / (1 / LC(A, ci0 + ck0) + 1 / LC(A, ck0) + 1 / LC(A, ci0) + 1 / LC(A, 0)))
#define HARM_XY(A) (4 / (1 / LC(A, ci0 + cj0) + 1 / LC(A, cj0) + 1 / LC(A, ci0) + 1 / LC(A, 0)))
#define AV_YZ_Z(A) (0.25 * (LX(A, dk0 * xsz) + LX(A, di1 + dk0 * xsz) + LX(A, 0) + LX(A, di1)))                       /* xy arrays */
#define AV_YZ_Y(A) (0.25 * (LZ(A, dj0 * zsy) + LZ(A, di1 + dj0 * zsy) + LZ(A, 0) + LZ(A, di1)))                       /* xz arrays */
#define AV_XZ_Z(A) (0.25 * (LX(A, dk0 * xsz) + LX(A, dj1 * xsy + dk0 * xsz) + LX(A, 0) + LX(A, dj1 * xsy)))           /* xy arrays */
#define AV_XZ_X(A) (0.25 * (LY(A, di0) + LY(A, 0) + LY(A, dj1 * ysy) + LY(A, di0 + dj1 * ysy)))                       /* yz arrays */
#define AV_XY_Y(A) (0.25 * (LZ(A, dj0 * zsy) + LZ(A, 0) + LZ(A, dj0 * zsy + dk1 * zsz) + LZ(A, dk1 * zsz)))           /* xz arrays */
#define AV_XY_X(A) (0.25 * (LY(A, di0) + LY(A, 0) + LY(A, di0 + dk1 * ysz) + LY(A, dk1 * ysz)))                       /* yz arrays */
    if (i <= nx && j <= ny + 1 && k <= nz + 1) {  // ---- yz edge
        const size_t v = IX3(nx, ny + 1, i, j, k);
        const double t[6] = {AV_YZ(a.txx_i), AV_YZ(a.tyy_i), AV_YZ(a.tzz_i), __ldg(a.tyz_i + v), AV_YZ_Y(a.txz_i), AV_YZ_Z(a.txy_i)};
        const double to[6] = {AV_YZ(a.oxx), AV_YZ(a.oyy), AV_YZ(a.ozz), __ldg(a.oyz + v), AV_YZ_Y(a.oxz), AV_YZ_Z(a.oxy)};
        const double e[6] = {AV_YZ(a.exx), AV_YZ(a.eyy), AV_YZ(a.ezz), __ldg(a.eyz + v), AV_YZ_Y(a.exz), AV_YZ_Z(a.exy)};
        vc3_edge<3, DIAG, NP>(a, pt, a.ph_yz, nyz, v, HARM_YZ(eta), AV_YZ(a.theta), t, to, e, a.lamyz, a.tyz_o, a.pyz);
    }
    if (i <= nx + 1 && j <= ny && k <= nz + 1) {  // ---- xz edge
        const size_t v = IX3(nx + 1, ny, i, j, k);
        const double t[6] = {AV_XZ(a.txx_i), AV_XZ(a.tyy_i), AV_XZ(a.tzz_i), AV_XZ_X(a.tyz_i), __ldg(a.txz_i + v), AV_XZ_Z(a.txy_i)};
        const double to[6] = {AV_XZ(a.oxx), AV_XZ(a.oyy), AV_XZ(a.ozz), AV_XZ_X(a.oyz), __ldg(a.oxz + v), AV_XZ_Z(a.oxy)};
        const double e[6] = {AV_XZ(a.exx), AV_XZ(a.eyy), AV_XZ(a.ezz), AV_XZ_X(a.eyz), __ldg(a.exz + v), AV_XZ_Z(a.exy)};
        vc3_edge<4, DIAG, NP>(a, pt, a.ph_xz, nxz, v, HARM_XZ(eta), AV_XZ(a.theta), t, to, e, a.lamxz, a.txz_o, a.pxz);
    }
    if (i <= nx + 1 && j <= ny + 1 && k <= nz) {  // ---- xy edge
        const size_t v = IX3(nx + 1, ny + 1, i, j, k);
        const double t[6] = {AV_XY(a.txx_i), AV_XY(a.tyy_i), AV_XY(a.tzz_i), AV_XY_X(a.tyz_i), AV_XY_Y(a.txz_i), __ldg(a.txy_i + v)};
        const double to[6] = {AV_XY(a.oxx), AV_XY(a.oyy), AV_XY(a.ozz), AV_XY_X(a.oyz), AV_XY_Y(a.oxz), __ldg(a.oxy + v)};
        const double e[6] = {AV_XY(a.exx), AV_XY(a.eyy), AV_XY(a.ezz), AV_XY_X(a.eyz), AV_XY_Y(a.exz), __ldg(a.exy + v)};
        vc3_edge<5, DIAG, NP>(a, pt, a.ph_xy, nxy, v, HARM_XY(eta), AV_XY(a.theta), t, to, e, a.lamxy, a.txy_o, a.pxy);
    }
    if (i <= nx && j <= ny && k <= nz) {  // ---- centre  StressKernels.jl:923-986 (plain products and sums: no @muladd there)
        // cache_tensors  StressUpdate.jl:248-301 (_av_yz/_av_xz/_av_xy: mysum order k → j → i starting from 0.0, quirk Q15)
        const size_t c = cb;
        const double eij[6] = {__ldg(a.exx + c), __ldg(a.eyy + c), __ldg(a.ezz + c),
                               0.25 * ((((0.0 + LY(a.eyz, 0)) + LY(a.eyz, ysy)) + LY(a.eyz, ysz)) + LY(a.eyz, ysy + ysz)),
                               0.25 * ((((0.0 + LZ(a.exz, 0)) + LZ(a.exz, 1)) + LZ(a.exz, zsz)) + LZ(a.exz, 1 + zsz)),
                               0.25 * ((((0.0 + LX(a.exy, 0)) + LX(a.exy, 1)) + LX(a.exy, xsy)) + LX(a.exy, 1 + xsy))};
        double tij[6] = {__ldg(a.txx_i + c), __ldg(a.tyy_i + c), __ldg(a.tzz_i + c), a.tyzc[c], a.txzc[c], a.txyc[c]};
        const double tijo[6] = {__ldg(a.oxx + c), __ldg(a.oyy + c), __ldg(a.ozz + c), __ldg(a.oyzc + c), __ldg(a.oxzc + c), __ldg(a.oxyc + c)};
        Mix<NP> m;
        mix_load<NP>(pt, a.ph_c, nc, c, m);
        vc3_centre<DIAG, NP>(a, pt, m, c, __ldg(eta + c), a.theta[c], eij, tij, tijo, a.lam[c]);
    }
}

// State of one edge family while the six Voigt components stream through: visco-elastic factors, the running sum of the second invariant of
// the trial stress (reference order sqrt(0.5 (T0² + T1² + T2²) + T3² + T4² + T5²)) and the own component.  Lets the three edges of a node
// advance as independent instruction streams (ILP) before the rarely taken plastic branches.
struct EdgeAcc {
    double etav, Pv, _Gdt, dtr, acc, t_own, d_own, trial_own;
};
template <int Q, bool OWN>
__device__ __forceinline__ void edge_comp(EdgeAcc &E, double t, double to, double e)
{
    const double d = jr_stress_increment(t, to, E.etav, e, E._Gdt, E.dtr);
    const double T = t + d;
    if (Q == 0) E.acc = T * T;
    else if (Q < 3) E.acc = E.acc + T * T;
    else if (Q == 3) E.acc = 0.5 * E.acc + T * T;
    else E.acc = E.acc + T * T;
    if (OWN) { E.t_own = t; E.d_own = d; E.trial_own = T; }
}
template <bool DIAG, int NP>
__device__ __forceinline__ void edge_finish(const V3 &a, const jr_phase_tab &pt, const Mix<NP> &m, const EdgeAcc &E, size_t v, double *__restrict__ lamv,
                                            double *__restrict__ tau_out, double *__restrict__ epl)
{
    const double tII = sqrt(E.acc);
    double tn = E.t_own + E.d_own, ep = 0.0;
    if (m.is_pl && tII != 0.0) {
        const double Fv = mix_yield_F<NP>(pt, m, E.Pv, tII);
        if (Fv > 0) {
            double dQdP, dFdP;
            mix_dP<NP>(pt, m, dQdP, dFdP);
            const double volume = isinf(m.Kb) ? 0.0 : m.Kb * a.dt * dFdP * dQdP;
            const double l = (1.0 - a.rel) * lamv[v] + a.rel * (fmax(Fv, 0.0) / (E.etav * E.dtr + m.eta_reg + volume));
            lamv[v] = l;
            ep = l * mix_dQdt<NP, true>(pt, m, E.trial_own, tII);
            tn = E.t_own + fma(-2.0, E.etav * ep * E.dtr, E.d_own);
        }
    }
    tau_out[v] = tn;
    if (DIAG) epl[v] = ep;
}

// ---- shared-memory staged variant ----------------------------------------------------------------------------------------------------
// A CTA of 32 × TYS nodes of plane k first copies the tiles of the 20 neighbour-read arrays it needs (2 planes × (TYS+1) rows × 33
// columns each; four tile origins: cell, yz, xz, xy families) global → shared with cp.async — ≈ 46 independent 8-byte copies in flight per
// thread and no registers held — then every node computes from shared memory with clamp-free indices.  CTAs that touch the high-side
// boundary planes (where the clamped and the raw indices of the reference differ between families) take the global-memory body above.
// The z-neighbour CTA re-reads one of the two planes: L2 serves it (CTAs are scheduled plane by plane).
#ifndef TYS
#define TYS 8
#endif
#define SROW 33
#define SPLANE (SROW * (TYS + 1))
#define STILE (2 * SPLANE)
enum { SL_eta, SL_theta, SL_txx, SL_tyy, SL_tzz, SL_oxx, SL_oyy, SL_ozz, SL_exx, SL_eyy, SL_ezz,   // cell family
       SL_tyz, SL_oyz, SL_eyz,                                                                       // yz family
       SL_txz, SL_oxz, SL_exz,                                                                       // xz family
       SL_txy, SL_oxy, SL_exy, SL_COUNT };

__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}

template <bool DIAG, int NP>
__device__ __noinline__ void vc3_stress_body_call(const V3 &a, const jr_phase_tab &pt) { vc3_stress_body<DIAG, NP>(a, pt); }

template <int NP>
__device__ __forceinline__ void ratios_load(const jr_phase_tab &pt, const double *__restrict__ ph, size_t stride, size_t q, double (&r)[NP])
{
#pragma unroll
    for (int p = 0; p < NP; p++) r[p] = p < pt.n ? __ldg(ph + (size_t)p * stride + q) : 0.0;
}
template <int NP>
__device__ __forceinline__ void mix_from(const jr_phase_tab &pt, const double (&r)[NP], Mix<NP> &m)
{
    m.G = 0.0; m.Kb = 0.0; m.eta_reg = 0.0; m.is_pl = false;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        m.r[p] = r[p];
        if (p < pt.n) {
            m.G += (r[p] == 0.0) ? 0.0 : pt.G[p] * r[p];
            m.Kb += (r[p] == 0.0) ? 0.0 : pt.Kb[p] * r[p];
            const bool pl = (r[p] != 0.0) && pt.has_pl[p];
            if (pl) m.is_pl = true;
            m.eta_reg += (pl ? pt.eta_vp[p] : 0.0) * r[p];
        }
    }
}

// tile accessors of the staged stress kernels: g<s, I, J, K>() = entry of array slot s at column tx+I, row ty+J, plane K (0 / 1) of the thread's
// 2 × 2 × 2 neighbourhood — per family: cell (i0|ic, j0|jc, k0|kc); yz (i0|ic, j|j+1, k|k+1); xz (i|i+1, j0|jc, k|k+1); xy (i|i+1, j|j+1, k0|kc)
struct TilePair {   // k_vc3_stress_sm: two planes per array, staged per CTA
    const double *sm;
    int base;       // ty · SROW + tx
    template <int s, int I, int J, int K> __device__ __forceinline__ double g() const { return sm[s * STILE + K * SPLANE + J * SROW + I + base]; }
};
template <int ZPLANE_, int ZROW_>
struct TileRing {   // k_vc3_stress_zm: three-plane ring per array; the cell / xy families hold planes (k−1, k), the yz / xz families (k, k+1)
    const double *sm;
    int base;       // ty · ZROW + tx
    int aLo, aHi, bLo, bHi;   // ring offsets (slot · ZPLANE) of the two planes of either group at this z-step
    template <int s, int I, int J, int K> __device__ __forceinline__ double g() const
    {
        const int off = (s >= SL_tyz && s < SL_txy) ? (K ? bHi : bLo) : (K ? aHi : aLo);
        return sm[s * (3 * ZPLANE_) + off + J * ZROW_ + I + base];
    }
};

// the arithmetic of one node (centre first, then the three edges) from the staged tiles and the node's own-position operands
template <bool DIAG, int NP, class Tile>
__device__ __forceinline__ void vc3_sm_compute(const V3 &a, const jr_phase_tab &pt, const Tile &A, size_t c, size_t vyz, size_t vxz, size_t vxy,
                                               const double (&rc)[NP], const double (&ryz)[NP], const double (&rxz)[NP], const double (&rxy)[NP],
                                               double c_tyz, double c_txz, double c_txy, double c_oyz, double c_oxz, double c_oxy, double c_lam)
{
#define S(s, I, J, K) A.template g<(s), (I), (J), (K)>()
#define SAV_YZ(s) (0.25 * (S(s, 1, 0, 0) + S(s, 1, 1, 0) + S(s, 1, 0, 1) + S(s, 1, 1, 1)))
#define SAV_XZ(s) (0.25 * (S(s, 0, 1, 0) + S(s, 1, 1, 0) + S(s, 0, 1, 1) + S(s, 1, 1, 1)))
#define SAV_XY(s) (0.25 * (S(s, 0, 0, 1) + S(s, 1, 0, 1) + S(s, 0, 1, 1) + S(s, 1, 1, 1)))
#define SHARM_YZ(s) jr_div_nr(4.0, jr_inv_nr(S(s, 1, 0, 0)) + jr_inv_nr(S(s, 1, 1, 0)) + jr_inv_nr(S(s, 1, 0, 1)) + jr_inv_nr(S(s, 1, 1, 1)))
#define SHARM_XZ(s) jr_div_nr(4.0, jr_inv_nr(S(s, 0, 1, 0)) + jr_inv_nr(S(s, 1, 1, 0)) + jr_inv_nr(S(s, 0, 1, 1)) + jr_inv_nr(S(s, 1, 1, 1)))
#define SHARM_XY(s) jr_div_nr(4.0, jr_inv_nr(S(s, 0, 0, 1)) + jr_inv_nr(S(s, 1, 0, 1)) + jr_inv_nr(S(s, 0, 1, 1)) + jr_inv_nr(S(s, 1, 1, 1)))
    /* harmonic means: η is strictly positive and in the normal range, so the branch-free IEEE-exact reciprocal / quotient sequences of
       tma.cuh give the same bits as 1 / x and 4 / x without the slow-path call scaffolding (15 reciprocals per node) */
#define SAV_YZ_Y(s) (0.25 * (S(s, 0, 0, 0) + S(s, 1, 0, 0) + S(s, 0, 1, 0) + S(s, 1, 1, 0)))   /* xz family: (ic,j0,kc),(i1,j0,kc),(ic,jc,kc),(i1,jc,kc) */
#define SAV_YZ_Z(s) (0.25 * (S(s, 0, 0, 0) + S(s, 1, 0, 0) + S(s, 0, 0, 1) + S(s, 1, 0, 1)))   /* xy family: (ic,jc,k0),(i1,jc,k0),(ic,jc,kc),(i1,jc,kc) */
#define SAV_XZ_X(s) (0.25 * (S(s, 0, 0, 0) + S(s, 1, 0, 0) + S(s, 1, 1, 0) + S(s, 0, 1, 0)))   /* yz family: (i0,jc,kc),(ic,jc,kc),(ic,j1,kc),(i0,j1,kc) */
#define SAV_XZ_Z(s) (0.25 * (S(s, 0, 0, 0) + S(s, 0, 1, 0) + S(s, 0, 0, 1) + S(s, 0, 1, 1)))   /* xy family: (ic,jc,k0),(ic,j1,k0),(ic,jc,kc),(ic,j1,kc) */
#define SAV_XY_X(s) (0.25 * (S(s, 0, 0, 0) + S(s, 1, 0, 0) + S(s, 0, 0, 1) + S(s, 1, 0, 1)))   /* yz family: (i0,jc,kc),(ic,jc,kc),(i0,jc,k1),(ic,jc,k1) */
#define SAV_XY_Y(s) (0.25 * (S(s, 0, 0, 0) + S(s, 0, 1, 0) + S(s, 0, 0, 1) + S(s, 0, 1, 1)))   /* xz family: (ic,j0,kc),(ic,jc,kc),(ic,j0,k1),(ic,jc,k1) */
    {   // ---- centre: cell (i, j, k) = S(·, 1, 1, 1); edge gathers in mysum order (quirk Q15)
        const double eij[6] = {S(SL_exx, 1, 1, 1), S(SL_eyy, 1, 1, 1), S(SL_ezz, 1, 1, 1),
                               0.25 * ((((0.0 + S(SL_eyz, 1, 0, 0)) + S(SL_eyz, 1, 1, 0)) + S(SL_eyz, 1, 0, 1)) + S(SL_eyz, 1, 1, 1)),
                               0.25 * ((((0.0 + S(SL_exz, 0, 1, 0)) + S(SL_exz, 1, 1, 0)) + S(SL_exz, 0, 1, 1)) + S(SL_exz, 1, 1, 1)),
                               0.25 * ((((0.0 + S(SL_exy, 0, 0, 1)) + S(SL_exy, 1, 0, 1)) + S(SL_exy, 0, 1, 1)) + S(SL_exy, 1, 1, 1))};
        double tij[6] = {S(SL_txx, 1, 1, 1), S(SL_tyy, 1, 1, 1), S(SL_tzz, 1, 1, 1), c_tyz, c_txz, c_txy};
        const double tijo[6] = {S(SL_oxx, 1, 1, 1), S(SL_oyy, 1, 1, 1), S(SL_ozz, 1, 1, 1), c_oyz, c_oxz, c_oxy};
        Mix<NP> mc;
        mix_from<NP>(pt, rc, mc);
        vc3_centre<DIAG, NP>(a, pt, mc, c, S(SL_eta, 1, 1, 1), S(SL_theta, 1, 1, 1), eij, tij, tijo, c_lam);
    }
    // ---- the three edges advance together (straight-line code: three independent dependency chains), then the plastic branches
    Mix<NP> myz, mxz, mxy;
    mix_from<NP>(pt, ryz, myz);
    mix_from<NP>(pt, rxz, mxz);
    mix_from<NP>(pt, rxy, mxy);
    EdgeAcc Eyz, Exz, Exy;
    Eyz.etav = SHARM_YZ(SL_eta); Exz.etav = SHARM_XZ(SL_eta); Exy.etav = SHARM_XY(SL_eta);
    Eyz.Pv = SAV_YZ(SL_theta); Exz.Pv = SAV_XZ(SL_theta); Exy.Pv = SAV_XY(SL_theta);
    Eyz._Gdt = jr_inv(myz.G * a.dt); Exz._Gdt = jr_inv(mxz.G * a.dt); Exy._Gdt = jr_inv(mxy.G * a.dt);
    Eyz.dtr = jr_inv_nr(a.th + Eyz.etav * Eyz._Gdt + 1.0);   // operand ≥ 1
    Exz.dtr = jr_inv_nr(a.th + Exz.etav * Exz._Gdt + 1.0);
    Exy.dtr = jr_inv_nr(a.th + Exy.etav * Exy._Gdt + 1.0);
#define COMP(Q, OYZ, OXZ, OXY, TYZ, TXZ, TXY, OLDYZ, OLDXZ, OLDXY, EPYZ, EPXZ, EPXY)                                                   \
    edge_comp<Q, OYZ>(Eyz, TYZ, OLDYZ, EPYZ);                                                                                          \
    edge_comp<Q, OXZ>(Exz, TXZ, OLDXZ, EPXZ);                                                                                          \
    edge_comp<Q, OXY>(Exy, TXY, OLDXY, EPXY);
    COMP(0, false, false, false, SAV_YZ(SL_txx), SAV_XZ(SL_txx), SAV_XY(SL_txx), SAV_YZ(SL_oxx), SAV_XZ(SL_oxx), SAV_XY(SL_oxx), SAV_YZ(SL_exx),
         SAV_XZ(SL_exx), SAV_XY(SL_exx))
    COMP(1, false, false, false, SAV_YZ(SL_tyy), SAV_XZ(SL_tyy), SAV_XY(SL_tyy), SAV_YZ(SL_oyy), SAV_XZ(SL_oyy), SAV_XY(SL_oyy), SAV_YZ(SL_eyy),
         SAV_XZ(SL_eyy), SAV_XY(SL_eyy))
    COMP(2, false, false, false, SAV_YZ(SL_tzz), SAV_XZ(SL_tzz), SAV_XY(SL_tzz), SAV_YZ(SL_ozz), SAV_XZ(SL_ozz), SAV_XY(SL_ozz), SAV_YZ(SL_ezz),
         SAV_XZ(SL_ezz), SAV_XY(SL_ezz))
    // yz component: own on the yz edge (yz family (ic, j, k) = S(·, 1, 0, 0)); av_clamped_xz_x, av_clamped_xy_x elsewhere
    COMP(3, true, false, false, S(SL_tyz, 1, 0, 0), SAV_XZ_X(SL_tyz), SAV_XY_X(SL_tyz), S(SL_oyz, 1, 0, 0), SAV_XZ_X(SL_oyz), SAV_XY_X(SL_oyz),
         S(SL_eyz, 1, 0, 0), SAV_XZ_X(SL_eyz), SAV_XY_X(SL_eyz))
    // xz component: own on the xz edge (xz family (i, jc, k) = S(·, 0, 1, 0))
    COMP(4, false, true, false, SAV_YZ_Y(SL_txz), S(SL_txz, 0, 1, 0), SAV_XY_Y(SL_txz), SAV_YZ_Y(SL_oxz), S(SL_oxz, 0, 1, 0), SAV_XY_Y(SL_oxz),
         SAV_YZ_Y(SL_exz), S(SL_exz, 0, 1, 0), SAV_XY_Y(SL_exz))
    // xy component: own on the xy edge (xy family (i, j, kc) = S(·, 0, 0, 1))
    COMP(5, false, false, true, SAV_YZ_Z(SL_txy), SAV_XZ_Z(SL_txy), S(SL_txy, 0, 0, 1), SAV_YZ_Z(SL_oxy), SAV_XZ_Z(SL_oxy), S(SL_oxy, 0, 0, 1),
         SAV_YZ_Z(SL_exy), SAV_XZ_Z(SL_exy), S(SL_exy, 0, 0, 1))
#undef COMP
    edge_finish<DIAG, NP>(a, pt, myz, Eyz, vyz, a.lamyz, a.tyz_o, a.pyz);
    edge_finish<DIAG, NP>(a, pt, mxz, Exz, vxz, a.lamxz, a.txz_o, a.pxz);
    edge_finish<DIAG, NP>(a, pt, mxy, Exy, vxy, a.lamxy, a.txy_o, a.pxy);
#undef S
#undef SAV_YZ
#undef SAV_XZ
#undef SAV_XY
#undef SHARM_YZ
#undef SHARM_XZ
#undef SHARM_XY
#undef SAV_YZ_Y
#undef SAV_YZ_Z
#undef SAV_XZ_X
#undef SAV_XZ_Z
#undef SAV_XY_X
#undef SAV_XY_Y
}

template <bool DIAG, int NP>
__global__ void __launch_bounds__(32 * TYS, 16 / TYS) k_vc3_stress_sm(const __grid_constant__ V3 a, const __grid_constant__ jr_phase_tab pt)
{
    extern __shared__ double sm[];
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int ib = (blockIdx.x + a.bo_x) * 32 + 1, jb = (blockIdx.y + a.bo_y) * TYS + 1, k = blockIdx.z + a.bo_z + 1;   // first node of the CTA (1-based)
    // nodes of the z-marching kernel's region (k_vc3_stress_zm): a CTA entirely inside it has nothing to do, one that straddles its rim
    // takes the per-node body (which skips the region's nodes)
    if (k <= a.zm_k && ib <= a.zm_i && jb <= a.zm_j) {
        if (ib + 31 <= a.zm_i && jb + TYS - 1 <= a.zm_j) return;
        vc3_stress_body_call<DIAG, NP>(a, pt);
        return;
    }
    // fast CTAs: every node has i+1 ≤ nx, j+1 ≤ ny, k+1 ≤ nz (no high-side clamp is active)
    if (!(ib + 31 <= nx - 1 && jb + TYS - 1 <= ny - 1 && k <= nz - 1)) {
        vc3_stress_body_call<DIAG, NP>(a, pt);
        return;
    }
    const int i = ib + tx, j = jb + ty;
    {   // ---- stage the tiles
        const double *src[SL_COUNT] = {a.eta_o, a.theta, a.txx_i, a.tyy_i, a.tzz_i, a.oxx, a.oyy, a.ozz, a.exx, a.eyy, a.ezz,
                                       a.tyz_i, a.oyz, a.eyz, a.txz_i, a.oxz, a.exz, a.txy_i, a.oxy, a.exy};
#pragma unroll
        for (int it = 0; it < (STILE + 32 * TYS - 1) / (32 * TYS); it++) {
            const int e = tid + it * 32 * TYS;
            if (e < STILE) {
                const int p = e / SPLANE, r = (e - p * SPLANE) / SROW, c = e - p * SPLANE - r * SROW;
                // tile origins (low-side clamp to index 1): cell (i−1, j−1, k−1); yz (i−1, j, k); xz (i, j−1, k); xy (i, j, k−1)
                const int im = max(ib - 1 + c, 1), jm = max(jb - 1 + r, 1), km = max(k - 1 + p, 1), ip = ib + c, jp = jb + r, kp = k + p;
                const size_t oc = IX3(nx, ny, im, jm, km), oy = IX3(nx, ny + 1, im, jp, kp), oz = IX3(nx + 1, ny, ip, jm, kp),
                             ox = IX3(nx + 1, ny + 1, ip, jp, km);
#pragma unroll
                for (int s = 0; s < SL_COUNT; s++) {
                    const size_t o = s < SL_tyz ? oc : (s < SL_txz ? oy : (s < SL_txy ? oz : ox));
                    cp_async8(sm + s * STILE + e, src[s] + o);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // global reads that do not go through the tiles are issued while the copies are in flight
    const size_t nc = (size_t)nx * ny * nz, nyz = (size_t)nx * (ny + 1) * (nz + 1), nxz = (size_t)(nx + 1) * ny * (nz + 1),
                 nxy = (size_t)(nx + 1) * (ny + 1) * nz;
    const size_t c = IX3(nx, ny, i, j, k), vyz = IX3(nx, ny + 1, i, j, k), vxz = IX3(nx + 1, ny, i, j, k), vxy = IX3(nx + 1, ny + 1, i, j, k);
    // L2 prefetch for the CTA of the same tile on plane k + 1 (CTAs are scheduled plane by plane, one plane ≈ one wave): the one new
    // plane of each family it will stage, and its own-position operands.  No registers held; its cp.async / loads then hit L2.
    if (a.pf_next && k + 1 <= nz - 1) {
        const double *src[SL_COUNT] = {a.eta_o, a.theta, a.txx_i, a.tyy_i, a.tzz_i, a.oxx, a.oyy, a.ozz, a.exx, a.eyy, a.ezz,
                                       a.tyz_i, a.oyz, a.eyz, a.txz_i, a.oxz, a.exz, a.txy_i, a.oxy, a.exy};
#pragma unroll
        for (int it = 0; it < (SPLANE + 32 * TYS - 1) / (32 * TYS); it++) {
            const int e = tid + it * 32 * TYS;
            if (e < SPLANE) {
                const int r = e / SROW, cc = e - r * SROW;
                const int im = max(ib - 1 + cc, 1), jm = max(jb - 1 + r, 1), ip = ib + cc, jp = jb + r;
                const size_t oc = IX3(nx, ny, im, jm, k + 1), oy = IX3(nx, ny + 1, im, jp, k + 2), oz = IX3(nx + 1, ny, ip, jm, k + 2),
                             ox = IX3(nx + 1, ny + 1, ip, jp, k + 1);
#pragma unroll
                for (int s = 0; s < SL_COUNT; s++) {
                    const size_t o = s < SL_tyz ? oc : (s < SL_txz ? oy : (s < SL_txy ? oz : ox));
                    jr_prefetch_l2(src[s] + o);
                }
            }
        }
        const size_t sc = (size_t)nx * ny, syz = (size_t)nx * (ny + 1), sxz = (size_t)(nx + 1) * ny, sxy = (size_t)(nx + 1) * (ny + 1);
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (p < pt.n) {
                jr_prefetch_l2(a.ph_yz + (size_t)p * nyz + vyz + syz); jr_prefetch_l2(a.ph_xz + (size_t)p * nxz + vxz + sxz);
                jr_prefetch_l2(a.ph_xy + (size_t)p * nxy + vxy + sxy); jr_prefetch_l2(a.ph_c + (size_t)p * nc + c + sc);
            }
        jr_prefetch_l2(a.tyzc + c + sc); jr_prefetch_l2(a.txzc + c + sc); jr_prefetch_l2(a.txyc + c + sc);
        jr_prefetch_l2(a.oyzc + c + sc); jr_prefetch_l2(a.oxzc + c + sc); jr_prefetch_l2(a.oxyc + c + sc);
        jr_prefetch_l2(a.lam + c + sc); jr_prefetch_l2(a.lamyz + vyz + syz); jr_prefetch_l2(a.lamxz + vxz + sxz); jr_prefetch_l2(a.lamxy + vxy + sxy);
    }
    double ryz[NP], rxz[NP], rxy[NP], rc[NP];
    ratios_load<NP>(pt, a.ph_c, nc, c, rc);
    // the centre's own-position operands, also in flight behind the copies (the centre is updated FIRST below, so they are consumed
    // before the three edges need the registers)
    const double c_tyz = a.tyzc[c], c_txz = a.txzc[c], c_txy = a.txyc[c];
    const double c_oyz = __ldg(a.oyzc + c), c_oxz = __ldg(a.oxzc + c), c_oxy = __ldg(a.oxyc + c), c_lam = a.lam[c];
    ratios_load<NP>(pt, a.ph_yz, nyz, vyz, ryz);
    ratios_load<NP>(pt, a.ph_xz, nxz, vxz, rxz);
    ratios_load<NP>(pt, a.ph_xy, nxy, vxy, rxy);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    TilePair A;
    A.sm = sm; A.base = ty * SROW + tx;
    vc3_sm_compute<DIAG, NP>(a, pt, A, c, vyz, vxz, vxy, rc, ryz, rxz, rxy, c_tyz, c_txz, c_txy, c_oyz, c_oxz, c_oxy, c_lam);
}

// ---- z-marching variant ----------------------------------------------------------------------------------------------------------------
// One CTA owns a column of 32 × TYZ nodes and marches a chunk of planes k = kbeg … kend−1.  Every array of the 20 keeps a ring of THREE tile
// planes in shared memory: a z-step needs two planes per array (cell / xy families: k−1 and k; yz / xz families: k and k+1), so one new plane
// per array arrives per step — half the staging traffic and instructions of k_vc3_stress_sm, no re-fetch of the shared plane — and it is
// requested one step ahead (cp.async into the free ring slot), so the copies run under the arithmetic of the current step.  The node's
// own-position operands (phase ratios of the four families, centre shear copies, λ) are loaded one step ahead into registers.  ONE barrier per
// step: behind it every copy for step k has landed and every thread has left step k−1, whose oldest planes the next request overwrites.
// Region: the nodes whose neighbourhood needs no high-side clamp (i ≤ 32·⌊(nx−1)/32⌋, j ≤ TYZ·⌊(ny−1)/TYZ⌋, k ≤ nz−1); the rim goes to
// k_vc3_stress_sm / the per-node body, which skip the region (V3::zm_*).  Same arithmetic (vc3_sm_compute), bit-identical results.
template <int NP>
struct OwnOps {
    double rc[NP], ryz[NP], rxz[NP], rxy[NP];
    double tyz, txz, txy, oyz, oxz, oxy, lam;
};
template <int NP>
__device__ __forceinline__ void own_load(const V3 &a, const jr_phase_tab &pt, size_t c, size_t vyz, size_t vxz, size_t vxy, size_t nc, size_t nyz, size_t nxz,
                                         size_t nxy, OwnOps<NP> &o)
{
    ratios_load<NP>(pt, a.ph_c, nc, c, o.rc);
    o.tyz = a.tyzc[c]; o.txz = a.txzc[c]; o.txy = a.txyc[c];
    o.oyz = __ldg(a.oyzc + c); o.oxz = __ldg(a.oxzc + c); o.oxy = __ldg(a.oxyc + c); o.lam = a.lam[c];
    ratios_load<NP>(pt, a.ph_yz, nyz, vyz, o.ryz);
    ratios_load<NP>(pt, a.ph_xz, nxz, vxz, o.rxz);
    ratios_load<NP>(pt, a.ph_xy, nxy, vxy, o.rxy);
}

template <int NP, int TYZ>
__global__ void __launch_bounds__(32 * TYZ, 1) k_vc3_stress_zm(const __grid_constant__ V3 a, const __grid_constant__ jr_phase_tab pt)
{
    extern __shared__ double sm[];
    constexpr int ZROW = 33, ZPLANE = ZROW * (TYZ + 1), ZTILE = 3 * ZPLANE, NTHR = 32 * TYZ, NIT = (ZPLANE + NTHR - 1) / NTHR;
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int ib = blockIdx.x * 32 + 1, jb = blockIdx.y * TYZ + 1;
    const int kbeg = blockIdx.z * a.zm_chunk + 1, kend = min(kbeg + a.zm_chunk, a.zm_k + 1);   // planes [kbeg, kend)
    const int i = ib + tx, j = jb + ty;
    const double *src[SL_COUNT] = {a.eta_o, a.theta, a.txx_i, a.tyy_i, a.tzz_i, a.oxx, a.oyy, a.ozz, a.exx, a.eyy, a.ezz,
                                   a.tyz_i, a.oyz, a.eyz, a.txz_i, a.oxz, a.exz, a.txy_i, a.oxy, a.exy};
    const size_t sc = (size_t)nx * ny, syz = (size_t)nx * (ny + 1), sxz = (size_t)(nx + 1) * ny, sxy = (size_t)(nx + 1) * (ny + 1);
    // the tile elements this thread stages (the same ones for every plane): offsets of plane 1 in the four families
    // (tile origins, low-side clamp to index 1: cell (i−1, j−1); yz (i−1, j); xz (i, j−1); xy (i, j))
    size_t e_oc[NIT], e_oy[NIT], e_oz[NIT], e_ox[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int e = tid + it * NTHR, r = e / ZROW, cc = e - r * ZROW;
        const int im = max(ib - 1 + cc, 1), jm = max(jb - 1 + r, 1), ip = ib + cc, jp = jb + r;
        e_oc[it] = IX3(nx, ny, im, jm, 1); e_oy[it] = IX3(nx, ny + 1, im, jp, 1); e_oz[it] = IX3(nx + 1, ny, ip, jm, 1); e_ox[it] = IX3(nx + 1, ny + 1, ip, jp, 1);
    }
    // group A = cell + xy families (plane P → ring slot P % 3), group B = yz + xz families
    auto stage_A = [&](int P) {
        const int slot = (P % 3) * ZPLANE;
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int e = tid + it * NTHR;
            if (e < ZPLANE) {
                const size_t oc = e_oc[it] + (size_t)(P - 1) * sc, ox = e_ox[it] + (size_t)(P - 1) * sxy;
#pragma unroll
                for (int s = 0; s < SL_tyz; s++) cp_async8(sm + s * ZTILE + slot + e, src[s] + oc);
#pragma unroll
                for (int s = SL_txy; s < SL_COUNT; s++) cp_async8(sm + s * ZTILE + slot + e, src[s] + ox);
            }
        }
    };
    auto stage_B = [&](int P) {
        const int slot = (P % 3) * ZPLANE;
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int e = tid + it * NTHR;
            if (e < ZPLANE) {
                const size_t oy = e_oy[it] + (size_t)(P - 1) * syz, oz = e_oz[it] + (size_t)(P - 1) * sxz;
#pragma unroll
                for (int s = SL_tyz; s < SL_txz; s++) cp_async8(sm + s * ZTILE + slot + e, src[s] + oy);
#pragma unroll
                for (int s = SL_txz; s < SL_txy; s++) cp_async8(sm + s * ZTILE + slot + e, src[s] + oz);
            }
        }
    };
    // prologue: the planes of step kbeg
    if (kbeg > 1) stage_A(kbeg - 1);
    stage_A(kbeg);
    stage_B(kbeg);
    stage_B(kbeg + 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const size_t nc = sc * nz, nyz = syz * (nz + 1), nxz = sxz * (nz + 1), nxy = sxy * nz;
    size_t c = IX3(nx, ny, i, j, kbeg), vyz = IX3(nx, ny + 1, i, j, kbeg), vxz = IX3(nx + 1, ny, i, j, kbeg), vxy = IX3(nx + 1, ny + 1, i, j, kbeg);
    OwnOps<NP> cur, nxt;
    own_load<NP>(a, pt, c, vyz, vxz, vxy, nc, nyz, nxz, nxy, cur);
    TileRing<ZPLANE, ZROW> A;
    A.sm = sm; A.base = ty * ZROW + tx;
    for (int k = kbeg; k < kend; ++k) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const bool more = k + 1 < kend;
        if (more) {   // the planes step k+1 adds, and its own-position operands
            stage_A(k + 1);
            stage_B(k + 2);
            asm volatile("cp.async.commit_group;" ::: "memory");
            own_load<NP>(a, pt, c + sc, vyz + syz, vxz + sxz, vxy + sxy, nc, nyz, nxz, nxy, nxt);
        }
        A.aLo = (k == 1 ? 1 : (k - 1) % 3) * ZPLANE; A.aHi = (k % 3) * ZPLANE;
        A.bLo = A.aHi; A.bHi = ((k + 1) % 3) * ZPLANE;
        vc3_sm_compute<false, NP>(a, pt, A, c, vyz, vxz, vxy, cur.rc, cur.ryz, cur.rxz, cur.rxy, cur.tyz, cur.txz, cur.txy, cur.oyz, cur.oxz, cur.oxy,
                                  cur.lam);
        if (more) cur = nxt;
        c += sc; vyz += syz; vxz += sxz; vxy += sxy;
    }
}

// global-memory variant: 3 CTAs of 256 threads per SM (≤ 80 registers, a few spilled doubles) — bound by the latency of its ≈ 280 loads per
// node; measured time falls with occupancy up to 24 warps/SM (5.75 ms per iteration at 128 registers → 5.15 ms at 80; no gain beyond)
template <bool DIAG, int NP>
__global__ void __launch_bounds__(256, 3) k_vc3_stress(const __grid_constant__ V3 a, const __grid_constant__ jr_phase_tab pt) { vc3_stress_body<DIAG, NP>(a, pt); }

// ---------------------------------------------------------------------------------------------------------------------------------
// compute_V! 3D  VelocityKernels.jl:182-242 (reads the NEW stresses, P = Pr_c, ητ)
template <bool DIAG>
__global__ void __launch_bounds__(256) k_vc3_vel(const __grid_constant__ V3 a)
{
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    int i, j;
    vc3_map_ij(nx, a.xfull_c, i, j);
    const int k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const double *P = a.P, *ett = a.etatau, *txx = a.txx_o, *tyy = a.tyy_o, *tzz = a.tzz_o, *tyz = a.tyz_o, *txz = a.txz_o, *txy = a.txy_o;
    if (a.pf_next && k + 2 <= nz) {
        // L2 prefetch of this thread's operands two planes up
        const size_t c2 = IX3(nx, ny, i, j, k + 2);
        jr_prefetch_l2(P + c2); jr_prefetch_l2(ett + c2); jr_prefetch_l2(txx + c2); jr_prefetch_l2(tyy + c2); jr_prefetch_l2(tzz + c2);
        jr_prefetch_l2(a.rgx + c2); jr_prefetch_l2(a.rgy + c2); jr_prefetch_l2(a.rgz + c2);
        jr_prefetch_l2(&XY(txy, i + 1, j + 1, k + 2)); jr_prefetch_l2(&XZ(txz, i + 1, j, k + 3)); jr_prefetch_l2(&YZ(tyz, i, j + 1, k + 3));
        jr_prefetch_l2(&VX(i + 1, j + 1, k + 3)); jr_prefetch_l2(&VY(i + 1, j + 1, k + 3)); jr_prefetch_l2(&VZ(i + 1, j + 1, k + 3));
    }
    // read-only operands go through ld.global.nc (LD(...)): the compiler may then issue all of them before the first V store
#define LD(x) __ldg(&(x))
    // the three velocities this thread updates are read BEFORE its first store (a later read-modify-write could not be hoisted above an
    // earlier store it might alias: three dependent DRAM round trips otherwise)
    const double vx_old = i <= nx - 1 ? VX(i + 1, j + 1, k + 1) : 0.0, vy_old = j <= ny - 1 ? VY(i + 1, j + 1, k + 1) : 0.0,
                 vz_old = k <= nz - 1 ? VZ(i + 1, j + 1, k + 1) : 0.0;
    const double Pc = LD(CC(P, i, j, k)), ec = LD(CC(ett, i, j, k));
    const double xy11 = LD(XY(txy, i + 1, j + 1, k)), xz11 = LD(XZ(txz, i + 1, j, k + 1)), yz11 = LD(YZ(tyz, i, j + 1, k + 1));
    if (i <= nx - 1) {
        const double R = (-LD(CC(txx, i, j, k)) + LD(CC(txx, i + 1, j, k))) * a._dx + a._dy * (xy11 - LD(XY(txy, i + 1, j, k))) +
                         a._dz * (xz11 - LD(XZ(txz, i + 1, j, k))) - (-Pc + LD(CC(P, i + 1, j, k))) * a._dx -
                         0.5 * (LD(CC(a.rgx, i, j, k)) + LD(CC(a.rgx, i + 1, j, k)));
        if (DIAG) a.Rx[IX3(nx - 1, ny, i, j, k)] = R;
        VX(i + 1, j + 1, k + 1) = vx_old + R * a.edt / (0.5 * (ec + LD(CC(ett, i + 1, j, k))));
    }
    if (j <= ny - 1) {
        const double R = a._dx * (xy11 - LD(XY(txy, i, j + 1, k))) + a._dy * (LD(CC(tyy, i, j + 1, k)) - LD(CC(tyy, i, j, k))) +
                         a._dz * (yz11 - LD(YZ(tyz, i, j + 1, k))) - (-Pc + LD(CC(P, i, j + 1, k))) * a._dy -
                         0.5 * (LD(CC(a.rgy, i, j, k)) + LD(CC(a.rgy, i, j + 1, k)));
        if (DIAG) a.Ry[IX3(nx, ny - 1, i, j, k)] = R;
        VY(i + 1, j + 1, k + 1) = vy_old + R * a.edt / (0.5 * (ec + LD(CC(ett, i, j + 1, k))));
    }
    if (k <= nz - 1) {
        const double R = a._dx * (xz11 - LD(XZ(txz, i, j, k + 1))) + a._dy * (yz11 - LD(YZ(tyz, i, j, k + 1))) +
                         (-LD(CC(tzz, i, j, k)) + LD(CC(tzz, i, j, k + 1))) * a._dz - (-Pc + LD(CC(P, i, j, k + 1))) * a._dz -
                         0.5 * (LD(CC(a.rgz, i, j, k)) + LD(CC(a.rgz, i, j, k + 1)));
        if (DIAG) a.Rz[IX3(nx, ny, i, j, k)] = R;
        VZ(i + 1, j + 1, k + 1) = vz_old + R * a.edt / (0.5 * (ec + LD(CC(ett, i, j, k + 1))));
    }
#undef LD
}

// compute_ρg! 3D stand-alone  BuoyancyForces.jl:38-60
__global__ void k_rhog3d(int nx, int ny, int nz, const __grid_constant__ jr_phase_tab pt, const double *__restrict__ ph_c, const double *__restrict__ T,
                         const double *__restrict__ Pa, double *__restrict__ rgx, double *__restrict__ rgy, double *__restrict__ rgz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const size_t c = IX3(nx, ny, i, j, k);
    const double Tc = T ? T[IX3(nx + 2, ny + 2, i + 1, j + 1, k + 1)] : 0.0, Pc = Pa ? Pa[c] : 0.0;
    const double rho = jr_ratio_density(pt, ph_c, (size_t)nx * ny * nz, c, Tc, Pc);
    if (!pt.g_scalar) { rgx[c] = rho * pt.g[0]; rgy[c] = rho * pt.g[1]; }
    rgz[c] = rho * pt.g[2];
}
// compute_viscosity_kernel! 3D (centres only)  Viscosity.jl:306-308,454-504
__global__ void k_viscosity3d(size_t n, const __grid_constant__ jr_phase_tab pt, const double *__restrict__ ph, double *__restrict__ eta, double nu,
                              double lo, double hi)
{
    const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (q < n) eta[q] = jr_clampd((1 - nu) * eta[q] + nu * jr_phase_viscosity(pt, ph, n, q), lo, hi);
}
__global__ void k_scale3(size_t n0, size_t n1, size_t n2, double *__restrict__ U0, double *__restrict__ U1, double *__restrict__ U2,
                         const double *__restrict__ V0, const double *__restrict__ V1, const double *__restrict__ V2, double f)
{
    const size_t tot = n0 + n1 + n2;
    for (size_t I = blockIdx.x * (size_t)blockDim.x + threadIdx.x; I < tot; I += (size_t)gridDim.x * blockDim.x) {
        if (I < n0) U0[I] = V0[I] * f;
        else if (I < n0 + n1) U1[I - n0] = V1[I - n0] * f;
        else U2[I - n0 - n1] = V2[I - n0 - n1] * f;
    }
}
// exit kernels ------------------------------------------------------------------------------------------------------------------
// compute_vorticity!(ωyz, ωxz, ωxy, V…, _di) over ni.+1  stress_rotation_particles.jl:32-51 (plain _d_*a at I)
__global__ void k_vorticity3d(const __grid_constant__ V3 a, double *__restrict__ wyz, double *__restrict__ wxz, double *__restrict__ wxy)
{
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx + 1 || j > ny + 1 || k > nz + 1) return;
    if (wyz && i <= nx && j <= ny + 1 && k <= nz + 1) YZ(wyz, i, j, k) = 0.5 * ((-VZ(i, j, k) + VZ(i, j + 1, k)) * a._dy - (-VY(i, j, k) + VY(i, j, k + 1)) * a._dz);
    if (wxz && i <= nx + 1 && j <= ny && k <= nz + 1) XZ(wxz, i, j, k) = 0.5 * ((-VX(i, j, k) + VX(i, j, k + 1)) * a._dz - (-VZ(i, j, k) + VZ(i + 1, j, k)) * a._dx);
    if (wxy && i <= nx + 1 && j <= ny + 1 && k <= nz) XY(wxy, i, j, k) = 0.5 * ((-VY(i, j, k) + VY(i + 1, j, k)) * a._dx - (-VX(i, j, k) + VX(i, j + 1, k)) * a._dy);
}
// shear2center! 3D  Interpolations.jl:313-323
__global__ void k_shear2center3d(int nx, int ny, int nz, double *__restrict__ yzc, double *__restrict__ xzc, double *__restrict__ xyc,
                                 const double *__restrict__ yz, const double *__restrict__ xz, const double *__restrict__ xy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const size_t c = IX3(nx, ny, i, j, k);
    yzc[c] = 0.25 * (YZ(yz, i, j, k) + YZ(yz, i, j + 1, k) + YZ(yz, i, j, k + 1) + YZ(yz, i, j + 1, k + 1));
    xzc[c] = 0.25 * (XZ(xz, i, j, k) + XZ(xz, i + 1, j, k) + XZ(xz, i, j, k + 1) + XZ(xz, i + 1, j, k + 1));
    xyc[c] = 0.25 * (XY(xy, i, j, k) + XY(xy, i + 1, j, k) + XY(xy, i, j + 1, k) + XY(xy, i + 1, j + 1, k));
}
// second_invariant_staggered of a 3D tensor with edge shear components (GeoParams; gathers MiniKernels.jl:196-204);
// mode 0: II = inv (tensor_invariant!) ; mode 1: II += inv * f (accumulate_tensor!  StressKernels.jl:394-408)
__global__ void k_inv_stag3d(int nx, int ny, int nz, double *__restrict__ II, const double *__restrict__ xx, const double *__restrict__ yy,
                             const double *__restrict__ zz, const double *__restrict__ yz, const double *__restrict__ xz, const double *__restrict__ xy,
                             int mode, double f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const size_t c = IX3(nx, ny, i, j, k);
    const double X = xx[c], Y = yy[c], Z = zz[c];
    const double a1 = YZ(yz, i, j, k), a2 = YZ(yz, i, j + 1, k), a3 = YZ(yz, i, j, k + 1), a4 = YZ(yz, i, j + 1, k + 1);
    const double b1 = XZ(xz, i, j, k), b2 = XZ(xz, i + 1, j, k), b3 = XZ(xz, i, j, k + 1), b4 = XZ(xz, i + 1, j, k + 1);
    const double c1 = XY(xy, i, j, k), c2 = XY(xy, i + 1, j, k), c3 = XY(xy, i, j + 1, k), c4 = XY(xy, i + 1, j + 1, k);
    const double yz2 = (((a1 * a1 + a2 * a2) + a3 * a3) + a4 * a4) / 4, xz2 = (((b1 * b1 + b2 * b2) + b3 * b3) + b4 * b4) / 4,
                 xy2 = (((c1 * c1 + c2 * c2) + c3 * c3) + c4 * c4) / 4;
    const double v = sqrt(0.5 * (X * X + Y * Y + Z * Z) + yz2 + xz2 + xy2);
    if (mode == 0) II[c] = v;
    else II[c] += v * f;
}
__global__ void k_axpy3(size_t n, double *__restrict__ A, const double *__restrict__ B, double f)
{
    const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (q < n) A[q] += f * B[q];
}

// ---------------------------------------------------------------------------------------------------------------------------------
// host drivers
enum { T_xx, T_yy, T_zz, T_yz, T_xz, T_xy, T_COUNT };
struct Plan3 {
    int nx, ny, nz;
    size_t nc, nyz, nxz, nxy;
    double *tau[2][T_COUNT];   // τ ping-pong sets (set 0 = the caller's arrays)
    double *eta[2];            // η ping-pong (set 0 = the caller's array)
    size_t tbytes[T_COUNT];
    V3 k;
    jr_phase_tab pt;
    bool multi;
    int32_t n[3], fs[6], ns[6], pe[6];
};

static int check3d_vc(const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *in)
{
    JR_REQUIRE(s && o, JR_ERR_ARG, "null fields/opts");
    JR_REQUIRE(s->ndim == 3, JR_ERR_SHAPE, "3D solver called with ndim=%d", s->ndim);
    JR_REQUIRE(s->n[0] >= 3 && s->n[1] >= 3 && s->n[2] >= 3, JR_ERR_SHAPE, "grid must be at least 3 cells per dimension");
    JR_REQUIRE(o->nout >= 1, JR_ERR_ARG, "nout must be >= 1");
    static const int req[] = {JR_F_P, JR_F_P0, JR_F_divV, JR_F_Q, JR_F_Vx, JR_F_Vy, JR_F_Vz, JR_F_txx, JR_F_tyy, JR_F_tzz, JR_F_tyz, JR_F_txz, JR_F_txy,
                              JR_F_tyz_c, JR_F_txz_c, JR_F_txy_c, JR_F_txx_o, JR_F_tyy_o, JR_F_tzz_o, JR_F_tyz_o, JR_F_txz_o, JR_F_txy_o, JR_F_tyz_o_c,
                              JR_F_txz_o_c, JR_F_txy_o_c, JR_F_exx, JR_F_eyy, JR_F_ezz, JR_F_eyz, JR_F_exz, JR_F_exy, JR_F_pxx, JR_F_pyy, JR_F_pzz, JR_F_pyz,
                              JR_F_pxz, JR_F_pxy, JR_F_tII, JR_F_eta_vep, JR_F_e_vol_pl, JR_F_EII_pl, JR_F_EVol_pl, JR_F_eta, JR_F_etatau, JR_F_Rx, JR_F_Ry,
                              JR_F_Rz, JR_F_RP, JR_F_rhogx, JR_F_rhogy, JR_F_rhogz};
    for (int q : req) JR_REQUIRE(s->f[q] != nullptr, JR_ERR_SHAPE, "required field '%s' is NULL", jr_field_name(q));
    JR_REQUIRE(in && in->ph_center && in->ph_xy && in->ph_yz && in->ph_xz, JR_ERR_SHAPE, "3D-VC needs phase ratios at centres and at the xy, yz, xz edges");
    return JR_OK;
}

static inline dim3 grid3(int nx, int ny, int nz) { return dim3((nx + 31) / 32, (ny + 7) / 8, nz); }
// remainder-column packing (vc3_map_ij): few (≤ 8) columns beyond the last full 32-wide tile are packed densely into one extra block column
static inline int xfull_of(int n0)
{
    static const bool on = !(getenv("JRB200_VC3_PACK") && atoi(getenv("JRB200_VC3_PACK")) == 0);
    const int rem = n0 % 32;
    return (on && rem > 0 && rem <= 8 && n0 >= 32) ? n0 / 32 : (n0 + 31) / 32;
}
static inline dim3 grid3p(int n0, int n1, int n2, int xfull) { return dim3(xfull * 32 < n0 ? xfull + 1 : xfull, (n1 + 7) / 8, n2); }
static const dim3 BLK3(32, 8, 1);

static int plan3_begin(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *in, Plan3 *p)
{
    memset(p, 0, sizeof(*p));
    const int nx = p->nx = s->n[0], ny = p->ny = s->n[1], nz = p->nz = s->n[2];
    p->nc = (size_t)nx * ny * nz; p->nyz = (size_t)nx * (ny + 1) * (nz + 1); p->nxz = (size_t)(nx + 1) * ny * (nz + 1); p->nxy = (size_t)(nx + 1) * (ny + 1) * nz;
    const size_t b[T_COUNT] = {p->nc * 8, p->nc * 8, p->nc * 8, p->nyz * 8, p->nxz * 8, p->nxy * 8};
    size_t off[T_COUNT], tot = 0;
    for (int q = 0; q < T_COUNT; q++) { p->tbytes[q] = b[q]; off[q] = tot; tot += (b[q] + 255) & ~(size_t)255; }
    const size_t o_eta = tot; tot += (p->nc * 8 + 255) & ~(size_t)255;
    const size_t o_th = tot; tot += (p->nc * 8 + 255) & ~(size_t)255;
    const size_t o_lam = tot; tot += (p->nc * 8 + 255) & ~(size_t)255;
    const size_t o_lyz = tot; tot += (p->nyz * 8 + 255) & ~(size_t)255;
    const size_t o_lxz = tot; tot += (p->nxz * 8 + 255) & ~(size_t)255;
    const size_t o_lxy = tot; tot += (p->nxy * 8 + 255) & ~(size_t)255;
    void *base = nullptr;
    int st = jr_ctx_scratch(ctx, "stokes3d_vc", tot, &base);
    if (st) return st;
    char *B = (char *)base;
    double *user[T_COUNT] = {F(txx), F(tyy), F(tzz), F(tyz), F(txz), F(txy)};
    for (int q = 0; q < T_COUNT; q++) { p->tau[0][q] = user[q]; p->tau[1][q] = (double *)(B + off[q]); }
    p->eta[0] = F(eta); p->eta[1] = (double *)(B + o_eta);
    for (int q = 0; q < 3; q++) p->n[q] = s->n[q];
    for (int q = 0; q < 6; q++) { p->fs[q] = o->free_slip[q]; p->ns[q] = o->no_slip[q]; p->pe[q] = o->periodic[q]; }
    p->multi = ctx->comm && ctx->comm->active;
    if ((st = jr_make_phase_tab(in, &p->pt))) return st;
    JR_REQUIRE(!p->pt.any_soft, JR_ERR_UNSUPPORTED, "cohesion softening is supported by the 2D multiphase solve only (the 3D-VC kernels do not carry EII to the edges)");
    V3 &k = p->k;
    k.nx = nx; k.ny = ny; k.nz = nz;
    k._dx = o->_di[0]; k._dy = o->_di[1]; k._dz = o->_di[2]; k.dt = o->dt; k.r = o->r; k.th = o->theta_dtau; k.edt = o->eta_dtau;
    k.rel = o->lambda_relaxation; k.nu = o->viscosity_relaxation; k.cut_lo = o->visc_cutoff_lo; k.cut_hi = o->visc_cutoff_hi;
    {
        const char *e = getenv("JRB200_VC3_PREFETCH");
        k.pf_next = (e && atoi(e) == 0) ? 0 : 1;
    }
    k.xfull_c = xfull_of(k.nx); k.xfull_n = xfull_of(k.nx + 1);
    k.Vx = F(Vx); k.Vy = F(Vy); k.Vz = F(Vz); k.theta = (double *)(B + o_th); k.P = F(P); k.P0 = F(P0); k.Q = F(Q); k.etatau = F(etatau);
    k.exx = F(exx); k.eyy = F(eyy); k.ezz = F(ezz); k.eyz = F(eyz); k.exz = F(exz); k.exy = F(exy);
    k.tyzc = F(tyz_c); k.txzc = F(txz_c); k.txyc = F(txy_c);
    k.oxx = F(txx_o); k.oyy = F(tyy_o); k.ozz = F(tzz_o); k.oyz = F(tyz_o); k.oxz = F(txz_o); k.oxy = F(txy_o);
    k.oyzc = F(tyz_o_c); k.oxzc = F(txz_o_c); k.oxyc = F(txy_o_c);
    k.lam = (double *)(B + o_lam); k.lamyz = (double *)(B + o_lyz); k.lamxz = (double *)(B + o_lxz); k.lamxy = (double *)(B + o_lxy);
    k.rgx = F(rhogx); k.rgy = F(rhogy); k.rgz = F(rhogz); k.T = F(T); k.Pargs = F(Pargs); k.dTargs = F(dTargs); k.dT_ghosted = o->dT_ghosted;
    k.ph_c = in->ph_center; k.ph_xy = in->ph_xy; k.ph_yz = in->ph_yz; k.ph_xz = in->ph_xz;
    k.divV = F(divV); k.RP = F(RP); k.pxx = F(pxx); k.pyy = F(pyy); k.pzz = F(pzz); k.pyz = F(pyz); k.pxz = F(pxz); k.pxy = F(pxy);
    k.tII = F(tII); k.eta_vep = F(eta_vep); k.e_vol_pl = F(e_vol_pl); k.Rx = F(Rx); k.Ry = F(Ry); k.Rz = F(Rz);
    return JR_OK;
}

// pre-loop  Stokes3D.jl:493-509
static int pre_VC3(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, Plan3 *p)
{
    cudaStream_t st = ctx->stream;
    const V3 &k = p->k;
    JR_CUDA(cudaMemcpyAsync(F(P0), F(P), p->nc * 8, cudaMemcpyDeviceToDevice, st));      // @copy stokes.P0 stokes.P   :493
    JR_CUDA(cudaMemcpyAsync(k.theta, F(P), p->nc * 8, cudaMemcpyDeviceToDevice, st));    // θ = deepcopy(stokes.P)     :494
    JR_CUDA(cudaMemsetAsync(k.lam, 0, p->nc * 8, st));                                   // λ, λv_* = 0                :495-498
    JR_CUDA(cudaMemsetAsync(k.lamyz, 0, p->nyz * 8, st));
    JR_CUDA(cudaMemsetAsync(k.lamxz, 0, p->nxz * 8, st));
    JR_CUDA(cudaMemsetAsync(k.lamxy, 0, p->nxy * 8, st));
    JR_CUDA(cudaMemcpyAsync(F(etatau), F(eta), p->nc * 8, cudaMemcpyDeviceToDevice, st));  // ητ = deepcopy(η)         :502
    k_rhog3d<<<grid3(p->nx, p->ny, p->nz), BLK3, 0, st>>>(p->nx, p->ny, p->nz, p->pt, k.ph_c, F(T), F(Pargs), F(rhogx), F(rhogy), F(rhogz));   // compute_ρg!  :505
    // compute_viscosity! (εII form, relaxation 1)  :506 — independent of εII for the LinearViscous subset
    k_viscosity3d<<<(unsigned)((p->nc + 255) / 256), 256, 0, st>>>(p->nc, p->pt, k.ph_c, F(eta), 1.0, o->visc_cutoff_lo, o->visc_cutoff_hi);
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

template <int NP>
static void launch_prep_np(bool diag, bool maxloc, dim3 grd, cudaStream_t st, const V3 &k, const jr_phase_tab &pt)
{
    if (k.dTargs) {  // thermal-stress pressure form (args.ΔT): its own instantiations, the common path keeps its register budget
        if (diag) { if (maxloc) k_vc3_prep<true, true, NP, true><<<grd, BLK3, 0, st>>>(k, pt); else k_vc3_prep<true, false, NP, true><<<grd, BLK3, 0, st>>>(k, pt); }
        else { if (maxloc) k_vc3_prep<false, true, NP, true><<<grd, BLK3, 0, st>>>(k, pt); else k_vc3_prep<false, false, NP, true><<<grd, BLK3, 0, st>>>(k, pt); }
        return;
    }
    if (diag) { if (maxloc) k_vc3_prep<true, true, NP><<<grd, BLK3, 0, st>>>(k, pt); else k_vc3_prep<true, false, NP><<<grd, BLK3, 0, st>>>(k, pt); }
    else { if (maxloc) k_vc3_prep<false, true, NP><<<grd, BLK3, 0, st>>>(k, pt); else k_vc3_prep<false, false, NP><<<grd, BLK3, 0, st>>>(k, pt); }
}
static void launch_prep(bool diag, bool maxloc, int nphase, dim3 grd, cudaStream_t st, const V3 &k, const jr_phase_tab &pt)
{
    if (nphase <= 1) launch_prep_np<1>(diag, maxloc, grd, st, k, pt);
    else if (nphase == 2) launch_prep_np<2>(diag, maxloc, grd, st, k, pt);
    else if (nphase == 3) launch_prep_np<3>(diag, maxloc, grd, st, k, pt);
    else if (nphase == 4) launch_prep_np<4>(diag, maxloc, grd, st, k, pt);
    else launch_prep_np<JR_MAX_PHASES>(diag, maxloc, grd, st, k, pt);
}

// z-marching stress kernel (k_vc3_stress_zm): rows per CTA (0 = off: the per-plane staged kernel everywhere), planes per CTA
static int zm_rows()
{
    // default OFF: measured slower than the per-plane kernel (257^3, three phases: 3.80 ms per iteration with 8 rows, 4.81 ms with 12, against
    // 3.30 ms) — the three-plane ring leaves room for one CTA per SM (8 / 12 warps instead of 16), and the kernel is bound by the latency of its
    // FP64 / shared-memory chains, not by DRAM (profiles/r02_vc3_zmarch.md).  Opt-in: JRB200_VC3_ZM=8 | 12.  (Read per call: the tests switch
    // variants inside one process.)
    int v = 0;
    if (const char *e = getenv("JRB200_VC3_ZM")) v = atoi(e);
    if (v != 0 && v != 8 && v != 12) v = 0;
    return v;
}
static int zm_chunk_of(int planes, int ctas_per_layer, int sm_count)
{
    if (const char *e = getenv("JRB200_VC3_ZM_CHUNK")) { const int c = atoi(e); if (c >= 1) return c; }
    // one CTA per SM: pick the number of z-chunks (≥ 16 planes each, so the two-plane prologue stays small) whose CTA count fills whole waves best
    int best_n = 1;
    double best = -1.0;
    for (int n = 1; n <= planes / 16 || n == 1; n++) {
        const long tot = (long)n * ctas_per_layer;
        const long waves = (tot + sm_count - 1) / sm_count;
        double eff = (double)tot / (double)(waves * sm_count);
        if (waves < 4) eff *= 0.25 * waves;   // too few waves: the pipeline fill of every CTA is exposed
        if (eff > best + 1e-9) { best = eff; best_n = n; }
    }
    return (planes + best_n - 1) / best_n;
}

template <int NP, int TYZ>
static void launch_zm(dim3 g, cudaStream_t st, const V3 &k, const jr_phase_tab &pt)
{
    constexpr size_t smem = (size_t)SL_COUNT * 3 * 33 * (TYZ + 1) * sizeof(double);
    static bool attr = false;   // per instantiation
    if (!attr) {
        cudaFuncSetAttribute(k_vc3_stress_zm<NP, TYZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    k_vc3_stress_zm<NP, TYZ><<<g, dim3(32, TYZ, 1), smem, st>>>(k, pt);
}

// returns the number of launches
template <int NP>
static int launch_stress_np(bool diag, dim3 grd, cudaStream_t st, V3 k, const jr_phase_tab &pt, int sm_count)
{
    static const bool use_sm = !(getenv("JRB200_VC_STRESS_GLOBAL") && atoi(getenv("JRB200_VC_STRESS_GLOBAL")));
    k.bo_x = k.bo_y = k.bo_z = 0; k.zm_i = k.zm_j = k.zm_k = 0; k.zm_chunk = 1;
    if (use_sm) {
        const size_t smem = (size_t)SL_COUNT * STILE * sizeof(double);
        static bool attr = false;   // per instantiation
        if (!attr) {
            cudaFuncSetAttribute(k_vc3_stress_sm<true, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_vc3_stress_sm<false, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr = true;
        }
        const dim3 blk(32, TYS, 1), g2(grd.x, (grd.y * 8 + TYS - 1) / TYS, grd.z);
        const int rows = zm_rows();
        const int zi = ((k.nx - 1) / 32) * 32, zj = rows ? ((k.ny - 1) / rows) * rows : 0, zk = k.nz - 1;
        if (!diag && rows && TYS == 8 && zi >= 32 && zj >= rows && zk >= 2) {
            // interior: z-marching CTAs; rim (high-side boundary columns / rows / planes): the per-plane kernel on three slabs of its grid
            k.zm_i = zi; k.zm_j = zj; k.zm_k = zk;
            k.zm_chunk = zm_chunk_of(zk, (zi / 32) * (zj / rows), sm_count);
            const dim3 gz(zi / 32, zj / rows, (zk + k.zm_chunk - 1) / k.zm_chunk);
            if (rows == 12) launch_zm<NP, 12>(gz, st, k, pt);
            else launch_zm<NP, 8>(gz, st, k, pt);
            int n = 1;
            const int by0 = zj / 8, bx0 = zi / 32;   // first block row / column of the per-plane grid that is not entirely inside the region
            V3 r = k;
            r.bo_z = zk;   // (a) the planes above the region
            if ((int)g2.z > zk) { k_vc3_stress_sm<false, NP><<<dim3(g2.x, g2.y, g2.z - zk), blk, smem, st>>>(r, pt); n++; }
            r.bo_z = 0; r.bo_y = by0;   // (b) the block rows beyond the region
            if ((int)g2.y > by0) { k_vc3_stress_sm<false, NP><<<dim3(g2.x, g2.y - by0, zk), blk, smem, st>>>(r, pt); n++; }
            r.bo_y = 0; r.bo_x = bx0;   // (c) the block columns beyond the region, below (b)
            if ((int)g2.x > bx0 && by0 > 0) { k_vc3_stress_sm<false, NP><<<dim3(g2.x - bx0, by0, zk), blk, smem, st>>>(r, pt); n++; }
            return n;
        }
        if (diag) k_vc3_stress_sm<true, NP><<<g2, blk, smem, st>>>(k, pt);
        else k_vc3_stress_sm<false, NP><<<g2, blk, smem, st>>>(k, pt);
        return 1;
    }
    if (diag) k_vc3_stress<true, NP><<<grd, BLK3, 0, st>>>(k, pt);
    else k_vc3_stress<false, NP><<<grd, BLK3, 0, st>>>(k, pt);
    return 1;
}
static int launch_stress(bool diag, int nphase, dim3 grd, cudaStream_t st, const V3 &k, const jr_phase_tab &pt, int sm_count)
{
    if (nphase <= 1) return launch_stress_np<1>(diag, grd, st, k, pt, sm_count);
    else if (nphase == 2) return launch_stress_np<2>(diag, grd, st, k, pt, sm_count);
    else if (nphase == 3) return launch_stress_np<3>(diag, grd, st, k, pt, sm_count);
    else if (nphase == 4) return launch_stress_np<4>(diag, grd, st, k, pt, sm_count);
    return launch_stress_np<JR_MAX_PHASES>(diag, grd, st, k, pt, sm_count);
}

// one PT iteration: τ set (it & 1) → set ((it + 1) & 1), η likewise
static int plan3_iter(jr_context *ctx, Plan3 *p, int64_t it, bool diag, const jr_fields *s, const jr_stokes_opts *o)
{
    V3 k = p->k;
    double *const *I = p->tau[it & 1], *const *O = p->tau[(it + 1) & 1];
    k.txx_i = I[T_xx]; k.tyy_i = I[T_yy]; k.tzz_i = I[T_zz]; k.tyz_i = I[T_yz]; k.txz_i = I[T_xz]; k.txy_i = I[T_xy];
    k.txx_o = O[T_xx]; k.tyy_o = O[T_yy]; k.tzz_o = O[T_zz]; k.tyz_o = O[T_yz]; k.txz_o = O[T_xz]; k.txy_o = O[T_xy];
    k.eta_i = p->eta[it & 1]; k.eta_o = p->eta[(it + 1) & 1];
    cudaStream_t st = ctx->stream;
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    int rc;
    if (p->multi) {  // compute_maxloc!(ητ, η); update_halo!(ητ)  :514-515
        const int32_t w3[3] = {1, 1, 1};
        if ((rc = jr_launch_maxloc3d(ctx, k.etatau, k.eta_i, p->n, w3))) return rc;
        const jr_harr H = jr_harr_dense(k.etatau, p->n, p->n);
        if ((rc = jr_comm_halo(ctx, &H, 1))) return rc;
    }
    launch_prep(diag, !p->multi, p->pt.n, grid3p(nx, ny, nz, k.xfull_c), st, k, p->pt);
    ctx->launches += 1 + launch_stress(diag, p->pt.n, grid3p(nx + 1, ny + 1, nz + 1, k.xfull_n), st, k, p->pt, ctx->sm_count);
    JR_CHECK_LAUNCH();
    if (p->multi) {  // update_halo!(τyz); update_halo!(τxz); update_halo!(τxy)  :578-580
        const int32_t eyz[3] = {nx, ny + 1, nz + 1}, exz[3] = {nx + 1, ny, nz + 1}, exy[3] = {nx + 1, ny + 1, nz};
        const jr_harr H[3] = {jr_harr_dense(k.tyz_o, eyz, p->n), jr_harr_dense(k.txz_o, exz, p->n), jr_harr_dense(k.txy_o, exy, p->n)};
        if ((rc = jr_comm_halo(ctx, H, 3))) return rc;
    }
    if (diag) k_vc3_vel<true><<<grid3p(nx, ny, nz, k.xfull_c), BLK3, 0, st>>>(k);
    else k_vc3_vel<false><<<grid3p(nx, ny, nz, k.xfull_c), BLK3, 0, st>>>(k);
    ctx->launches++;
    const size_t nVx = (size_t)(nx + 1) * (ny + 2) * (nz + 2), nVy = (size_t)(nx + 2) * (ny + 1) * (nz + 2), nVz = (size_t)(nx + 2) * (ny + 2) * (nz + 1);
    if (diag && F(Ux) && F(Uy) && F(Uz)) {  // velocity2displacement!(stokes, dt) BEFORE flow_bcs!  :594
        k_scale3<<<ctx->sm_count * 8, 256, 0, st>>>(nVx, nVy, nVz, F(Ux), F(Uy), F(Uz), k.Vx, k.Vy, k.Vz, o->dt);
        ctx->launches++;
    }
    JR_CHECK_LAUNCH();
    if ((rc = jr_launch_flow_bcs3d(ctx, k.Vx, k.Vy, k.Vz, p->n, p->fs, p->ns, p->pe))) return rc;   // flow_bcs!  :595
    if (p->multi) {  // update_halo!(@velocity(stokes)...)  :596
        const int32_t eVx[3] = {nx + 1, ny + 2, nz + 2}, eVy[3] = {nx + 2, ny + 1, nz + 2}, eVz[3] = {nx + 2, ny + 2, nz + 1};
        const jr_harr H[3] = {jr_harr_dense(k.Vx, eVx, p->n), jr_harr_dense(k.Vy, eVy, p->n), jr_harr_dense(k.Vz, eVz, p->n)};
        if ((rc = jr_comm_halo(ctx, H, 3))) return rc;
    }
    return JR_OK;
}

// bring the final τ and η (sets niter & 1) back into the caller's arrays; expose the solver-local λ
static int plan3_finish(jr_context *ctx, const jr_fields *s, Plan3 *p, int64_t niter)
{
    cudaStream_t st = ctx->stream;
    if (niter & 1) {
        for (int q = 0; q < T_COUNT; q++) JR_CUDA(cudaMemcpyAsync(p->tau[0][q], p->tau[1][q], p->tbytes[q], cudaMemcpyDeviceToDevice, st));
        JR_CUDA(cudaMemcpyAsync(p->eta[0], p->eta[1], p->nc * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (F(lam)) JR_CUDA(cudaMemcpyAsync(F(lam), p->k.lam, p->nc * 8, cudaMemcpyDeviceToDevice, st));
    return JR_OK;
}

// norms  Stokes3D.jl:601-611 (quirk Q4: ‖R‖₂ / ((nx_g−1)(ny_g−1)(nz_g−1)), RP by the LOCAL length)
static int norms3d_vc(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, Plan3 *p, double e[4])
{
    void *slots_v = nullptr;
    int st = jr_ctx_scratch(ctx, "norm_slots", 16 * sizeof(double), &slots_v);
    if (st) return st;
    double *slots = (double *)slots_v;
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    const int32_t nRx[3] = {nx - 1, ny, nz}, nRy[3] = {nx, ny - 1, nz}, nRz[3] = {nx, ny, nz - 1}, nP[3] = {nx, ny, nz};
    if ((st = jr_launch_sumsq(ctx, F(Rx), nRx, 1, slots + 0))) return st;
    if ((st = jr_launch_sumsq(ctx, F(Ry), nRy, 1, slots + 1))) return st;
    if ((st = jr_launch_sumsq(ctx, F(Rz), nRz, 1, slots + 2))) return st;
    if ((st = jr_launch_sumsq(ctx, F(RP), nP, 0, slots + 3))) return st;
    if ((st = jr_comm_allreduce_dev(ctx, slots, 4, 0))) return st;   // norm_mpi: Allreduce(sum) of the local sums of squares
    JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, slots, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    const double ng = (double)(o->n_g[0] - 1) * (o->n_g[1] - 1) * (o->n_g[2] - 1);
    e[0] = sqrt(ctx->h_pinned[0]) / ng;
    e[1] = sqrt(ctx->h_pinned[1]) / ng;
    e[2] = sqrt(ctx->h_pinned[2]) / ng;
    e[3] = sqrt(ctx->h_pinned[3]) / ((double)nx * ny * nz);
    return JR_OK;
}

// exit kernels  Stokes3D.jl:641-655
static int post_VC3(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, Plan3 *p)
{
    const int nx = p->nx, ny = p->ny, nz = p->nz;
    cudaStream_t st = ctx->stream;
    if (F(wyz) || F(wxz) || F(wxy)) { k_vorticity3d<<<grid3(nx + 1, ny + 1, nz + 1), BLK3, 0, st>>>(p->k, F(wyz), F(wxz), F(wxy)); ctx->launches++; }
    if (F(eyz_c) && F(exz_c) && F(exy_c)) { k_shear2center3d<<<grid3(nx, ny, nz), BLK3, 0, st>>>(nx, ny, nz, F(eyz_c), F(exz_c), F(exy_c), F(eyz), F(exz), F(exy)); ctx->launches++; }
    if (F(pyz_c) && F(pxz_c) && F(pxy_c)) { k_shear2center3d<<<grid3(nx, ny, nz), BLK3, 0, st>>>(nx, ny, nz, F(pyz_c), F(pxz_c), F(pxy_c), F(pyz), F(pxz), F(pxy)); ctx->launches++; }
    if (F(dyz_c) && F(dxz_c) && F(dxy_c) && F(dyz) && F(dxz) && F(dxy)) {
        k_shear2center3d<<<grid3(nx, ny, nz), BLK3, 0, st>>>(nx, ny, nz, F(dyz_c), F(dxz_c), F(dxy_c), F(dyz), F(dxz), F(dxy));
        ctx->launches++;
    }
    k_inv_stag3d<<<grid3(nx, ny, nz), BLK3, 0, st>>>(nx, ny, nz, F(EII_pl), F(pxx), F(pyy), F(pzz), F(pyz), F(pxz), F(pxy), 1, o->dt);   // accumulate_tensor!
    k_axpy3<<<(unsigned)((p->nc + 255) / 256), 256, 0, st>>>(p->nc, F(EVol_pl), F(e_vol_pl), o->dt);                                      // accumulate_vol!
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    // multi_copy! τ → τ_o (edge set over ni.+1, centre set over ni)
    double *dst[9] = {F(txx_o), F(tyy_o), F(tzz_o), F(tyz_o), F(txz_o), F(txy_o), F(tyz_o_c), F(txz_o_c), F(txy_o_c)};
    double *src[9] = {F(txx), F(tyy), F(tzz), F(tyz), F(txz), F(txy), F(tyz_c), F(txz_c), F(txy_c)};
    const size_t by[9] = {p->nc * 8, p->nc * 8, p->nc * 8, p->nyz * 8, p->nxz * 8, p->nxy * 8, p->nc * 8, p->nc * 8, p->nc * 8};
    for (int q = 0; q < 9; q++) JR_CUDA(cudaMemcpyAsync(dst[q], src[q], by[q], cudaMemcpyDeviceToDevice, st));
    return JR_OK;
}

static void fill_result3(jr_context *ctx, jr_stokes_result *res, int64_t iter, int64_t cont, double err)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    res->iter = iter; res->nhist = cont; res->err = err;
    res->time_s = ms * 1e-3;
    res->kernel_launches = ctx->launches;
}

extern "C" {

int jr_stokes3d_iterate_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, int64_t niter, int finish,
                           jr_stokes_result *res)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    int st = check3d_vc(s, o, vc);
    if (st) return st;
    JR_CUDA(cudaSetDevice(ctx->device));
    Plan3 p;
    ctx->launches = 0;
    if ((st = plan3_begin(ctx, s, o, vc, &p))) return st;
    if ((st = pre_VC3(ctx, s, o, &p))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int64_t it = 0; it < niter; it++)
        if ((st = plan3_iter(ctx, &p, it, (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) || it == niter - 1, s, o))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if ((st = plan3_finish(ctx, s, &p, niter))) return st;
    if (finish && (st = post_VC3(ctx, s, o, &p))) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res) fill_result3(ctx, res, niter, 0, NAN);
    return JR_OK;
}

int jr_stokes3d_solve_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, jr_stokes_result *res)
{
    JR_REQUIRE(ctx && res, JR_ERR_ARG, "null context/result");
    int st = check3d_vc(s, o, vc);
    if (st) return st;
    JR_REQUIRE(res->err_evo1 && res->err_evo2 && res->norm_Rx && res->norm_Ry && res->norm_Rz && res->norm_divV, JR_ERR_ARG,
               "result history arrays must be provided");
    JR_CUDA(cudaSetDevice(ctx->device));
    Plan3 p;
    ctx->launches = 0;
    if ((st = plan3_begin(ctx, s, o, vc, &p))) return st;
    if ((st = pre_VC3(ctx, s, o, &p))) return st;
    double err_it1 = 1.0, err = INFINITY;
    int64_t iter = 0, cont = 0;
    int status = JR_OK;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    while (iter < 2 || (((err / err_it1) > o->eps_rel && err > o->eps_abs) && iter <= o->iterMax)) {  // Stokes3D.jl:511
        const int64_t next = iter + 1;
        // the loop can only end right after a sample (or at iter = 2 / beyond iterMax): diagnostics written there reproduce the final state
        const bool diag = (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) || (next % o->nout == 0) || next > o->iterMax || next <= 2;
        if ((st = plan3_iter(ctx, &p, iter, diag, s, o))) return st;
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            double e[4];
            if ((st = norms3d_vc(ctx, s, o, &p, e))) return st;
            res->norm_Rx[cont] = e[0]; res->norm_Ry[cont] = e[1]; res->norm_Rz[cont] = e[2]; res->norm_divV[cont] = e[3];
            err = fmax(fmax(e[0], e[1]), fmax(e[2], e[3]));
            if (std::isnan(e[0]) || std::isnan(e[1]) || std::isnan(e[2]) || std::isnan(e[3])) err = NAN;
            res->err_evo1[cont] = err; res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), fmax(res->norm_Rz[0], res->norm_divV[0]));
            if (std::isnan(err)) { status = JR_ERR_NAN; break; }  // isnan(err) && error("NaN(s)")  Stokes3D.jl:631
        }
    }
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if ((st = plan3_finish(ctx, s, &p, iter))) return st;
    if (status == JR_OK && (st = post_VC3(ctx, s, o, &p))) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    fill_result3(ctx, res, iter, cont, err);
    if (status) jr_set_error("NaN(s)");
    return status;
}

int jr_compute_viscosity3d(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, double nu)
{
    JR_REQUIRE(ctx && s && o && vc && F(eta) && vc->ph_center, JR_ERR_ARG, "jr_compute_viscosity3d: null argument");
    jr_phase_tab pt;
    int st = jr_make_phase_tab(vc, &pt);
    if (st) return st;
    const size_t nc = (size_t)s->n[0] * s->n[1] * s->n[2];
    k_viscosity3d<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(nc, pt, vc->ph_center, F(eta), nu, o->visc_cutoff_lo, o->visc_cutoff_hi);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_compute_rhog3d(jr_context *ctx, const jr_fields *s, const jr_vc_inputs *vc)
{
    JR_REQUIRE(ctx && s && vc && F(rhogx) && F(rhogy) && F(rhogz) && vc->ph_center, JR_ERR_ARG, "jr_compute_rhog3d: null argument");
    jr_phase_tab pt;
    int st = jr_make_phase_tab(vc, &pt);
    if (st) return st;
    k_rhog3d<<<grid3(s->n[0], s->n[1], s->n[2]), BLK3, 0, ctx->stream>>>(s->n[0], s->n[1], s->n[2], pt, vc->ph_center, F(T), F(Pargs), F(rhogx), F(rhogy), F(rhogz));
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_tensor_invariant3d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *zz, const double *yz, const double *xz,
                          const double *xy, const int32_t n[3])
{
    JR_REQUIRE(ctx && II && xx && yy && zz && yz && xz && xy && n, JR_ERR_ARG, "jr_tensor_invariant3d: null argument");
    k_inv_stag3d<<<grid3(n[0], n[1], n[2]), BLK3, 0, ctx->stream>>>(n[0], n[1], n[2], II, xx, yy, zz, yz, xz, xy, 0, 0.0);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_shear2center3d(jr_context *ctx, double *yz_c, double *xz_c, double *xy_c, const double *yz, const double *xz, const double *xy, const int32_t n[3])
{
    JR_REQUIRE(ctx && yz_c && xz_c && xy_c && yz && xz && xy && n, JR_ERR_ARG, "jr_shear2center3d: null argument");
    k_shear2center3d<<<grid3(n[0], n[1], n[2]), BLK3, 0, ctx->stream>>>(n[0], n[1], n[2], yz_c, xz_c, xy_c, yz, xz, xy);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_accumulate_tensor3d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *zz, const double *yz, const double *xz,
                           const double *xy, const int32_t n[3], double dt)
{
    JR_REQUIRE(ctx && II && xx && yy && zz && yz && xz && xy && n, JR_ERR_ARG, "jr_accumulate_tensor3d: null argument");
    k_inv_stag3d<<<grid3(n[0], n[1], n[2]), BLK3, 0, ctx->stream>>>(n[0], n[1], n[2], II, xx, yy, zz, yz, xz, xy, 1, dt);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_accumulate_vol(jr_context *ctx, double *EVol_pl, const double *e_vol_pl, size_t count, double dt)
{
    JR_REQUIRE(ctx && EVol_pl && e_vol_pl, JR_ERR_ARG, "jr_accumulate_vol: null argument");
    k_axpy3<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(count, EVol_pl, e_vol_pl, dt);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

}  // extern "C"
