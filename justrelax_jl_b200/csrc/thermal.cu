// thermal.cu — heatdiffusion_PT! (2D + 3D, array form and rheology form) of libjrb200, sm_100a.
//
// Replaces the while-loop of `_heatdiffusion_PT!` (src/thermal_diffusion/DiffusionPT_solver.jl:34-149, 181-305) and its
// kernels compute_flux! / update_T! / check_res! / update_ΔT! / adiabatic_heating / compute_pt_thermal_arrays!
// (DiffusionPT_kernels.jl, DiffusionPT_coefficients.jl:105-151) and thermal_bcs! (BoundaryConditions.jl:39-54).
//
// Kernels (all x-contiguous / coalesced, one thread per node, FP64, -fmad=false so the operation order is the
// reference's):
//   k_th_pt      per-cell θr_dτ, dτ_ρ from ρCp(T,P,phases), K(phases)          (rheology form with phase ratios: every iteration)
//   k_th_flux    qT ← (qT·θ̄ + q)/(1 + θ̄), q = −K̄ ∂T; all 2–3 directions of a node in one thread; the raw flux qT2 is only
//                written on the iterations whose residual is sampled (the reference rewrites it every iteration and reads
//                it only in check_res!)
//   k_th_update  T ← (dτ_ρ(−∇·qT + Told ρCp/dt + sources) + T)/(1 + dτ_ρ ρCp/dt), Dirichlet mask branch, and on sampled
//                iterations the residual ResT with its squared-norm partial (warp shuffle → one slot per block)
//   k_th_bc      ghost layer: constant value → no flux → periodic as ONE race-free gather over the six faces
// HBM traffic per cell and iteration (3D, array form): flux reads T,K,θ,q×3 + writes q×3 = 9 passes; update reads
// q×3,T,Told,H,Hs,ρCp,dτ_ρ + writes T = 10 passes → 152 B/cell against A_eff = (2·4 + 7)·8 = 120 B/cell.
#include "common.cuh"
#include "comm.cuh"
#include "tma.cuh"

struct ThDims {
    int nd, nx, ny, nz, gx, gy, gz;
};
__host__ __device__ inline size_t th_ti(const ThDims &d, int i, int j, int k) { return ((size_t)k * d.gy + j) * d.gx + i; }
__host__ __device__ inline size_t th_ci(const ThDims &d, int i, int j, int k) { return ((size_t)k * d.ny + j) * d.nx + i; }

#define TH_MAX_PHASES 8
struct ThTable {
    int nphase;
    jr_thermal_phase p[TH_MAX_PHASES];
};

struct ThArgs {
    jr_thermal_fields f;
    ThDims d;
    ThTable tab;
    double _di[3], _dt, dt, L, Vpdtau, dir_const;
    int form;
    int cf_lo[3], cf_hi[3];
    double cfv_lo[3], cfv_hi[3];
    int write_q2, write_res, next_pt;   // next_pt: k_th_update also writes θr_dτ, dτ_ρ of the NEXT iteration (update_pt_thermal_arrays! fused)
    double *res_part;  // per-block partial sums of ResT²
    int k_var;         // some phase has a T, P dependent conductivity (TP_Conductivity): K̄ is formed inside compute_flux! every iteration
    double *Kf[3];     // rheology form: face conductivities K̄ (static for the ConstantConductivity subset), computed once per call
};

// ---- GeoParams subset (restated from the published definitions; same operation order as oracle/thermal.c) ----
__device__ __forceinline__ double th_density(const jr_thermal_phase &p, double T, double P)
{
    if (p.rho_kind == 1) return p.rho0 * (1.0 - p.alpha * (T - p.T0) + p.beta * (P - p.P0));
    if (p.rho_kind == 2) return p.rho0 * (1.0 - p.alpha * (T - p.T0));
    return p.rho0;
}
// compute_conductivity: ConstantConductivity, or TP_Conductivity k = (a + b / (T + c)) · (1 + d · P)
__device__ __forceinline__ double th_cond(const jr_thermal_phase &p, double T, double P)
{
    return p.k_kind == 1 ? (p.k_a + p.k_b / (T + p.k_c)) * (1.0 + p.k_d * P) : p.k;
}
// fn_ratio(fn, rheology, ratio, args)  src/phases/phases.jl:18-30 (a ratio equal to one returns that phase alone)
__device__ __forceinline__ double th_rhoCp(const ThTable &t, const double *ph, size_t stride, size_t idx, double T, double P)
{
    if (!ph) return t.p[0].Cp * th_density(t.p[0], T, P);
    double x = 0.0;
    for (int q = 0; q < t.nphase; q++) {
        const double r = ph[(size_t)q * stride + idx];
        const double v = t.p[q].Cp * th_density(t.p[q], T, P);
        if (r == 1.0) return v * r;
        x += (r == 0.0) ? 0.0 : v * r;
    }
    return x;
}
__device__ __forceinline__ double th_K(const ThTable &t, const double *ph, size_t stride, size_t idx, double T, double P)
{
    if (!ph) return th_cond(t.p[0], T, P);
    double x = 0.0;
    for (int q = 0; q < t.nphase; q++) {
        const double r = ph[(size_t)q * stride + idx];
        if (r == 1.0) return th_cond(t.p[q], T, P) * r;
        x += (r == 0.0) ? 0.0 : th_cond(t.p[q], T, P) * r;
    }
    return x;
}
// fn_ratio(fn, rheology, ratio)  phases.jl:5-16
__device__ __forceinline__ double th_Hr(const ThTable &t, const double *ph, size_t stride, size_t idx)
{
    if (!ph) return t.p[0].has_Hr ? t.p[0].Hr : 0.0;
    double x = 0.0;
    for (int q = 0; q < t.nphase; q++) {
        const double r = ph[(size_t)q * stride + idx];
        x += (r == 0.0) ? 0.0 : t.p[q].Hr * r;
    }
    return x;
}
__device__ __forceinline__ double th_alpha(const ThTable &t, const double *ph, size_t stride, size_t idx)
{
    if (!ph) return t.p[0].alpha;
    double x = 0.0;
    for (int q = 0; q < t.nphase; q++) {
        const double r = ph[(size_t)q * stride + idx];
        x += (r == 0.0) ? 0.0 : t.p[q].alpha * r;
    }
    return x;
}


// the same phase-weighted evaluations from a register copy of the ratios (loaded once per cell); identical operation order
struct ThRatios { double r[TH_MAX_PHASES]; };
__device__ __forceinline__ void th_load_ratios(const ThTable &t, const double *__restrict__ ph, size_t stride, size_t idx, ThRatios &R)
{
#pragma unroll
    for (int q = 0; q < TH_MAX_PHASES; q++) R.r[q] = q < t.nphase ? ph[(size_t)q * stride + idx] : 0.0;
}
__device__ __forceinline__ double th_rhoCp_r(const ThTable &t, const ThRatios &R, double T, double P)
{
    double x = 0.0, out = 0.0;
    bool done = false;
#pragma unroll
    for (int q = 0; q < TH_MAX_PHASES; q++)
        if (q < t.nphase && !done) {
            const double v = t.p[q].Cp * th_density(t.p[q], T, P);
            if (R.r[q] == 1.0) { out = v * R.r[q]; done = true; }
            else x += (R.r[q] == 0.0) ? 0.0 : v * R.r[q];
        }
    return done ? out : x;
}
__device__ __forceinline__ double th_K_r(const ThTable &t, const ThRatios &R, double T, double P)
{
    double x = 0.0, out = 0.0;
    bool done = false;
#pragma unroll
    for (int q = 0; q < TH_MAX_PHASES; q++)
        if (q < t.nphase && !done) {
            if (R.r[q] == 1.0) { out = th_cond(t.p[q], T, P) * R.r[q]; done = true; }
            else x += (R.r[q] == 0.0) ? 0.0 : th_cond(t.p[q], T, P) * R.r[q];
        }
    return done ? out : x;
}
__device__ __forceinline__ double th_Hr_r(const ThTable &t, const ThRatios &R)
{
    double x = 0.0;
#pragma unroll
    for (int q = 0; q < TH_MAX_PHASES; q++)
        if (q < t.nphase) x += (R.r[q] == 0.0) ? 0.0 : t.p[q].Hr * R.r[q];
    return x;
}

#define TH_PI 3.141592653589793

// compute_pt_thermal_arrays!  DiffusionPT_coefficients.jl:105-151
__global__ void k_th_pt(const __grid_constant__ ThArgs a)
{
    const ThDims &d = a.d;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
    if (i >= d.nx || j >= d.ny) return;
    const size_t nc = (size_t)d.nx * d.ny * d.nz, c = th_ci(d, i, j, k);
    const double T = a.f.T[th_ti(d, i + 1, j + 1, d.nd == 3 ? k + 1 : 0)], P = a.f.P ? a.f.P[c] : 0.0;
    const double rhoCp = th_rhoCp(a.tab, a.f.phase_c, nc, c, T, P);
    const double _K = 1.0 / th_K(a.tab, a.f.phase_c, nc, c, T, P);
    const double _Re = 1.0 / (TH_PI + sqrt(TH_PI * TH_PI + rhoCp * (a.L * a.L) * _K * a._dt));
    a.f.theta_r_dtau[c] = a.L / a.Vpdtau * _Re;
    a.f.dtau_rho[c] = a.Vpdtau * a.L * _K * _Re;
}

// compute_flux!  (node I handles the x-, y-, z-face with index I; launch range ni .+ 1)
template <int DIM>
__device__ __forceinline__ void th_flux_dim(const ThArgs &a, int i, int j, int k)
{
    const ThDims &d = a.d;
    const int nc3[3] = {d.nx, d.ny, d.nz};
    int e[3] = {d.nx, d.ny, d.nz};
    e[DIM] += 1;
    const int I[3] = {i, j, k};
    if (i >= e[0] || j >= e[1] || k >= e[2]) return;
    double *q = DIM == 0 ? a.f.qTx : DIM == 1 ? a.f.qTy : a.f.qTz, *q2 = DIM == 0 ? a.f.qTx2 : DIM == 1 ? a.f.qTy2 : a.f.qTz2;
    const size_t qi = ((size_t)k * e[1] + j) * e[0] + i;
    if (I[DIM] == 0 && a.cf_lo[DIM]) { q[qi] = a.cfv_lo[DIM]; return; }
    if (I[DIM] == e[DIM] - 1 && a.cf_hi[DIM]) { q[qi] = a.cfv_hi[DIM]; return; }
    int L[3] = {i, j, k}, R[3] = {i, j, k};
    L[DIM] = jr_clamp(I[DIM] - 1, 0, nc3[DIM] - 1);
    R[DIM] = jr_clamp(I[DIM], 0, nc3[DIM] - 1);
    const size_t cL = th_ci(d, L[0], L[1], L[2]), cR = th_ci(d, R[0], R[1], R[2]);
    const int g3 = d.nd == 3;
    int tl[3] = {i + 1, j + 1, g3 ? k + 1 : 0}, th[3] = {i + 1, j + 1, g3 ? k + 1 : 0};
    tl[DIM] = I[DIM]; th[DIM] = I[DIM] + 1;
    const double Tl = a.f.T[th_ti(d, tl[0], tl[1], tl[2])], Th = a.f.T[th_ti(d, th[0], th[1], th[2])];
    double K;
    if (a.form == 0) K = (a.f.K[cL] + a.f.K[cR]) * 0.5;
    else {
        // face phase ratios indexed with the clamped CENTRE indices (quirk Q9)
        if (a.Kf[DIM]) K = a.Kf[DIM][qi];
        else {
            const double *phf = DIM == 0 ? a.f.phase_x : DIM == 1 ? a.f.phase_y : a.f.phase_z;
            const size_t ps = (size_t)e[0] * e[1] * e[2];
            const size_t pL = ((size_t)L[2] * e[1] + L[1]) * e[0] + L[0], pR = ((size_t)R[2] * e[1] + R[1]) * e[0] + R[0];
            // args of compute_conductivity: T = the mean of the two nodes adjacent to the face, P of the (clamped) cell on either side
            // (DiffusionPT_kernels.jl:93-100, 391-402); only TP_Conductivity looks at them
            const double Tf = (Tl + Th) * 0.5, PL = (a.k_var && a.f.P) ? a.f.P[cL] : 0.0, PR = (a.k_var && a.f.P) ? a.f.P[cR] : 0.0;
            K = (th_K(a.tab, phf, ps, pL, Tf, PL) + th_K(a.tab, phf, ps, pR, Tf, PR)) * 0.5;
        }
    }
    const double th_ = (a.f.theta_r_dtau[cL] + a.f.theta_r_dtau[cR]) * 0.5;
    const double qx = -K * (Th - Tl) * a._di[DIM];
    if (a.write_q2) q2[qi] = qx;
    q[qi] = (q[qi] * th_ + qx) / (1.0 + th_);
}

// K̄ at the faces of dimension DIM: (K(cL) + K(cR)) / 2 with the face phase ratios indexed by the clamped centre indices (quirk Q9) — the
// expression compute_flux! evaluates every iteration; for ConstantConductivity it does not change during a solve
template <int DIM>
__device__ __forceinline__ void th_kface_dim(const ThArgs &a, int i, int j, int k)
{
    const ThDims &d = a.d;
    const int nc3[3] = {d.nx, d.ny, d.nz};
    int e[3] = {d.nx, d.ny, d.nz};
    e[DIM] += 1;
    const int I[3] = {i, j, k};
    if (i >= e[0] || j >= e[1] || k >= e[2]) return;
    int L[3] = {i, j, k}, R[3] = {i, j, k};
    L[DIM] = jr_clamp(I[DIM] - 1, 0, nc3[DIM] - 1);
    R[DIM] = jr_clamp(I[DIM], 0, nc3[DIM] - 1);
    const size_t qi = ((size_t)k * e[1] + j) * e[0] + i;
    if (a.form == 0) {  // array form: K at the two (clamped) cells
        a.Kf[DIM][qi] = (a.f.K[th_ci(d, L[0], L[1], L[2])] + a.f.K[th_ci(d, R[0], R[1], R[2])]) * 0.5;
        return;
    }
    const double *phf = !a.f.phase_c ? nullptr : DIM == 0 ? a.f.phase_x : DIM == 1 ? a.f.phase_y : a.f.phase_z;
    const size_t ps = (size_t)e[0] * e[1] * e[2];
    const size_t pL = ((size_t)L[2] * e[1] + L[1]) * e[0] + L[0], pR = ((size_t)R[2] * e[1] + R[1]) * e[0] + R[0];
    a.Kf[DIM][qi] = (th_K(a.tab, phf, ps, pL, 0.0, 0.0) + th_K(a.tab, phf, ps, pR, 0.0, 0.0)) * 0.5;   // (constant conductivities only: th_prepare_kface)
}
__global__ void __launch_bounds__(256) k_th_kface(const __grid_constant__ ThArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
    th_kface_dim<0>(a, i, j, k);
    th_kface_dim<1>(a, i, j, k);
    if (a.d.nd == 3) th_kface_dim<2>(a, i, j, k);
}

__global__ void __launch_bounds__(256) k_th_flux(const __grid_constant__ ThArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
    th_flux_dim<0>(a, i, j, k);
    th_flux_dim<1>(a, i, j, k);
    if (a.d.nd == 3) th_flux_dim<2>(a, i, j, k);
}

__device__ __forceinline__ double th_div(const ThArgs &a, int i, int j, int k, bool second)
{
    const ThDims &d = a.d;
    const double *qx = second ? a.f.qTx2 : a.f.qTx, *qy = second ? a.f.qTy2 : a.f.qTy, *qz = second ? a.f.qTz2 : a.f.qTz;
    const size_t ix = ((size_t)k * d.ny + j) * (d.nx + 1) + i, iy = ((size_t)k * (d.ny + 1) + j) * d.nx + i;
    double s = (qx[ix + 1] - qx[ix]) * a._di[0] + (qy[iy + d.nx] - qy[iy]) * a._di[1];
    if (d.nd == 3) {
        const size_t iz = ((size_t)k * d.ny + j) * d.nx + i;
        s = s + (qz[iz + (size_t)d.nx * d.ny] - qz[iz]) * a._di[2];
    }
    return s;
}

// update_T! (+ check_res! on sampled iterations: the residual is evaluated with the UPDATED T and the raw fluxes qT2 of
// this iteration, exactly what check_res! sees when it runs after thermal_bcs! — it only reads interior T)
__global__ void __launch_bounds__(256) k_th_update(const __grid_constant__ ThArgs a)
{
    const ThDims &d = a.d;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
    double r2 = 0.0;
    if (i < d.nx && j < d.ny) {
        const size_t nc = (size_t)d.nx * d.ny * d.nz, c = th_ci(d, i, j, k), t = th_ti(d, i + 1, j + 1, d.nd == 3 ? k + 1 : 0);
        const bool dir = a.f.dir_mask && a.f.dir_mask[t] != 0.0;
        double Tn;
        const double Tc = a.f.T[t], Told = a.f.Told[t];
        double rhoCp = 0.0, src = 0.0, Hr = 0.0;
        ThRatios R;
        const bool hasR = a.form == 1 && a.f.phase_c;
        if (hasR) th_load_ratios(a.tab, a.f.phase_c, nc, c, R);
        if (dir) {
            // apply_mask!: A = inv(m)·A + m·B  (src/mask/mask.jl:51-52)
            const double m = a.f.dir_mask[t], B = a.f.dir_value ? a.f.dir_value[t] : a.dir_const;
            Tn = (1 - m) * Tc + m * B;
        } else if (a.form == 0) {
            rhoCp = a.f.rhoCp[c];
            const double dtr = a.f.dtau_rho[c];
            Tn = (dtr * (-(th_div(a, i, j, k, false)) + Told * rhoCp * a._dt + a.f.H[c] + a.f.shear_heating[c]) + Tc) / (1.0 + dtr * rhoCp * a._dt);
        } else {
            const double P = a.f.P ? a.f.P[c] : 0.0;
            if (hasR) { rhoCp = th_rhoCp_r(a.tab, R, Tc, P); Hr = th_Hr_r(a.tab, R); }
            else { rhoCp = th_rhoCp(a.tab, a.f.phase_c, nc, c, Tc, P); Hr = th_Hr(a.tab, a.f.phase_c, nc, c); }
            const double dtr = a.f.dtau_rho[c];
            Tn = (dtr * (-(th_div(a, i, j, k, false)) + Told * rhoCp * a._dt + Hr + a.f.H[c] + a.f.shear_heating[c] +
                         a.f.adiabatic[c] * Tc) + Tc) / (1.0 + dtr * rhoCp * a._dt);
        }
        a.f.T[t] = Tn;
        if (a.next_pt) {
            // update_pt_thermal_arrays! of the next iteration (solver.jl:233-234; DiffusionPT_coefficients.jl:105-151): it reads interior T only,
            // which neither thermal_bcs! nor update_halo!(T) touches, so evaluating it here from Tn is the same arithmetic on the same inputs
            const double rc = th_rhoCp_r(a.tab, R, Tn, a.f.P ? a.f.P[c] : 0.0);
            const double _K = 1.0 / th_K_r(a.tab, R, Tn, a.f.P ? a.f.P[c] : 0.0);
            const double _Re = 1.0 / (TH_PI + sqrt(TH_PI * TH_PI + rc * (a.L * a.L) * _K * a._dt));
            a.f.theta_r_dtau[c] = a.L / a.Vpdtau * _Re;
            a.f.dtau_rho[c] = a.Vpdtau * a.L * _K * _Re;
        }
        if (a.write_res) {
            double Rr = 0.0;
            if (!dir) {
                if (a.form == 0) Rr = -rhoCp * (Tn - Told) * a._dt - th_div(a, i, j, k, true) + a.f.H[c] + a.f.shear_heating[c];
                else {
                    const double rc2 = hasR ? th_rhoCp_r(a.tab, R, Tn, a.f.P ? a.f.P[c] : 0.0) : th_rhoCp(a.tab, a.f.phase_c, nc, c, Tn, a.f.P ? a.f.P[c] : 0.0);
                    Rr = -rc2 * (Tn - Told) * a._dt - th_div(a, i, j, k, true) + Hr + a.f.H[c] + a.f.shear_heating[c] + a.f.adiabatic[c] * Tn;
                }
            }
            a.f.ResT[c] = Rr;
            r2 = Rr * Rr;
        }
    }
    if (a.write_res) {
        __shared__ double sm[32];
        const double s = jr_block_sum(r2, sm);
        if (threadIdx.x == 0 && threadIdx.y == 0) a.res_part[(size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused compute_flux! + update_T! (+ update_pt_thermal_arrays! of the next iteration) for 3D, one launch per PT iteration.
// Jacobi-exact = the two-kernel sequence: every flux is formed from T_in, θ_in, q_in (ping-pong sets A ↔ B, so no thread reads what
// another one writes in this launch) with the expressions of th_flux_dim, every temperature from the new fluxes with the
// expressions of k_th_update — operation for operation, so the results are bit-identical to k_th_flux → k_th_update.
// A thread owns the column (i, j) of a z-chunk and marches in z: T, θ of the planes k−1, k, k+1 and the z-flux of the lower face
// roll through registers; the four lateral fluxes of a cell are formed by the cell itself (the x-high / y-high ones redundantly
// with the neighbour: 5 instead of 3 flux evaluations per cell, no shared memory, no barrier).  Index arithmetic is
// incremental (one pointer bump per array and plane).
// HBM traffic per cell (rheology form, N phases): reads T, K̄×3, θ, q×3, Told, dτ_ρ, H, H_s, A, P, ratios×N; writes q×3, T, θ, dτ_ρ
// = 19 + N passes against 24 + N of the two-kernel path (q is not re-read, T and θ are read once).
// phase-weighted evaluations from a register copy of the ratios with a COMPILE-TIME phase bound NP (same operation order as
// th_rhoCp_r / th_K_r / th_Hr_r: phases in table order, a ratio equal to one returns that phase alone)
template <int NP> struct ThRatiosN { double r[NP]; };
template <int NP>
__device__ __forceinline__ double th_rhoCp_n(const ThTable &t, const ThRatiosN<NP> &R, double T, double P)
{
    double x = 0.0, out = 0.0;
    bool done = false;
#pragma unroll
    for (int q = 0; q < NP; q++)
        if (q < t.nphase && !done) {
            const double v = t.p[q].Cp * th_density(t.p[q], T, P);
            if (R.r[q] == 1.0) { out = v * R.r[q]; done = true; }
            else x += (R.r[q] == 0.0) ? 0.0 : v * R.r[q];
        }
    return done ? out : x;
}
template <int NP>
__device__ __forceinline__ double th_K_n(const ThTable &t, const ThRatiosN<NP> &R)
{
    double x = 0.0, out = 0.0;
    bool done = false;
#pragma unroll
    for (int q = 0; q < NP; q++)
        if (q < t.nphase && !done) {
            if (R.r[q] == 1.0) { out = t.p[q].k * R.r[q]; done = true; }
            else x += (R.r[q] == 0.0) ? 0.0 : t.p[q].k * R.r[q];
        }
    return done ? out : x;
}
template <int NP>
__device__ __forceinline__ double th_Hr_n(const ThTable &t, const ThRatiosN<NP> &R)
{
    double x = 0.0;
#pragma unroll
    for (int q = 0; q < NP; q++)
        if (q < t.nphase) x += (R.r[q] == 0.0) ? 0.0 : t.p[q].Hr * R.r[q];
    return x;
}
struct ThPP {
    const double *T_in, *th_in, *q_in[3];
    double *T_out, *th_out, *q_out[3];
    int kchunk;
    int xfull;   // block columns that hold full 32-column tiles; block column xfull (if any) packs the nx mod 32 remainder columns
};
// the flux expression of th_flux_dim (compute_flux!, DiffusionPT_kernels.jl:6-158)
__device__ __forceinline__ double th_flux_pt(double q_old, double K, double Tl, double Th, double thL, double thR, double _d)
{
    const double th_ = (thL + thR) * 0.5;
    const double qx = -K * (Th - Tl) * _d;
    return jr_div_nr(q_old * th_ + qx, 1.0 + th_);  // IEEE-exact quotient (denominator ≥ 1), no slow-path call
}
__device__ __forceinline__ void th_pf(const double *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int FORM, int NP, bool PF, int MINB>  // NP = 0: no phase ratios; else compile-time phase bound; PF: L2 prefetch; MINB: CTAs/SM
__global__ void __launch_bounds__(256, MINB) k_th_fused3(const __grid_constant__ ThArgs a, const __grid_constant__ ThPP pp)
{
    const ThDims &d = a.d;
    const int nx = d.nx, ny = d.ny, nz = d.nz;
    // thread → column (i, j).  Full 32-column tiles: a warp = one x-row of the tile.  The nx mod 32 remainder columns (nx = 257: ONE
    // column, which would otherwise cost a ninth of the CTAs at 1/32 lane use) are packed densely into the CTAs of block column
    // `pp.xfull` — uncoalesced, but rem/nx of the cells.
    int i, j;
    if ((int)blockIdx.x < pp.xfull) {
        i = blockIdx.x * 32 + threadIdx.x; j = blockIdx.y * 8 + threadIdx.y;
    } else {
        const int rem = nx - pp.xfull * 32, lin = blockIdx.y * 256 + threadIdx.y * 32 + threadIdx.x;
        j = lin / rem; i = pp.xfull * 32 + (lin - j * rem);
    }
    if (i >= nx || j >= ny) return;
    const int k0 = blockIdx.z * pp.kchunk, k1 = min(k0 + pp.kchunk, nz);
    if (k0 >= k1) return;
    const size_t nc = (size_t)nx * ny * nz;
    const size_t sT = (size_t)d.gx * d.gy, sC = (size_t)nx * ny, sX = (size_t)(nx + 1) * ny, sY = (size_t)nx * (ny + 1);
    // clamped lateral neighbours of the cell (θ is indexed with clamped CENTRE indices)
    const int oxm = i > 0 ? -1 : 0, oxp = i < nx - 1 ? 1 : 0, oym = j > 0 ? -nx : 0, oyp = j < ny - 1 ? nx : 0;
    const double _dx = a._di[0], _dy = a._di[1], _dz = a._di[2], _dt = a._dt;
    // running indices of plane k
    size_t t = th_ti(d, i + 1, j + 1, k0 + 1), c = th_ci(d, i, j, k0);
    size_t fx = ((size_t)k0 * ny + j) * (nx + 1) + i, fy = ((size_t)k0 * (ny + 1) + j) * nx + i;  // x-low / y-low face of the cell
    const double *__restrict__ T = pp.T_in, *__restrict__ TH = pp.th_in;
    const double *__restrict__ qxi = pp.q_in[0], *__restrict__ qyi = pp.q_in[1], *__restrict__ qzi = pp.q_in[2];
    const double *__restrict__ Kx = a.Kf[0], *__restrict__ Ky = a.Kf[1], *__restrict__ Kz = a.Kf[2];
    // z carry: T(k−1), T(k); θ at the clamped cells k−1, k; the new flux through the lower z-face k
    double Tm = __ldg(T + t - sT), Tc = __ldg(T + t);
    double thc = __ldg(TH + c), thm = k0 > 0 ? __ldg(TH + c - sC) : thc;
    double qz_lo = th_flux_pt(__ldg(qzi + c), __ldg(Kz + c), Tm, Tc, thm, thc, _dz);  // z-face k has the index of cell k
    const double L_V = a.L / a.Vpdtau, VL = a.Vpdtau * a.L;
    constexpr bool HASR = NP > 0;
    constexpr int NR = NP > 0 ? NP : 1;
    for (int k = k0; k < k1; ++k) {
        // ---- L2 prefetch of the NEXT plane's operands (no registers held; the loads of the next step then hit L2) ----
        if (PF && k + 1 < k1) {
            th_pf(T + t + 2 * sT); th_pf(a.f.Told + t + sT);
            if (k + 2 < nz) th_pf(TH + c + 2 * sC);
            th_pf(qxi + fx + sX); th_pf(qyi + fy + sY); th_pf(qzi + c + 2 * sC);
            th_pf(Kx + fx + sX); th_pf(Ky + fy + sY); th_pf(Kz + c + 2 * sC);
            th_pf(a.f.dtau_rho + c + sC); th_pf(a.f.H + c + sC); th_pf(a.f.shear_heating + c + sC);
            if (FORM == 0) th_pf(a.f.rhoCp + c + sC);
            else {
                th_pf(a.f.adiabatic + c + sC);
                if (a.f.P) th_pf(a.f.P + c + sC);
                if (HASR) {
#pragma unroll
                    for (int q = 0; q < NR; q++)
                        if (q < a.tab.nphase) th_pf(a.f.phase_c + (size_t)q * nc + c + sC);
                }
            }
        }
        // ---- loads of this plane (independent: issued back to back) ----
        const double Tp = __ldg(T + t + sT);
        const double thp = k < nz - 1 ? __ldg(TH + c + sC) : thc;
        const double Txm = __ldg(T + t - 1), Txp = __ldg(T + t + 1), Tym = __ldg(T + t - d.gx), Typ = __ldg(T + t + d.gx);
        const double thxm = __ldg(TH + c + oxm), thxp = __ldg(TH + c + oxp), thym = __ldg(TH + c + oym), thyp = __ldg(TH + c + oyp);
        const double qxl_o = __ldg(qxi + fx), qxh_o = __ldg(qxi + fx + 1), qyl_o = __ldg(qyi + fy), qyh_o = __ldg(qyi + fy + nx);
        const double qzh_o = __ldg(qzi + c + sC);
        const double Kxl = __ldg(Kx + fx), Kxh = __ldg(Kx + fx + 1), Kyl = __ldg(Ky + fy), Kyh = __ldg(Ky + fy + nx), Kzh = __ldg(Kz + c + sC);
        const double Told = __ldg(a.f.Told + t), dtr = a.f.dtau_rho[c], Hc = __ldg(a.f.H + c), Hs = __ldg(a.f.shear_heating + c);
        double rhoCp_a = 0.0, adi = 0.0, P = 0.0;
        ThRatiosN<NR> R;
        if (FORM == 0) rhoCp_a = __ldg(a.f.rhoCp + c);
        else {
            adi = __ldg(a.f.adiabatic + c);
            P = a.f.P ? __ldg(a.f.P + c) : 0.0;
            if (HASR) {
#pragma unroll
                for (int q = 0; q < NR; q++) R.r[q] = q < a.tab.nphase ? __ldg(a.f.phase_c + (size_t)q * nc + c) : 0.0;
            }
        }
        // every load of this plane is in flight before the first use (the compiler otherwise sinks them into dependent batches)
        asm volatile("" ::: "memory");
        // ---- fluxes (x-low, x-high, y-low, y-high, z-high) ----
        const double qx_lo = th_flux_pt(qxl_o, Kxl, Txm, Tc, thxm, thc, _dx);
        const double qx_hi = th_flux_pt(qxh_o, Kxh, Tc, Txp, thc, thxp, _dx);
        const double qy_lo = th_flux_pt(qyl_o, Kyl, Tym, Tc, thym, thc, _dy);
        const double qy_hi = th_flux_pt(qyh_o, Kyh, Tc, Typ, thc, thyp, _dy);
        const double qz_hi = th_flux_pt(qzh_o, Kzh, Tc, Tp, thc, thp, _dz);
        pp.q_out[0][fx] = qx_lo;
        if (i == nx - 1) pp.q_out[0][fx + 1] = qx_hi;
        pp.q_out[1][fy] = qy_lo;
        if (j == ny - 1) pp.q_out[1][fy + nx] = qy_hi;
        pp.q_out[2][c] = qz_lo;
        if (k == nz - 1) pp.q_out[2][c + sC] = qz_hi;
        // ---- update_T!  (th_div: ((∂x qx) + (∂y qy)) + (∂z qz)) ----
        double div = (qx_hi - qx_lo) * _dx + (qy_hi - qy_lo) * _dy;
        div = div + (qz_hi - qz_lo) * _dz;
        double Tn;
        if (FORM == 0) {
            const double rhoCp = rhoCp_a;
            Tn = jr_div_nr(dtr * (-(div) + Told * rhoCp * _dt + Hc + Hs) + Tc, 1.0 + dtr * rhoCp * _dt);
        } else {
            double rhoCp, Hr;
            if (HASR) { rhoCp = th_rhoCp_n<NR>(a.tab, R, Tc, P); Hr = th_Hr_n<NR>(a.tab, R); }
            else { rhoCp = th_rhoCp(a.tab, nullptr, nc, c, Tc, P); Hr = th_Hr(a.tab, nullptr, nc, c); }
            Tn = jr_div_nr(dtr * (-(div) + Told * rhoCp * _dt + Hr + Hc + Hs + adi * Tc) + Tc, 1.0 + dtr * rhoCp * _dt);
            if (HASR) {
                // update_pt_thermal_arrays! of the next iteration from the new interior T (as k_th_update with next_pt)
                const double rc = th_rhoCp_n<NR>(a.tab, R, Tn, P);
                const double _K = jr_inv_nr(th_K_n<NR>(a.tab, R));
                const double _Re = jr_inv_nr(TH_PI + sqrt(TH_PI * TH_PI + rc * (a.L * a.L) * _K * _dt));
                pp.th_out[c] = L_V * _Re;
                a.f.dtau_rho[c] = VL * _K * _Re;
            }
        }
        pp.T_out[t] = Tn;
        // ---- roll ----
        Tm = Tc; Tc = Tp; thm = thc; thc = thp; qz_lo = qz_hi;
        t += sT; c += sC; fx += sX; fy += sY;
    }
}

__global__ void k_th_sum_parts(const double *__restrict__ part, size_t n, double *__restrict__ out)
{
    __shared__ double sm[32];
    double acc = 0.0;
    for (size_t q = threadIdx.x; q < n; q += blockDim.x) acc += part[q];
    const double s = jr_block_sum(acc, sm);
    if (threadIdx.x == 0) *out = s;
}

// thermal_bcs!: constant_value → no_flux → periodic (BoundaryConditions.jl:46-54) as one race-free gather: every ghost
// element is computed from the interior element it mirrors; the LAST active kind of a face wins (periodic over no-flux
// over constant value, quirk Q16).  Ghost edges/corners (never read on the path; thread-order dependent in the reference)
// chain the per-dimension rules in x, y, z order.
struct ThBc {
    double *T;
    ThDims d;
    int kind_lo[3], kind_hi[3];  // 0 none, 1 constant value, 2 no flux, 3 periodic
    double val_lo[3], val_hi[3];
};
__global__ void k_th_bc(const __grid_constant__ ThBc b)
{
    const ThDims &d = b.d;
    const int plane = blockIdx.z, dim = plane >> 1, hi = plane & 1;
    const int g[3] = {d.gx, d.gy, d.gz};
    const int u = dim == 0 ? 1 : 0, v = dim == 2 ? 1 : 2;
    const int p = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y * blockDim.y + threadIdx.y;
    if (p >= g[u] || q >= g[v]) return;
    if ((hi ? b.kind_hi[dim] : b.kind_lo[dim]) == 0) return;
    int c[3];
    c[dim] = hi ? g[dim] - 1 : 0; c[u] = p; c[v] = q;
    int s[3] = {c[0], c[1], c[2]};
    bool flip[3] = {false, false, false};
    double val[3] = {0, 0, 0};
    for (int e = 0; e < d.nd; e++) {
        const bool lo_e = c[e] == 0, hi_e = c[e] == g[e] - 1;
        if (!lo_e && !hi_e) continue;
        const int kind = lo_e ? b.kind_lo[e] : b.kind_hi[e];
        if (kind == 0) {
            if (e != dim) return;  // an edge whose other face has no condition: owned by that face's (absent) rule → leave
            continue;
        }
        if (kind == 3) s[e] = lo_e ? g[e] - 2 : 1;
        else s[e] = lo_e ? 1 : g[e] - 2;
        if (kind == 1) { flip[e] = true; val[e] = lo_e ? b.val_lo[e] : b.val_hi[e]; }
    }
    double x = b.T[th_ti(d, s[0], s[1], s[2])];
    for (int e = 0; e < d.nd; e++)
        if (flip[e]) x = 2 * val[e] - x;
    b.T[th_ti(d, c[0], c[1], c[2])] = x;
}

__global__ void k_th_sub(double *__restrict__ out, const double *__restrict__ a, const double *__restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = a[i] - b[i];
}

// adiabatic_heating  DiffusionPT_kernels.jl:720-731
__global__ void k_th_adiabatic(const __grid_constant__ ThArgs a, const double *__restrict__ P, const double *__restrict__ P0, size_t nc)
{
    for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < nc; c += (size_t)gridDim.x * blockDim.x)
        a.f.adiabatic[c] = (P[c] - P0[c]) * th_alpha(a.tab, a.f.phase_c, nc, c) * a._dt;
}

// ---------------------------------------------------------------------------------------------------------------
static void face_map(int nd, int lo[3], int hi[3])
{
    lo[0] = 0; hi[0] = 1;
    if (nd == 2) { lo[1] = 5; hi[1] = 4; lo[2] = hi[2] = -1; }
    else { lo[1] = 2; hi[1] = 3; lo[2] = 5; hi[2] = 4; }
}

static int th_check(const jr_thermal_fields *f, const jr_thermal_opts *o, ThArgs &a)
{
    JR_REQUIRE(f && o, JR_ERR_ARG, "null thermal fields/opts");
    JR_REQUIRE(f->ndim == 2 || f->ndim == 3, JR_ERR_SHAPE, "thermal solver: ndim = %d", f->ndim);
    for (int q = 0; q < f->ndim; q++) JR_REQUIRE(f->n[q] >= 2, JR_ERR_SHAPE, "thermal grid must have at least 2 cells per dimension");
    JR_REQUIRE(f->T && f->Told && f->dT && f->qTx && f->qTy && f->qTx2 && f->qTy2 && f->H && f->shear_heating && f->ResT && f->theta_r_dtau &&
                   f->dtau_rho && (f->ndim == 2 || (f->qTz && f->qTz2)),
               JR_ERR_SHAPE, "a required ThermalArrays / PTThermalCoeffs field is NULL");
    JR_REQUIRE(o->form == 0 || o->form == 1, JR_ERR_ARG, "thermal opts: form = %d", o->form);
    if (o->form == 0) JR_REQUIRE(f->K && f->rhoCp, JR_ERR_SHAPE, "array form needs K and ρCp");
    else {
        JR_REQUIRE(f->adiabatic, JR_ERR_SHAPE, "rheology form needs thermal.adiabatic");
        JR_REQUIRE(o->phases && o->nphase >= 1 && o->nphase <= TH_MAX_PHASES, JR_ERR_UNSUPPORTED,
                   "rheology form needs 1..%d phases in the thermal table (got %d)", TH_MAX_PHASES, o->nphase);
        for (int q = 0; q < o->nphase; q++)
            JR_REQUIRE(o->phases[q].rho_kind >= 0 && o->phases[q].rho_kind <= 2, JR_ERR_UNSUPPORTED,
                       "density law %d of phase %d is outside the supported GeoParams subset", o->phases[q].rho_kind, q);
        if (f->phase_c) JR_REQUIRE(f->phase_x && f->phase_y && (f->ndim == 2 || f->phase_z), JR_ERR_SHAPE, "phase ratios need centre and face arrays");
    }
    JR_REQUIRE(o->nout >= 1, JR_ERR_ARG, "nout must be >= 1");
    a.f = *f;
    a.d.nd = f->ndim; a.d.nx = f->n[0]; a.d.ny = f->n[1]; a.d.nz = f->ndim == 3 ? f->n[2] : 1;
    a.d.gx = a.d.nx + 2; a.d.gy = a.d.ny + 2; a.d.gz = f->ndim == 3 ? a.d.nz + 2 : 1;
    a.tab.nphase = o->form == 1 ? o->nphase : 0;
    a.k_var = 0;
    for (int q = 0; q < a.tab.nphase; q++) {
        a.tab.p[q] = o->phases[q];
        JR_REQUIRE(o->phases[q].k_kind == 0 || o->phases[q].k_kind == 1, JR_ERR_UNSUPPORTED, "phase %d: conductivity law %d outside the supported subset", q,
                   o->phases[q].k_kind);
        if (o->phases[q].k_kind == 1) a.k_var = 1;
    }
    for (int q = 0; q < 3; q++) a._di[q] = o->_di[q];
    a.dt = o->dt; a._dt = 1.0 / o->dt; a.L = o->max_lxyz; a.Vpdtau = o->Vpdtau; a.dir_const = o->dir_const; a.form = o->form;
    int lo[3], hi[3];
    face_map(f->ndim, lo, hi);
    for (int q = 0; q < 3; q++) {
        const bool on = q < f->ndim;
        a.cf_lo[q] = on ? o->cf_active[lo[q]] : 0; a.cf_hi[q] = on ? o->cf_active[hi[q]] : 0;
        a.cfv_lo[q] = on ? o->cf_value[lo[q]] : 0.0; a.cfv_hi[q] = on ? o->cf_value[hi[q]] : 0.0;
    }
    a.write_q2 = 0; a.write_res = 0; a.res_part = nullptr;
    a.Kf[0] = a.Kf[1] = a.Kf[2] = nullptr;
    return JR_OK;
}

static int th_launch_bc(jr_context *ctx, double *T, const ThDims &d, const jr_thermal_opts *o)
{
    ThBc b;
    b.T = T; b.d = d;
    int lo[3], hi[3];
    face_map(d.nd, lo, hi);
    bool any = false;
    for (int q = 0; q < 3; q++) {
        b.kind_lo[q] = b.kind_hi[q] = 0; b.val_lo[q] = b.val_hi[q] = 0.0;
        if (q >= d.nd) continue;
        const int fl = lo[q], fh = hi[q];
        if (o->cv_active[fl]) { b.kind_lo[q] = 1; b.val_lo[q] = o->cv_value[fl]; }
        if (o->cv_active[fh]) { b.kind_hi[q] = 1; b.val_hi[q] = o->cv_value[fh]; }
        if (o->no_flux[fl]) b.kind_lo[q] = 2;
        if (o->no_flux[fh]) b.kind_hi[q] = 2;
        if (o->periodic[fl]) b.kind_lo[q] = 3;
        if (o->periodic[fh]) b.kind_hi[q] = 3;
        any = any || b.kind_lo[q] || b.kind_hi[q];
    }
    if (!any) return JR_OK;
    int m = d.gx > d.gy ? d.gx : d.gy;
    m = m > d.gz ? m : d.gz;
    dim3 blk(32, 8, 1), grid((m + 31) / 32, d.nd == 3 ? (m + 7) / 8 : 1, 2 * d.nd);
    if (d.nd == 2) blk = dim3(128, 1, 1), grid = dim3((m + 127) / 128, 1, 4);
    k_th_bc<<<grid, blk, 0, ctx->stream>>>(b);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

static int th_halo_of(jr_context *ctx, const ThArgs &a, double *T)
{
    if (!ctx->comm) return JR_OK;
    const int32_t ext[3] = {a.d.gx, a.d.gy, a.d.gz}, nc[3] = {a.d.nx, a.d.ny, a.d.nz};
    const jr_harr H = jr_harr_dense(T, ext, nc);
    return jr_comm_halo(ctx, &H, 1);  // update_halo!(thermal.T)  DiffusionPT_solver.jl:110,261
}
static int th_halo(jr_context *ctx, const ThArgs &a) { return th_halo_of(ctx, a, a.f.T); }

// rheology form with phase ratios: face conductivities once per call (scratch owned by the context)
static int th_prepare_kface(jr_context *ctx, ThArgs &a, bool all_forms = false)
{
    if (!(a.form == 1 && a.f.phase_c) && !all_forms) return JR_OK;
    if (a.form == 1 && a.k_var) return JR_OK;   // TP_Conductivity: K̄ depends on the iterate (k_th_flux forms it)
    const ThDims &d = a.d;
    const size_t nfx = (size_t)(d.nx + 1) * d.ny * d.nz, nfy = (size_t)d.nx * (d.ny + 1) * d.nz, nfz = d.nd == 3 ? (size_t)d.nx * d.ny * (d.nz + 1) : 0;
    void *buf = nullptr;
    int st = jr_ctx_scratch(ctx, "th_kface", (nfx + nfy + nfz) * sizeof(double), &buf);
    if (st) return st;
    ThArgs b = a;
    b.Kf[0] = (double *)buf; b.Kf[1] = b.Kf[0] + nfx; b.Kf[2] = d.nd == 3 ? b.Kf[1] + nfy : nullptr;
    dim3 blk(32, 8, 1), g1((d.nx + 32) / 32, (d.ny + 8) / 8, d.nd == 3 ? d.nz + 1 : 1);
    k_th_kface<<<g1, blk, 0, ctx->stream>>>(b);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    a.Kf[0] = b.Kf[0]; a.Kf[1] = b.Kf[1]; a.Kf[2] = b.Kf[2];
    return JR_OK;
}

// one PT iteration; `sample` = this iteration's residual is read (iter % nout == 0)
// have_pt: θr_dτ, dτ_ρ of this iteration were already written by the previous k_th_update; write_next: this iteration's k_th_update writes the
// next iteration's (never on an iteration the loop can end on, so the arrays the caller sees are the reference's)
static int th_iter(jr_context *ctx, ThArgs &a, const jr_thermal_opts *o, bool sample, double *d_sum, bool have_pt, bool write_next)
{
    const ThDims &d = a.d;
    dim3 blk(32, 8, 1);
    dim3 g0((d.nx + 31) / 32, (d.ny + 7) / 8, d.nz), g1((d.nx + 32) / 32, (d.ny + 8) / 8, d.nd == 3 ? d.nz + 1 : 1);
    const bool pt_dyn = a.form == 1 && a.f.phase_c;
    a.next_pt = (pt_dyn && write_next) ? 1 : 0;
    if (pt_dyn && !have_pt) {
        k_th_pt<<<g0, blk, 0, ctx->stream>>>(a);  // update_pt_thermal_arrays!  solver.jl:233-234 (later iterations: fused into k_th_update)
        ctx->launches++;
    }
    a.write_q2 = sample; a.write_res = sample;
    if (sample) {
        void *part = nullptr;
        const size_t nb = (size_t)g0.x * g0.y * g0.z;
        int st = jr_ctx_scratch(ctx, "th_res_part", nb * sizeof(double), &part);
        if (st) return st;
        a.res_part = (double *)part;
    }
    k_th_flux<<<g1, blk, 0, ctx->stream>>>(a);
    k_th_update<<<g0, blk, 0, ctx->stream>>>(a);
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    int st = th_launch_bc(ctx, a.f.T, d, o);
    if (st) return st;
    if ((st = th_halo(ctx, a))) return st;
    if (sample) {
        k_th_sum_parts<<<1, 256, 0, ctx->stream>>>(a.res_part, (size_t)g0.x * g0.y * g0.z, d_sum);
        ctx->launches++;
        JR_CHECK_LAUNCH();
    }
    return JR_OK;
}

// ---- fused path (3D): set A = the caller's T, θr_dτ, qT arrays, set B = scratch owned by the context; iterations run in PAIRS
// (A → B, B → A) so the state is back in the caller's arrays whenever anything else (sampled iteration, exit) looks at it.
struct ThFused {
    bool ok = false, pt_dyn = false;
    double *Tb = nullptr, *thb = nullptr, *qb[3] = {nullptr, nullptr, nullptr};
    int kchunk = 12;
    bool prefetch = true;
    bool pack_rem = false;  // pack the nx mod 32 remainder columns (JRB200_TH_PACK=1).  Measured at 257³: 0.653 ms against 0.602 ms plain —
                            // the packed column's z-marching CTAs run uncoalesced for their whole chunk and become the tail; the same
                            // packing pays in the one-plane-per-CTA 3D-VC kernels
    int minb = 4;   // resident CTAs per SM the kernel is compiled for (2: 115 registers, 3: 78, 4: 64 — still no spills —, 5: 48 with spills)
};
static int th_fused_prepare(jr_context *ctx, ThArgs &a, ThFused &F)
{
    const ThDims &d = a.d;
    F.ok = d.nd == 3 && !a.f.dir_mask && !(a.form == 1 && a.k_var);   // (the fused kernel reads K̄ planes formed once per call)
    for (int q = 0; q < 3; q++) F.ok = F.ok && !a.cf_lo[q] && !a.cf_hi[q];
    if (const char *e = getenv("JRB200_TH_FUSED")) F.ok = F.ok && atoi(e) != 0;
    if (!F.ok) return th_prepare_kface(ctx, a);
    F.pt_dyn = a.form == 1 && a.f.phase_c;
    if (const char *e = getenv("JRB200_TH_KCHUNK")) F.kchunk = atoi(e) > 0 ? atoi(e) : F.kchunk;
    if (const char *e = getenv("JRB200_TH_PREFETCH")) F.prefetch = atoi(e) != 0;
    if (const char *e = getenv("JRB200_TH_PACK")) F.pack_rem = atoi(e) != 0;
    if (const char *e = getenv("JRB200_TH_MINB")) F.minb = (atoi(e) >= 2 && atoi(e) <= 5) ? atoi(e) : F.minb;
    const size_t nc = (size_t)d.nx * d.ny * d.nz, ng = (size_t)d.gx * d.gy * d.gz;
    const size_t nfx = (size_t)(d.nx + 1) * d.ny * d.nz, nfy = (size_t)d.nx * (d.ny + 1) * d.nz, nfz = (size_t)d.nx * d.ny * (d.nz + 1);
    void *p = nullptr;
    int st;
    if ((st = jr_ctx_scratch(ctx, "th_fused_T", ng * sizeof(double), &p))) return st;
    F.Tb = (double *)p;
    if ((st = jr_ctx_scratch(ctx, "th_fused_q", (nfx + nfy + nfz) * sizeof(double), &p))) return st;
    F.qb[0] = (double *)p; F.qb[1] = F.qb[0] + nfx; F.qb[2] = F.qb[1] + nfy;
    if (F.pt_dyn) {
        if ((st = jr_ctx_scratch(ctx, "th_fused_th", nc * sizeof(double), &p))) return st;
        F.thb = (double *)p;
    }
    // ghost elements no boundary condition touches keep their value in the reference: seed set B with them
    JR_CUDA(cudaMemcpyAsync(F.Tb, a.f.T, ng * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return th_prepare_kface(ctx, a, true);
}
static int th_fused_launch(jr_context *ctx, const ThArgs &a, const ThFused &F, bool a_to_b)
{
    const ThDims &d = a.d;
    ThPP pp;
    double *qA[3] = {a.f.qTx, a.f.qTy, a.f.qTz};
    pp.T_in = a_to_b ? a.f.T : F.Tb; pp.T_out = a_to_b ? F.Tb : a.f.T;
    for (int q = 0; q < 3; q++) { pp.q_in[q] = a_to_b ? qA[q] : F.qb[q]; pp.q_out[q] = a_to_b ? F.qb[q] : qA[q]; }
    if (F.pt_dyn) { pp.th_in = a_to_b ? a.f.theta_r_dtau : F.thb; pp.th_out = a_to_b ? F.thb : a.f.theta_r_dtau; }
    else { pp.th_in = a.f.theta_r_dtau; pp.th_out = a.f.theta_r_dtau; }
    pp.kchunk = F.kchunk;
    // remainder columns are packed when they are few (else the last block column is an ordinary, partly filled tile)
    const int rem = d.nx % 32;
    const bool pack = F.pack_rem && rem > 0 && rem <= 8 && d.nx >= 32;
    pp.xfull = pack ? d.nx / 32 : (d.nx + 31) / 32;
    dim3 blk(32, 8, 1), grid(pack ? pp.xfull + 1 : pp.xfull, (d.ny + 7) / 8, (d.nz + F.kchunk - 1) / F.kchunk);
    const bool pf = F.prefetch;
#define TH_LAUNCH_B(FORM_, NP_, B_)                                                                 \
    do {                                                                                            \
        if (pf) k_th_fused3<FORM_, NP_, true, B_><<<grid, blk, 0, ctx->stream>>>(a, pp);            \
        else k_th_fused3<FORM_, NP_, false, B_><<<grid, blk, 0, ctx->stream>>>(a, pp);              \
    } while (0)
#define TH_LAUNCH(FORM_, NP_)                                                                       \
    do {                                                                                            \
        if (F.minb == 2) TH_LAUNCH_B(FORM_, NP_, 2);                                                \
        else if (F.minb == 3) TH_LAUNCH_B(FORM_, NP_, 3);                                           \
        else if (F.minb == 5) TH_LAUNCH_B(FORM_, NP_, 5);                                           \
        else TH_LAUNCH_B(FORM_, NP_, 4);                                                            \
    } while (0)
    if (a.form == 0) TH_LAUNCH(0, 0);
    else if (!F.pt_dyn) TH_LAUNCH(1, 0);
    else if (a.tab.nphase <= 2) TH_LAUNCH(1, 2);
    else if (a.tab.nphase <= 3) TH_LAUNCH(1, 3);
    else if (a.tab.nphase <= 4) TH_LAUNCH(1, 4);
    else TH_LAUNCH(1, TH_MAX_PHASES);
#undef TH_LAUNCH
#undef TH_LAUNCH_B
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}
// two PT iterations whose residual nobody samples: A → B, thermal_bcs!(B), halo(B), B → A, thermal_bcs!(A), halo(A)
static int th_fused_pair(jr_context *ctx, const ThArgs &a, const jr_thermal_opts *o, const ThFused &F)
{
    int st;
    if ((st = th_fused_launch(ctx, a, F, true))) return st;
    if ((st = th_launch_bc(ctx, F.Tb, a.d, o))) return st;
    if ((st = th_halo_of(ctx, a, F.Tb))) return st;
    if ((st = th_fused_launch(ctx, a, F, false))) return st;
    if ((st = th_launch_bc(ctx, a.f.T, a.d, o))) return st;
    return th_halo_of(ctx, a, a.f.T);
}

extern "C" {

int jr_thermal_bcs(jr_context *ctx, double *T, int32_t ndim, const int32_t n[3], const jr_thermal_opts *o)
{
    JR_REQUIRE(ctx && T && n && o, JR_ERR_ARG, "jr_thermal_bcs: null argument");
    JR_REQUIRE(ndim == 2 || ndim == 3, JR_ERR_SHAPE, "jr_thermal_bcs: ndim = %d", ndim);
    JR_CUDA(cudaSetDevice(ctx->device));
    ThDims d;
    d.nd = ndim; d.nx = n[0]; d.ny = n[1]; d.nz = ndim == 3 ? n[2] : 1;
    d.gx = d.nx + 2; d.gy = d.ny + 2; d.gz = ndim == 3 ? d.nz + 2 : 1;
    int st = th_launch_bc(ctx, T, d, o);
    if (st) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_thermal_pt_arrays(jr_context *ctx, const jr_thermal_fields *f, const jr_thermal_opts *o)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    ThArgs a;
    int st = th_check(f, o, a);
    if (st) return st;
    JR_REQUIRE(o->form == 1, JR_ERR_ARG, "jr_thermal_pt_arrays needs the rheology form");
    JR_CUDA(cudaSetDevice(ctx->device));
    dim3 blk(32, 8, 1), g0((a.d.nx + 31) / 32, (a.d.ny + 7) / 8, a.d.nz);
    k_th_pt<<<g0, blk, 0, ctx->stream>>>(a);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_thermal_iterate(jr_context *ctx, const jr_thermal_fields *f, const jr_thermal_opts *o, int64_t niter, jr_thermal_result *res)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    ThArgs a;
    int st = th_check(f, o, a);
    if (st) return st;
    JR_CUDA(cudaSetDevice(ctx->device));
    void *slot = nullptr;
    if ((st = jr_ctx_scratch(ctx, "th_sum", 16 * sizeof(double), &slot))) return st;
    ctx->launches = 0;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    ThFused F;
    if ((st = th_fused_prepare(ctx, a, F))) return st;
    for (int64_t it = 0; it < niter;) {
        // pairs of iterations that are neither the first (PT coefficients still to be formed) nor the last (sampled)
        if (F.ok && (it > 0 || !F.pt_dyn) && it + 2 <= niter - 1) {
            if ((st = th_fused_pair(ctx, a, o, F))) return st;
            it += 2;
            continue;
        }
        if ((st = th_iter(ctx, a, o, it == niter - 1, (double *)slot, it > 0, it < niter - 1))) return st;
        it++;
    }
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res) {
        float ms = 0.f;
        JR_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const size_t nc = (size_t)a.d.nx * a.d.ny * a.d.nz;
        res->iter = niter; res->nhist = 0;
        res->err = niter > 0 ? sqrt(ctx->h_pinned[0]) * (1.0 / sqrt((double)nc)) : NAN;
        res->time_s = ms * 1e-3; res->kernel_launches = ctx->launches;
    }
    return JR_OK;
}

int jr_heatdiffusion_PT(jr_context *ctx, const jr_thermal_fields *f, const jr_thermal_opts *o, const double *stokes_P, const double *stokes_P0,
                        jr_thermal_result *res)
{
    JR_REQUIRE(ctx && res, JR_ERR_ARG, "null context/result");
    ThArgs a;
    int st = th_check(f, o, a);
    if (st) return st;
    JR_CUDA(cudaSetDevice(ctx->device));
    const ThDims &d = a.d;
    const size_t nc = (size_t)d.nx * d.ny * d.nz, ng = (size_t)d.gx * d.gy * d.gz;
    const double _sq_len_RT = 1.0 / sqrt((double)nc);
    void *slot = nullptr;
    if ((st = jr_ctx_scratch(ctx, "th_sum", 16 * sizeof(double), &slot))) return st;
    ctx->launches = 0;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    JR_CUDA(cudaMemcpyAsync(f->Told, f->T, ng * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));  // @copy thermal.Told thermal.T
    if (o->form == 1 && stokes_P && stokes_P0) {
        k_th_adiabatic<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(a, stokes_P, stokes_P0, nc);
        ctx->launches++;
    }
    ThFused F;
    if ((st = th_fused_prepare(ctx, a, F))) return st;
    int64_t iter = 0, cont = 0;
    double err = 2 * o->eps;
    bool have_pt = false;
    while (err > o->eps && iter < o->iterMax) {
        // two iterations in a row that are not sampled and not the last the loop can run (the condition cannot change in
        // between: err is only updated at samples): the fused kernel, set A → B → A
        if (F.ok && (have_pt || !F.pt_dyn) && (iter + 1) % o->nout != 0 && (iter + 2) % o->nout != 0 && iter + 2 < o->iterMax) {
            if ((st = th_fused_pair(ctx, a, o, F))) return st;
            iter += 2;
            continue;
        }
        const bool sample = (iter + 1) % o->nout == 0;
        const bool write_next = !(sample || iter + 1 >= o->iterMax);
        if ((st = th_iter(ctx, a, o, sample, (double *)slot, have_pt, write_next))) return st;
        have_pt = write_next;
        iter += 1;
        if (sample) {
            // err = norm(ResT)/√(prod(ni)): rank-local in the reference (quirk Q8); with a communicator the squared norm is
            // all-reduced so that every rank takes the same decision (documented deviation, identical on one rank)
            if (ctx->comm && (st = jr_comm_allreduce_dev(ctx, (double *)slot, 1, 0))) return st;
            JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            JR_CUDA(cudaStreamSynchronize(ctx->stream));
            double s = ctx->h_pinned[0];
            if (ctx->comm) s /= (double)ctx->comm->nranks;
            err = sqrt(s) * _sq_len_RT;
            if (cont < res->cap && res->norm_ResT && res->iter_count) { res->norm_ResT[cont] = err; res->iter_count[cont] = iter; }
            cont += 1;
            if (std::isnan(err)) break;  // the reference's loop also ends on NaN (err > ϵ is false)
        }
    }
    k_th_sub<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(f->dT, f->T, f->Told, ng);  // update_ΔT!
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    JR_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->iter = iter; res->nhist = cont; res->err = err; res->time_s = ms * 1e-3; res->kernel_launches = ctx->launches;
    return JR_OK;
}

} // extern "C"
