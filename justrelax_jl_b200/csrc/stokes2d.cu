// stokes2d.cu — the 2D Stokes PT loops of libjrb200 (sm_100a): ONE fused kernel per PT iteration.
//
//   2D-V2  visco-elastic, arrays G, K         src/stokes/Stokes2D.jl:181-325   (config 2, SolCx)
//   2D-VC  multiphase visco-elasto-plastic     src/stokes/Stokes2D.jl:577-866   (config 3, shear band)
//
// The reference launches per iteration: compute_maxloc! (VC) → compute_∇V! → compute_P! → [update_ρg!] → compute_strain_rate! →
// compute_τ! | update_stresses_center_vertex_ps! → [update_viscosity_τII!] → compute_V! → velocity2displacement! → flow_bcs!
// (7-10 launches, ≈ 60-120 array passes).  Here a CTA owns a (TX-2)×(TY-2) tile of cells/vertices plus a one-node rim that it
// recomputes; the three dependent stages (∇V,P,ε → τ,λ,P → V + boundary ghosts) communicate through shared memory.  State
// arrays are ping-ponged (set `in` → set `out`) which makes the iteration a race-free Jacobi step — exactly the semantics of
// the reference's separate launches, and for the racy VEP kernel (quirk Q7) the schedule the oracle declares canonical.
// Arithmetic is operation for operation the reference's (fma only where it writes fma/muladd; -fmad=false).
// Diagnostics nobody reads inside the loop (∇V, ε, ε_pl, RP, τII, η_vep, U, ρg) are only stored on the iterations whose
// result can be observed (every `nout`, and the last).
#include <algorithm>
#include "rheo.cuh"
#include "tma.cuh"
#include "comm.cuh"
#include "stokes2d_resident.cuh"

#define F(name) (s->f[JR_F_##name])
// 2D arrays stay far below 2^31 elements (check2d refuses larger grids): the index arithmetic of this file runs in 32 bits
#undef IX2
#define IX2(n1, i, j) (((j) - 1) * (n1) + ((i) - 1))
#define TX 32
#define TY 16            /* default tile height (threads in y) */
#define NT (TX * TY)

struct K2 {
    int nx, ny;
    double _dx, _dy, dt, r, th, edt, rel, nu, cut_lo, cut_hi;
    int fs_l, fs_r, fs_t, fs_b, ns_l, ns_r, ns_t, ns_b;
    int dT_ghosted;   // args.ΔT is (ni.+2), indexed ΔT[i, j] without offset (the reference's compute_P_kernel!)
    int dbc;   // flow_bcs isa DisplacementBoundaryConditions: flow_bcs! acts on U = V·dt, V keeps its ghosts / boundary faces
    // strain-increment form (INC): the displacement is part of the ping-pong state (Δε of the next iteration reads it at neighbours)
    const double *Ux_i, *Uy_i;
    double *Ux_o, *Uy_o, *dxx, *dyy, *dxy, *divU;
    // ping-pong state
    const double *Vx_i, *Vy_i, *P_i, *txx_i, *tyy_i, *txy_i, *th_i, *txyc_i, *lam_i, *lamv_i, *eta_i, *etav_i;
    double *Vx_o, *Vy_o, *P_o, *txx_o, *tyy_o, *txy_o, *th_o, *txyc_o, *lam_o, *lamv_o, *eta_o, *etav_o;
    // read-only inputs
    const double *P0, *Q, *K, *G, *etatau, *txxo, *tyyo, *txyo, *txyco, *rhogx, *rhogy, *T, *Pargs, *dTargs, *ph_c, *ph_v, *EII;
    // diagnostics (written when DIAG)
    double *divV, *RP, *exx, *eyy, *exy, *pxx, *pyy, *pxy, *tII, *eta_vep, *e_vol_pl, *Ux, *Uy, *rhogx_w, *rhogy_w, *etatau_w;
};

__device__ __forceinline__ double inv2(double xx, double yy, double xy) { return sqrt(0.5 * (xx * xx + yy * yy) + xy * xy); }
__device__ __forceinline__ void jr_prefetch_l1(const double *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// TYT = threads in y (tile height incl. the one-node rim): chosen per grid so that the CTA count fills whole waves (plan2_tile)
// INC (VC only): strain-increment form (kwarg strain_increment = true, Stokes2D.jl:659-730; StressKernels.jl:1147-1302): Δε from the
// displacement U, ε = Δε/dt, stress increments from Δε; the planes s_exx / s_eyy / s_exy then hold Δε and readers scale by 1/dt.
// RARE (VC only): the seldom-used options — args.ΔT (thermal-stress pressure form), DisplacementBoundaryConditions, cohesion softening,
// and INC — are compiled only into the RARE instantiations, so the common path keeps its register budget (no spills)
template <bool VC, bool DIAG, int TYT, bool INC = false, bool RARE = false>
__global__ void __launch_bounds__(TX * TYT, TYT <= 16 ? 2 : 1) k_stokes2d(const __grid_constant__ K2 a, const __grid_constant__ jr_phase_tab pt)
{
    constexpr int NTT = TX * TYT;
    extern __shared__ double sm[];
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * TX + tx;
    const int i = blockIdx.x * (TX - 2) + tx, j = blockIdx.y * (TYT - 2) + ty;  // 1-based node index; 0 = idle rim thread
    const int nx = a.nx, ny = a.ny;
    const bool vert = i >= 1 && j >= 1 && i <= nx + 1 && j <= ny + 1;
    const bool cell = vert && i <= nx && j <= ny;
    const bool own = tx >= 1 && tx <= TX - 2 && ty >= 1 && ty <= TYT - 2;
    const size_t nc = (size_t)nx * ny, nv = (size_t)(nx + 1) * (ny + 1);
    const size_t c = cell ? IX2(nx, i, j) : 0, v = vert ? IX2(nx + 1, i, j) : 0;

    double *s_th = sm, *s_exx = sm + NTT, *s_eyy = sm + 2 * NTT, *s_exy = sm + 3 * NTT, *s_ett = sm + 4 * NTT, *s_rgx = sm + 5 * NTT,
           *s_rgy = sm + 6 * NTT, *s_txxn = sm + 7 * NTT, *s_tyyn = sm + 8 * NTT, *s_txyn = sm + 9 * NTT, *s_Pn = sm + 10 * NTT,
           *s_txx = sm + 11 * NTT, *s_tyy = sm + 12 * NTT, *s_txxo = sm + 13 * NTT, *s_tyyo = sm + 14 * NTT, *s_eta = sm + 15 * NTT;

    // ---------------- stage 1: ητ, ∇V, P|θ, ε (centres) and εxy (vertices) ------------------------------------------------------
    // operands that are only read in stage 2 (after the first barrier): pull their lines into L1 now, so those loads do not expose
    // a DRAM / L2 round trip in the middle of the CTA (no registers held)
    if (vert) { jr_prefetch_l1(a.txy_i + v); jr_prefetch_l1(a.txyo + v); }
    if (VC) {
        if (vert) {
            jr_prefetch_l1(a.lamv_i + v);
            if (a.etav_o) jr_prefetch_l1(a.etav_i + v);
            for (int p = 0; p < pt.n; p++) jr_prefetch_l1(a.ph_v + (size_t)p * nv + v);
        }
        if (cell) { jr_prefetch_l1(a.txyc_i + c); jr_prefetch_l1(a.txyco + c); jr_prefetch_l1(a.lam_i + c); }
    }
    double divV = 0.0, RP = 0.0, thn = 0.0, exx = 0.0, eyy = 0.0, exy = 0.0, ett = 0.0, eta = 0.0, rgx = 0.0, rgy = 0.0;
    double txx = 0.0, tyy = 0.0, txxo = 0.0, tyyo = 0.0, Kc = 0.0, Gc = 0.0, divU = 0.0;
    const bool dbc = RARE && a.dbc;
    const double _dt = jr_inv(a.dt);
    if (cell) {
        eta = a.eta_i[c];
        if (VC) {  // compute_maxloc!(ητ, η; window = (1,1))  Stokes2D.jl:654, Utils.jl:409-461
            double x = -INFINITY;
            for (int jj = j - 1; jj <= j + 1; jj++)
                for (int ii = i - 1; ii <= i + 1; ii++) {
                    const double e = a.eta_i[IX2(nx, jr_clamp(ii, 1, nx), jr_clamp(jj, 1, ny))];
                    if (e > x) x = e;
                }
            ett = x;
        } else
            ett = a.etatau[c];
        const double dVx = (-a.Vx_i[IX2(nx + 1, i, j + 1)] + a.Vx_i[IX2(nx + 1, i + 1, j + 1)]) * a._dx;
        const double dVy = (-a.Vy_i[IX2(nx + 2, i + 1, j)] + a.Vy_i[IX2(nx + 2, i + 1, j + 1)]) * a._dy;
        divV = dVx + dVy;  // compute_∇V!  VelocityKernels.jl:3-6
        if (VC) {
            Kc = jr_ratio_Kb(pt, a.ph_c, nc, c);
            Gc = jr_ratio_G(pt, a.ph_c, nc, c);
            thn = a.th_i[c];
        } else {
            Kc = a.K[c];
            Gc = a.G[c];
            thn = a.P_i[c];
        }
        // compute_P! with ητ (quirk Q5)  Stokes2D.jl:231-233, 664-677; PressureKernels.jl:186-195
        if (VC && RARE && a.dTargs)  // args.ΔT given: thermal-stress form  PressureKernels.jl:128-149,197-206
            jr_compute_P_point_dT(RP, thn, a.P0[c], divV, a.Q[c], a.dTargs[a.dT_ghosted ? IX2(nx + 2, i, j) : c], jr_ratio_alpha(pt, a.ph_c, nc, c), ett, Kc, Gc, a.dt, a.r, a.th);
        else
            jr_compute_P_point(RP, thn, a.P0[c], divV, a.Q[c], ett, Kc, Gc, a.dt, a.r, a.th);
        if (INC) {  // compute_∇V!(∇U, U) + compute_strain_rate!(Δε, ∇U, U)  Stokes2D.jl:660-662, 681-689
            const double dUx = (-a.Ux_i[IX2(nx + 1, i, j + 1)] + a.Ux_i[IX2(nx + 1, i + 1, j + 1)]) * a._dx;
            const double dUy = (-a.Uy_i[IX2(nx + 2, i + 1, j)] + a.Uy_i[IX2(nx + 2, i + 1, j + 1)]) * a._dy;
            divU = dUx + dUy;
            const double dU = divU * jr_inv(3.0);
            exx = dUx - dU;
            eyy = dUy - dU;
        } else {
            const double dV = divV * jr_inv(3.0);
            exx = dVx - dV;  // compute_strain_rate!  VelocityKernels.jl:10-44
            eyy = dVy - dV;
        }
        txx = a.txx_i[c];
        tyy = a.tyy_i[c];
        txxo = a.txxo[c];
        tyyo = a.tyyo[c];
        if (VC && !pt.rho_const) {  // update_ρg!  Stokes2D.jl:679; BuoyancyForces.jl:74-95 (args.T sampled at I+1, quirk Q17)
            const double Tc = a.T ? a.T[IX2(nx + 2, i + 1, j + 1)] : 0.0, Pc = a.Pargs ? a.Pargs[c] : 0.0;
            const double rho = jr_ratio_density(pt, a.ph_c, nc, c, Tc, Pc);
            rgx = pt.g_scalar ? a.rhogx[c] : rho * pt.g[0];
            rgy = rho * pt.g[2];
        } else {
            rgx = a.rhogx[c];
            rgy = a.rhogy[c];
        }
    }
    if (vert) {
        const double *Ax = INC ? a.Ux_i : a.Vx_i, *Ay = INC ? a.Uy_i : a.Vy_i;
        exy = 0.5 * (a._dy * (Ax[IX2(nx + 1, i, j + 1)] - Ax[IX2(nx + 1, i, j)]) + a._dx * (Ay[IX2(nx + 2, i + 1, j)] - Ay[IX2(nx + 2, i, j)]));
    }
    s_th[t] = thn; s_exx[t] = exx; s_eyy[t] = eyy; s_exy[t] = exy; s_ett[t] = ett; s_rgx[t] = rgx; s_rgy[t] = rgy;
    if (VC) { s_txx[t] = txx; s_tyy[t] = tyy; s_txxo[t] = txxo; s_tyyo[t] = tyyo; s_eta[t] = eta; }
    __syncthreads();

    // ---------------- stage 2: stresses --------------------------------------------------------------------------------------------
    double txyn = 0.0, pxy = 0.0, lamv = 0.0, txxn = 0.0, tyyn = 0.0, txycn = 0.0, Pn = thn, lam = 0.0, pxx = 0.0, pyy = 0.0, tII = 0.0,
           etavep = 0.0, evol = 0.0;
    if (!VC) {  // compute_τ! 2D visco-elastic  StressKernels.jl:63-91
        if (cell) {
            const double _Gdt = jr_inv(Gc * a.dt), dtr = jr_dtau_r(a.th, eta, _Gdt);
            txxn = txx + jr_stress_increment(txx, txxo, eta, exx, _Gdt, dtr);
            tyyn = tyy + jr_stress_increment(tyy, tyyo, eta, eyy, _Gdt, dtr);
        }
        if (vert) {  // _av_ai_clamped  MiniKernels.jl:76-80
            const int i0 = jr_clamp(i - 1, 1, nx), i1 = jr_clamp(i, 1, nx), j0 = jr_clamp(j - 1, 1, ny), j1 = jr_clamp(j, 1, ny);
            const double e = 0.25 * (a.eta_i[IX2(nx, i0, j0)] + a.eta_i[IX2(nx, i1, j0)] + a.eta_i[IX2(nx, i0, j1)] + a.eta_i[IX2(nx, i1, j1)]);
            const double g = 0.25 * (a.G[IX2(nx, i0, j0)] + a.G[IX2(nx, i1, j0)] + a.G[IX2(nx, i0, j1)] + a.G[IX2(nx, i1, j1)]);
            const double _Gdt = jr_inv(g * a.dt), dtr = jr_dtau_r(a.th, e, _Gdt), t0 = a.txy_i[v];
            txyn = t0 + jr_stress_increment(t0, a.txyo[v], e, exy, _Gdt, dtr);
        }
    } else {  // update_stresses_center_vertex_ps! 2D  StressKernels.jl:992-1144, Jacobi schedule
        if (vert && tx >= 1 && ty >= 1) {
            const int i0 = jr_clamp(i - 1, 1, nx), ic = jr_clamp(i, 1, nx), j0 = jr_clamp(j - 1, 1, ny), jc = jr_clamp(j, 1, ny);
            const int q00 = (ty + j0 - j) * TX + (tx + i0 - i), qcc = (ty + jc - j) * TX + (tx + ic - i), q0c = (ty + jc - j) * TX + (tx + i0 - i),
                      qc0 = (ty + j0 - j) * TX + (tx + ic - i);
#define AVC(p) (0.25 * (p[q00] + p[qcc] + p[q0c] + p[qc0]))
            const double Pv = AVC(s_th), exxv = AVC(s_exx), eyyv = AVC(s_eyy), txxv = AVC(s_txx), tyyv = AVC(s_tyy);
            const double txxov = AVC(s_txxo), tyyov = AVC(s_tyyo);
#undef AVC
            bool is_pl;
            double eta_reg, Gv, Kv, dQdP, dFdP;   // phase mixture at the vertex: one sweep over the ratios (jr_mix_sweep)
            jr_mix_sweep(pt, a.ph_v, nv, v, Gv, Kv, is_pl, eta_reg, dQdP, dFdP);
            const double _Gdt = jr_inv(Gv * a.dt), _G = jr_inv(Gv);
            // harmonic mean of η (> 0, normal range) and 1/(θ_dτ + η/(G dt) + 1) (operand ≥ 1): branch-free IEEE-exact sequences
            const double etav = jr_div_nr(4.0, jr_inv_nr(s_eta[q00]) + jr_inv_nr(s_eta[qcc]) + jr_inv_nr(s_eta[q0c]) + jr_inv_nr(s_eta[qc0]));
            const double dtr = INC ? jr_inv(a.th * a.dt + etav * _G + a.dt) : jr_inv_nr(a.th + etav * _Gdt + 1.0);
            const double txyv = a.txy_i[v];
            double dxx, dyy, dxy;
            if (INC) {  // compute_stress_increment(τ, τ_o, η, Δε, 1/G, dτ_r, dt)  StressKernels.jl:19-22
                dxx = jr_stress_increment_d(txxv, txxov, etav, exxv, _G, dtr, a.dt);
                dyy = jr_stress_increment_d(tyyv, tyyov, etav, eyyv, _G, dtr, a.dt);
                dxy = jr_stress_increment_d(txyv, a.txyo[v], etav, exy, _G, dtr, a.dt);
            } else {
                dxx = jr_stress_increment(txxv, txxov, etav, exxv, _Gdt, dtr);
                dyy = jr_stress_increment(tyyv, tyyov, etav, eyyv, _Gdt, dtr);
                dxy = jr_stress_increment(txyv, a.txyo[v], etav, exy, _Gdt, dtr);
            }
            const double trial[3] = {txxv + dxx, tyyv + dyy, txyv + dxy};
            const double tIIv = inv2(dxx + txxv, dyy + tyyv, dxy + txyv);
            // (∂Q/∂τxy itself only where the vertex yields — jr_plastic_dQ: the divisions of compute_plastic_gradients_phase)
            const double volume = isinf(Kv) ? 0.0 : Kv * a.dt * dFdP * dQdP;
            double Fv;
            if (RARE && pt.any_soft) {  // cohesion softening: EII interpolated to the vertex (av_clamped, StressKernels.jl:1031)
                const double EIIv = 0.25 * (a.EII[IX2(nx, i0, j0)] + a.EII[IX2(nx, ic, jc)] + a.EII[IX2(nx, i0, jc)] + a.EII[IX2(nx, ic, j0)]);
                Fv = jr_yield_F_soft(pt, a.ph_v, nv, v, Pv, tIIv, EIIv);
            } else
                Fv = jr_yield_F(pt, a.ph_v, nv, v, Pv, tIIv);
            lamv = a.lamv_i[v];
            if (is_pl && tIIv != 0.0 && Fv > 0) {
                lamv = fma(a.rel, fmax(Fv, 0.0) / ((INC ? etav * dtr * a.dt : etav * dtr) + eta_reg + volume), (1.0 - a.rel) * lamv);
                pxy = lamv * jr_plastic_dQ<3, 2>(pt, a.ph_v, nv, v, trial, tIIv);   // tIIv = second invariant of `trial` (a + b = b + a exactly)
                txyn = txyv + (INC ? fma(-2.0, etav * a.dt * pxy * dtr, dxy) : fma(-2.0, etav * pxy * dtr, dxy));
            } else {
                txyn = txyv + dxy;
                pxy = 0.0;
            }
        }
        __syncthreads();  // (no data hazard: keeps the vertex and centre register live ranges apart)
        if (cell && tx <= TX - 2 && ty <= TYT - 2) {
            const double _Gdt = jr_inv(Gc * a.dt);
            bool is_pl;
            double eta_reg, dQdP, dFdP, G_unused, K_unused;
            jr_mix_sweep(pt, a.ph_c, nc, c, G_unused, K_unused, is_pl, eta_reg, dQdP, dFdP);   // (G, K of the centre are stage-1 values: dead code here)
            const double _G = jr_inv(Gc);
            const double dtr = INC ? 1.0 / (a.th * a.dt + eta * _G + a.dt) : 1.0 / (a.th + eta * _Gdt + 1.0);
            // strain rate at the centre (INC: ε = Δε·(1/dt) element by element, then the same four-vertex average)
            const double eij[3] = {INC ? exx * _dt : exx, INC ? eyy * _dt : eyy,
                                   INC ? (((s_exy[t] * _dt + s_exy[t + 1] * _dt) + s_exy[t + TX] * _dt) + s_exy[t + TX + 1] * _dt) / 4
                                       : (((s_exy[t] + s_exy[t + 1]) + s_exy[t + TX]) + s_exy[t + TX + 1]) / 4};
            double tij[3] = {txx, tyy, a.txyc_i[c]};
            const double tijo[3] = {txxo, tyyo, a.txyco[c]};
            double dt_[3];
            if (INC) {
                const double dij[3] = {exx, eyy, (((s_exy[t] + s_exy[t + 1]) + s_exy[t + TX]) + s_exy[t + TX + 1]) / 4};
#pragma unroll
                for (int q = 0; q < 3; q++) dt_[q] = jr_stress_increment_d(tij[q], tijo[q], eta, dij[q], _G, dtr, a.dt);
            } else {
#pragma unroll
                for (int q = 0; q < 3; q++) dt_[q] = jr_stress_increment(tij[q], tijo[q], eta, eij[q], _Gdt, dtr);
            }
            tII = inv2(dt_[0] + tij[0], dt_[1] + tij[1], dt_[2] + tij[2]);
            const double trial[3] = {tij[0] + dt_[0], tij[1] + dt_[1], tij[2] + dt_[2]};
            const double volume = isinf(Kc) ? 0.0 : Kc * a.dt * dFdP * dQdP;
            const double Fc = (RARE && pt.any_soft) ? jr_yield_F_soft(pt, a.ph_c, nc, c, thn, tII, a.EII[c]) : jr_yield_F(pt, a.ph_c, nc, c, thn, tII);
            lam = a.lam_i[c];
            if (is_pl && tII != 0.0 && Fc > 0) {
                lam = fma(a.rel, fmax(Fc, 0.0) / ((INC ? eta * dtr * a.dt : eta * dtr) + eta_reg + volume), (1.0 - a.rel) * lam);
                const double tIIt = tII;   // second invariant of `trial`
                const double dQ[3] = {jr_plastic_dQ<3, 0>(pt, a.ph_c, nc, c, trial, tIIt), jr_plastic_dQ<3, 1>(pt, a.ph_c, nc, c, trial, tIIt),
                                      jr_plastic_dQ<3, 2>(pt, a.ph_c, nc, c, trial, tIIt)};
                double epl[3];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    epl[q] = lam * dQ[q];
                    dt_[q] = INC ? fma(-2.0, eta * a.dt * epl[q] * dtr, dt_[q]) : fma(-2.0, eta * epl[q] * dtr, dt_[q]);
                    tij[q] = dt_[q] + tij[q];
                }
                evol = -lam * dQdP;
                txxn = tij[0]; tyyn = tij[1]; txycn = tij[2];
                pxx = epl[0]; pyy = epl[1];
                tII = inv2(tij[0], tij[1], tij[2]);
            } else {
                evol = 0.0;
                txxn = dt_[0] + tij[0]; tyyn = dt_[1] + tij[1]; txycn = dt_[2] + tij[2];
                pxx = 0.0; pyy = 0.0;
            }
            etavep = tII * 0.5 * jr_inv(inv2(eij[0], eij[1], eij[2]));
            Pn = thn - (isinf(Kc) ? 0.0 : Kc * a.dt * lam * dQdP);
        }
    }
    s_txxn[t] = txxn; s_tyyn[t] = tyyn; s_txyn[t] = txyn; s_Pn[t] = Pn;
    __syncthreads();

    // ---------------- stage 3: stores, viscosity relaxation, velocities + boundary ghosts ------------------------------------
    if (!own) return;
    if (cell) {
        a.P_o[c] = Pn;
        a.txx_o[c] = txxn;
        a.tyy_o[c] = tyyn;
        if (VC) {
            a.th_o[c] = thn;
            a.txyc_o[c] = txycn;
            a.lam_o[c] = lam;
            // update_viscosity_τII! AFTER the stress kernel (quirk Q13)  Stokes2D.jl:747-758; Viscosity.jl:382-418
            a.eta_o[c] = jr_clampd((1 - a.nu) * eta + a.nu * jr_phase_viscosity(pt, a.ph_c, nc, c), a.cut_lo, a.cut_hi);
        }
        if (DIAG) {
            a.divV[c] = divV; a.RP[c] = RP; a.exx[c] = INC ? exx * _dt : exx; a.eyy[c] = INC ? eyy * _dt : eyy;
            if (INC) { a.dxx[c] = exx; a.dyy[c] = eyy; a.divU[c] = divU; }
            if (VC) {
                a.pxx[c] = pxx; a.pyy[c] = pyy; a.tII[c] = tII; a.eta_vep[c] = etavep; a.e_vol_pl[c] = evol; a.etatau_w[c] = ett;
                if (!pt.rho_const) { if (!pt.g_scalar) a.rhogx_w[c] = rgx; a.rhogy_w[c] = rgy; }
            }
        }
    }
    if (vert) {
        a.txy_o[v] = txyn;
        if (VC) {
            a.lamv_o[v] = lamv;
            if (a.etav_o) a.etav_o[v] = jr_clampd((1 - a.nu) * a.etav_i[v] + a.nu * jr_phase_viscosity(pt, a.ph_v, nv, v), a.cut_lo, a.cut_hi);
        }
        if (DIAG) { a.exy[v] = INC ? exy * _dt : exy; if (INC) a.dxy[v] = exy; if (VC) a.pxy[v] = pxy; }
    }
    // compute_V! VelocityKernels.jl:108-131 (V2) / :134-180 (VC, free-surface form) + flow_bcs! (no_slip! → free_slip!) as a gather
    if (vert && j <= ny) {  // Vx[i, j+1]
        const size_t e = IX2(nx + 1, i, j + 1);
        double vx;
        if (i >= 2 && i <= nx) {
            const double dP = (-s_Pn[t - 1] + s_Pn[t]) * a._dx, dt_xx = (-s_txxn[t - 1] + s_txxn[t]) * a._dx;
            const double dt_xy = (-s_txyn[t] + s_txyn[t + TX]) * a._dy, avf = (s_rgx[t - 1] + s_rgx[t]) * 0.5, ave = (s_ett[t - 1] + s_ett[t]) * 0.5;
            vx = a.Vx_i[e] + (-dP + dt_xx + dt_xy - avf) * a.edt / ave;
        } else
            vx = (!dbc && ((i == 1) ? a.ns_l : a.ns_r)) ? 0.0 : a.Vx_i[e];
        a.Vx_o[e] = vx;
        // velocity2displacement! runs BEFORE flow_bcs!; with DisplacementBoundaryConditions flow_bcs! then acts on U (V untouched)
        double *const Uxw = INC ? a.Ux_o : ((DIAG && a.Ux) ? a.Ux : nullptr);
        const bool nzero = (i == 1 && a.ns_l) || (i == nx + 1 && a.ns_r);   // no_slip! zeroes the boundary-normal face (all rows)
        const double ux = (dbc && nzero) ? 0.0 : ((i >= 2 && i <= nx) ? vx : a.Vx_i[e]) * a.dt;
        if (Uxw) Uxw[e] = ux;
        if (j == 1) {
            const size_t g = IX2(nx + 1, i, 1);
            if (!dbc) {
                a.Vx_o[g] = a.fs_b ? vx : (a.ns_b ? -vx : (nzero ? 0.0 : a.Vx_i[g]));
                if (Uxw) Uxw[g] = a.Vx_i[g] * a.dt;
            } else {
                a.Vx_o[g] = a.Vx_i[g];
                if (Uxw) Uxw[g] = a.fs_b ? ux : (a.ns_b ? -ux : (nzero ? 0.0 : a.Vx_i[g] * a.dt));
            }
        }
        if (j == ny) {
            const size_t g = IX2(nx + 1, i, ny + 2);
            if (!dbc) {
                a.Vx_o[g] = a.fs_t ? vx : (a.ns_t ? -vx : (nzero ? 0.0 : a.Vx_i[g]));
                if (Uxw) Uxw[g] = a.Vx_i[g] * a.dt;
            } else {
                a.Vx_o[g] = a.Vx_i[g];
                if (Uxw) Uxw[g] = a.fs_t ? ux : (a.ns_t ? -ux : (nzero ? 0.0 : a.Vx_i[g] * a.dt));
            }
        }
    }
    if (vert && i <= nx) {  // Vy[i+1, j]
        const size_t e = IX2(nx + 2, i + 1, j);
        double vy;
        if (j >= 2 && j <= ny) {
            const double dP = (-s_Pn[t - TX] + s_Pn[t]) * a._dy, dt_yy = (-s_tyyn[t - TX] + s_tyyn[t]) * a._dy;
            const double dt_xy = (-s_txyn[t] + s_txyn[t + 1]) * a._dx, avf = (s_rgy[t - TX] + s_rgy[t]) * 0.5, ave = (s_ett[t - TX] + s_ett[t]) * 0.5;
            const double Vy0 = a.Vy_i[e];
            if (VC) {
                const double drg = (s_rgy[t] - s_rgy[t - TX]) * a._dy;
                const double corr = Vy0 * drg * 1.0 * pt.fs;
                vy = Vy0 + (-dP + dt_yy + dt_xy - avf + corr) * a.edt / ave;
            } else
                vy = Vy0 + (-dP + dt_yy + dt_xy - avf) * a.edt / ave;
        } else
            vy = (!dbc && ((j == 1) ? a.ns_b : a.ns_t)) ? 0.0 : a.Vy_i[e];
        a.Vy_o[e] = vy;
        double *const Uyw = INC ? a.Uy_o : ((DIAG && a.Uy) ? a.Uy : nullptr);
        const bool nzero = (j == 1 && a.ns_b) || (j == ny + 1 && a.ns_t);
        const double uy = (dbc && nzero) ? 0.0 : ((j >= 2 && j <= ny) ? vy : a.Vy_i[e]) * a.dt;
        if (Uyw) Uyw[e] = uy;
        if (i == 1) {
            const size_t g = IX2(nx + 2, 1, j);
            if (!dbc) {
                a.Vy_o[g] = a.fs_l ? vy : (a.ns_l ? -vy : (nzero ? 0.0 : a.Vy_i[g]));
                if (Uyw) Uyw[g] = a.Vy_i[g] * a.dt;
            } else {
                a.Vy_o[g] = a.Vy_i[g];
                if (Uyw) Uyw[g] = a.fs_l ? uy : (a.ns_l ? -uy : (nzero ? 0.0 : a.Vy_i[g] * a.dt));
            }
        }
        if (i == nx) {
            const size_t g = IX2(nx + 2, nx + 2, j);
            if (!dbc) {
                a.Vy_o[g] = a.fs_r ? vy : (a.ns_r ? -vy : (nzero ? 0.0 : a.Vy_i[g]));
                if (Uyw) Uyw[g] = a.Vy_i[g] * a.dt;
            } else {
                a.Vy_o[g] = a.Vy_i[g];
                if (Uyw) Uyw[g] = a.fs_r ? uy : (a.ns_r ? -uy : (nzero ? 0.0 : a.Vy_i[g] * a.dt));
            }
        }
    }
}

// compute_Res!  VelocityKernels.jl:246-307 (fs_form: the free-surface variant the VC solver launches)
__global__ void k_res2d(int nx, int ny, double _dx, double _dy, int fs_form, double fs, const double *__restrict__ P, const double *__restrict__ txx,
                        const double *__restrict__ tyy, const double *__restrict__ txy, const double *__restrict__ fx, const double *__restrict__ fy,
                        const double *__restrict__ Vy, double *__restrict__ Rx, double *__restrict__ Ry)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > nx || j > ny) return;
    if (i <= nx - 1)
        Rx[IX2(nx - 1, i, j)] = (-txx[IX2(nx, i, j)] + txx[IX2(nx, i + 1, j)]) * _dx + (-txy[IX2(nx + 1, i + 1, j)] + txy[IX2(nx + 1, i + 1, j + 1)]) * _dy -
                                (-P[IX2(nx, i, j)] + P[IX2(nx, i + 1, j)]) * _dx - (fx[IX2(nx, i, j)] + fx[IX2(nx, i + 1, j)]) * 0.5;
    if (j <= ny - 1) {
        double R = (-tyy[IX2(nx, i, j)] + tyy[IX2(nx, i, j + 1)]) * _dy + (-txy[IX2(nx + 1, i, j + 1)] + txy[IX2(nx + 1, i + 1, j + 1)]) * _dx -
                   (-P[IX2(nx, i, j)] + P[IX2(nx, i, j + 1)]) * _dy - (fy[IX2(nx, i, j)] + fy[IX2(nx, i, j + 1)]) * 0.5;
        if (fs_form) {
            const double V = Vy[IX2(nx + 2, i + 1, j + 1)];
            const double drg = (fy[IX2(nx, i, j + 1)] - fy[IX2(nx, i, j)]) * _dy;
            R = R + (V * drg) * 1.0 * fs;
        }
        Ry[IX2(nx, i, j)] = R;
    }
}

// compute_ρg!  BuoyancyForces.jl:74-95
__global__ void k_rhog2d(int nx, int ny, const __grid_constant__ jr_phase_tab pt, const double *__restrict__ ph_c, const double *__restrict__ T,
                         const double *__restrict__ Pa, double *__restrict__ rgx, double *__restrict__ rgy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > nx || j > ny) return;
    const size_t c = IX2(nx, i, j);
    const double Tc = T ? T[IX2(nx + 2, i + 1, j + 1)] : 0.0, Pc = Pa ? Pa[c] : 0.0;
    const double rho = jr_ratio_density(pt, ph_c, (size_t)nx * ny, c, Tc, Pc);
    if (!pt.g_scalar) rgx[c] = rho * pt.g[0];
    rgy[c] = rho * pt.g[2];
}

// compute_viscosity_kernel! (centres and — 2D only — vertices, quirk Q19)  Viscosity.jl:282-323, 382-418
__global__ void k_viscosity2d(size_t n, const __grid_constant__ jr_phase_tab pt, const double *__restrict__ ph, double *__restrict__ eta, double nu,
                              double lo, double hi)
{
    const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (q < n) eta[q] = jr_clampd((1 - nu) * eta[q] + nu * jr_phase_viscosity(pt, ph, n, q), lo, hi);
}

// second_invariant_staggered(xx, yy, gather(xy)) — tensor_invariant!  StressKernels.jl:470-480;
// mode 0: II = inv; mode 1: II += inv * f  (accumulate_tensor!  StressKernels.jl:379-408)
__global__ void k_inv_stag2d(int nx, int ny, double *__restrict__ II, const double *__restrict__ xx, const double *__restrict__ yy,
                             const double *__restrict__ xy, int mode, double f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > nx || j > ny) return;
    const size_t c = IX2(nx, i, j);
    const double A = xy[IX2(nx + 1, i, j)], B = xy[IX2(nx + 1, i + 1, j)], Cc = xy[IX2(nx + 1, i, j + 1)], D = xy[IX2(nx + 1, i + 1, j + 1)];
    const double X = xx[c], Y = yy[c];
    const double v = sqrt(0.5 * (X * X + Y * Y) + (((A * A + B * B) + Cc * Cc) + D * D) / 4);
    if (mode == 0) II[c] = v;
    else II[c] += v * f;
}
// shear2center!  Interpolations.jl:306-311
__global__ void k_shear2center2d(int nx, int ny, double *__restrict__ cen, const double *__restrict__ ver)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > nx || j > ny) return;
    cen[IX2(nx, i, j)] = 0.25 * (ver[IX2(nx + 1, i, j)] + ver[IX2(nx + 1, i + 1, j)] + ver[IX2(nx + 1, i, j + 1)] + ver[IX2(nx + 1, i + 1, j + 1)]);
}
// compute_vorticity! 2D  stress_rotation_particles.jl:17-30
__global__ void k_vorticity2d(int nx, int ny, double _dx, double _dy, const double *__restrict__ Vx, const double *__restrict__ Vy, double *__restrict__ w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > nx + 1 || j > ny + 1) return;
    w[IX2(nx + 1, i, j)] = 0.5 * ((-Vy[IX2(nx + 2, i, j)] + Vy[IX2(nx + 2, i + 1, j)]) * _dx - (-Vx[IX2(nx + 1, i, j)] + Vx[IX2(nx + 1, i, j + 1)]) * _dy);
}
// accumulate_vol!  StressKernels.jl:422-438:  A += dt * B
__global__ void k_axpy(size_t n, double *__restrict__ A, const double *__restrict__ B, double f)
{
    const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (q < n) A[q] += f * B[q];
}

// flow_bcs! 2D stand-alone, in place: no_slip! (sequential broadcasts) → free_slip! → periodic_boundary!
// no_slip.jl:1-19, free_slip.jl:1-13, periodic.jl:15-35
__global__ void k_no_slip2_face(double *Ax, double *Ay, int nx, int ny, int face)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int n1x = nx + 1, n2x = ny + 2, n1y = nx + 2, n2y = ny + 1;
    if (face == 0) { if (q <= n2x) Ax[IX2(n1x, 1, q)] = 0; if (q <= n2y) Ay[IX2(n1y, 1, q)] = -Ay[IX2(n1y, 2, q)]; }
    else if (face == 1) { if (q <= n2x) Ax[IX2(n1x, n1x, q)] = 0; if (q <= n2y) Ay[IX2(n1y, n1y, q)] = -Ay[IX2(n1y, n1y - 1, q)]; }
    else if (face == 5) { if (q <= n1x) Ax[IX2(n1x, q, 1)] = -Ax[IX2(n1x, q, 2)]; if (q <= n1y) Ay[IX2(n1y, q, 1)] = 0; }
    else { if (q <= n1x) Ax[IX2(n1x, q, n2x)] = -Ax[IX2(n1x, q, n2x - 1)]; if (q <= n1y) Ay[IX2(n1y, q, n2y)] = 0; }
}
__global__ void k_free_slip2(double *Ax, double *Ay, int nx, int ny, int l, int r, int t, int b)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int n1x = nx + 1, n2x = ny + 2, n1y = nx + 2, n2y = ny + 1;
    if (q <= n1x) {
        if (b) Ax[IX2(n1x, q, 1)] = Ax[IX2(n1x, q, 2)];
        if (t) Ax[IX2(n1x, q, n2x)] = Ax[IX2(n1x, q, n2x - 1)];
    }
    if (q <= n2y) {
        if (l) Ay[IX2(n1y, 1, q)] = Ay[IX2(n1y, 2, q)];
        if (r) Ay[IX2(n1y, n1y, q)] = Ay[IX2(n1y, n1y - 1, q)];
    }
}
// periodic_boundary! 2D: the reference kernel has read/write overlaps between indices, so its single-thread order (i ascending)
// is the defined result; O(n) sequential on one thread — a rarely used boundary path, not the hot loop
__global__ void k_periodic2_seq(double *Ax, double *Ay, int nx, int ny, int l, int r, int t, int b)
{
    if (blockIdx.x | threadIdx.x) return;
    const int n1x = nx + 1, n2x = ny + 2, n1y = nx + 2, n2y = ny + 1;
    int n = n1y > n2x ? n1y : n2x;
    const int m = n1x > n2y ? n1x : n2y;
    n = n > m ? n : m;
    for (int i = 1; i <= n; i++) {
        if (i <= n2x && l) Ax[IX2(n1x, 1, i)] = Ax[IX2(n1x, n1x, i)];
        if (i <= n2y) {
            if (l) Ay[IX2(n1y, 1, i)] = Ay[IX2(n1y, n1y - 1, i)];
            if (r) Ay[IX2(n1y, n1y, i)] = Ay[IX2(n1y, 2, i)];
        }
        if (i <= n1x) {
            if (b) Ax[IX2(n1x, i, 1)] = Ax[IX2(n1x, i, n2x - 1)];
            if (t) Ax[IX2(n1x, i, n2x)] = Ax[IX2(n1x, i, 2)];
        }
        if (i <= n1y && b) Ay[IX2(n1y, i, 1)] = Ay[IX2(n1y, i, n2y)];
    }
}

static int launch_flow_bcs2d(jr_context *ctx, double *Ax, double *Ay, int nx, int ny, const int32_t fs[6], const int32_t ns[6], const int32_t pe[6])
{
    const int m = (nx > ny ? nx : ny) + 2, blocks = (m + 127) / 128;
    const int order[4] = {0, 1, 5, 4};
    for (int q = 0; q < 4; q++)
        if (ns[order[q]]) { k_no_slip2_face<<<blocks, 128, 0, ctx->stream>>>(Ax, Ay, nx, ny, order[q]); ctx->launches++; }
    if (fs[0] | fs[1] | fs[4] | fs[5]) { k_free_slip2<<<blocks, 128, 0, ctx->stream>>>(Ax, Ay, nx, ny, fs[0], fs[1], fs[4], fs[5]); ctx->launches++; }
    if (pe[0] | pe[1] | pe[4] | pe[5]) { k_periodic2_seq<<<1, 32, 0, ctx->stream>>>(Ax, Ay, nx, ny, pe[0], pe[1], pe[4], pe[5]); ctx->launches++; }
    JR_CHECK_LAUNCH();
    return JR_OK;
}

int jr_make_phase_tab(const jr_vc_inputs *vc, jr_phase_tab *out)
{
    JR_REQUIRE(vc && vc->phases, JR_ERR_ARG, "null rheology table");
    JR_REQUIRE(vc->nphase >= 1 && vc->nphase <= JR_MAX_PHASES, JR_ERR_UNSUPPORTED, "number of phases %d outside 1..%d", vc->nphase, JR_MAX_PHASES);
    memset(out, 0, sizeof(*out));
    out->n = vc->nphase;
    out->g_scalar = vc->g_scalar;
    out->fs = vc->free_surface;
    out->rho_const = 1;
    for (int d = 0; d < 3; d++) out->g[d] = vc->g[d];
    for (int p = 0; p < vc->nphase; p++) {
        const jr_stokes_phase &q = vc->phases[p];
        JR_REQUIRE(q.rho_kind >= 0 && q.rho_kind <= 2, JR_ERR_UNSUPPORTED, "phase %d: density law %d outside the supported subset", p, q.rho_kind);
        out->eta[p] = q.eta; out->G[p] = q.G; out->Kb[p] = q.Kb; out->C[p] = q.C; out->sinphi[p] = q.sinphi; out->cosphi[p] = q.cosphi;
        out->sinpsi[p] = q.sinpsi; out->eta_vp[p] = q.eta_vp; out->rho0[p] = q.rho0; out->alpha[p] = q.alpha; out->beta[p] = q.beta;
        out->T0[p] = q.T0; out->P0[p] = q.P0; out->has_pl[p] = q.has_pl; out->rho_kind[p] = q.rho_kind;
        {   // compute_viscosity_τII(CompositeRheology) with dt = Inf  (Viscosity.jl:510-522, 599-619): IEEE divisions, as on the device
            volatile double ie = 1.0 / q.eta, ig = 1.0 / (q.G * INFINITY);
            volatile double ec = 1.0 / (ie + ig);
            out->eta_c[p] = ec;
            out->ieta_c[p] = 1.0 / ec;
        }
        if (q.rho_kind != 0) out->rho_const = 0;
        JR_REQUIRE(q.soft_C_kind >= 0 && q.soft_C_kind <= 2, JR_ERR_UNSUPPORTED, "phase %d: softening law %d outside the supported subset", p, q.soft_C_kind);
        out->soft_kind[p] = q.has_pl ? q.soft_C_kind : 0;
        for (int e = 0; e < 6; e++) out->soft[p][e] = q.soft_C[e];
        if (out->soft_kind[p]) out->any_soft = 1;
    }
    return JR_OK;
}

// ------------------------------------------------------------------------------------------------------------------------------
// host drivers
enum { S_Vx, S_Vy, S_P, S_txx, S_tyy, S_txy, S_th, S_txyc, S_lam, S_lamv, S_eta, S_etav, S_Ux, S_Uy, S_COUNT };
struct Plan2 {
    bool vc, inc;   // inc: strain-increment form (the displacement joins the ping-pong state)
    bool rare;      // any of the seldom-used options is on: the RARE kernel instantiations
    int nx, ny;
    size_t bytes[S_COUNT];
    double *set[2][S_COUNT];
    K2 k;
    jr_phase_tab pt;
    bool periodic;
    int32_t fs[6], ns[6], pe[6];
    int ty;   // tile height (threads in y) of k_stokes2d for this grid
    V2ResArgs rargs;   // 2D-V2: batches of non-observable iterations with the state resident in shared memory (stokes2d_resident.cu)
    V2ResPlan rplan;
};

// ---- tile height ----------------------------------------------------------------------------------------------------------------
// The kernel is latency-bound per CTA (load tile → three dependent stages with two barriers → store), so a partial last wave costs
// almost a full one: 511² with 16-row tiles is 666 CTAs on 296 slots = 2.25 → 3 waves, with 18-row tiles 576 CTAs = 2 waves.
// Candidates are instantiated at compile time; the choice minimises waves × (fixed latency + resident rows per SM).
template <bool VC, bool DIAG, int TYT>
static int k2_attr(int *nb)
{
    const int smem = (VC ? 16 : 11) * TX * TYT * 8;
    JR_CUDA(cudaFuncSetAttribute(k_stokes2d<VC, DIAG, TYT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (nb) JR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k_stokes2d<VC, DIAG, TYT>, TX * TYT, smem));
    return JR_OK;
}
template <bool VC, int TYT>
static int k2_attr_both(int *nb)
{
    int st = k2_attr<VC, true, TYT>(nullptr);
    if (!st && VC) {   // the strain-increment instantiations share the tile choice of the ε form
        constexpr int smem = 16 * TX * TYT * 8;
        JR_CUDA(cudaFuncSetAttribute(k_stokes2d<true, true, TYT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        JR_CUDA(cudaFuncSetAttribute(k_stokes2d<true, false, TYT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        JR_CUDA(cudaFuncSetAttribute(k_stokes2d<true, true, TYT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        JR_CUDA(cudaFuncSetAttribute(k_stokes2d<true, false, TYT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    return st ? st : k2_attr<VC, false, TYT>(nb);
}
#define K2_TILES(X) X(12) X(14) X(16) X(18) X(20)
static int plan2_tile(jr_context *ctx, Plan2 *p)
{
    static int nb_cache[2][32] = {};   // [vc][ty] resident CTAs per SM (0 = not queried yet, −1 = does not fit)
    int best_ty = TY;
    double best = 1e300;
    int forced = 0;
    if (const char *e = getenv("JRB200_2D_TY")) forced = atoi(e);
#define X(T_)                                                                                                              \
    {                                                                                                                      \
        int &nb = nb_cache[p->vc ? 1 : 0][T_];                                                                             \
        if (nb == 0) {                                                                                                     \
            int q = 0;                                                                                                     \
            const int st = p->vc ? k2_attr_both<true, T_>(&q) : k2_attr_both<false, T_>(&q);                               \
            if (st) return st;                                                                                             \
            nb = q > 0 ? q : -1;                                                                                           \
        }                                                                                                                  \
        if (nb > 0) {                                                                                                      \
            const long ctas = (long)((p->nx + 1 + TX - 3) / (TX - 2)) * ((p->ny + 1 + T_ - 3) / (T_ - 2));                 \
            const long slots = (long)nb * ctx->sm_count;                                                                   \
            const long waves = (ctas + slots - 1) / slots;                                                                 \
            const double est = (double)waves * (16.0 + (double)nb * T_);                                                   \
            if (forced == T_ || (!forced && est < best * 0.999)) { best = forced == T_ ? -1.0 : est; best_ty = T_; }       \
        }                                                                                                                  \
    }
    K2_TILES(X)
#undef X
    p->ty = best_ty;
    if (getenv("JRB200_VERBOSE")) fprintf(stderr, "[jrb200] k_stokes2d<vc=%d>: %d x %d grid, tile height %d\n", (int)p->vc, p->nx, p->ny, p->ty);
    return JR_OK;
}

// the 2D solvers are single-rank: the fused iteration has no halo exchange (the reference cuts it at update_halo!(ητ),
// update_halo!(τxy), update_halo!(V): Stokes2D.jl:655,757,784) and the norms are not all-reduced — refuse loudly instead
// of returning rank-local answers
static int single_rank2d(const jr_context *ctx)
{
    JR_REQUIRE(!(ctx->comm && ctx->comm->active), JR_ERR_UNSUPPORTED,
               "the 2D Stokes solvers run on one non-periodic rank only (a communicator with %d ranks%s is attached to this context)", ctx->comm->nranks,
               ctx->comm->nranks > 1 ? "" : " and a periodic dimension");
    return JR_OK;
}

static int check2d(const jr_fields *s, const jr_stokes_opts *o, bool vc, const jr_vc_inputs *in)
{
    JR_REQUIRE(s && o, JR_ERR_ARG, "null fields/opts");
    JR_REQUIRE(s->ndim == 2, JR_ERR_SHAPE, "2D solver called with ndim=%d", s->ndim);
    JR_REQUIRE(s->n[0] >= 3 && s->n[1] >= 3, JR_ERR_SHAPE, "grid must be at least 3 cells per dimension");
    JR_REQUIRE(((long long)s->n[0] + 3) * ((long long)s->n[1] + 3) < (1ll << 30), JR_ERR_SHAPE, "2D grid too large for the 32-bit index arithmetic of the 2D kernels");
    JR_REQUIRE(o->nout >= 1, JR_ERR_ARG, "nout must be >= 1");
    static const int req_common[] = {JR_F_P, JR_F_P0, JR_F_divV, JR_F_Q, JR_F_Vx, JR_F_Vy, JR_F_txx, JR_F_tyy, JR_F_txy, JR_F_txx_o, JR_F_tyy_o, JR_F_txy_o,
                                     JR_F_exx, JR_F_eyy, JR_F_exy, JR_F_eta, JR_F_etatau, JR_F_Rx, JR_F_Ry, JR_F_RP, JR_F_rhogx, JR_F_rhogy};
    for (int q : req_common) JR_REQUIRE(s->f[q] != nullptr, JR_ERR_SHAPE, "required field '%s' is NULL", jr_field_name(q));
    if (!vc) {
        JR_REQUIRE(F(K) && F(G), JR_ERR_SHAPE, "2D-V2 needs the K and G arrays");
        JR_REQUIRE(!o->strain_increment && !o->displacement_bcs, JR_ERR_UNSUPPORTED,
                   "strain_increment / DisplacementBoundaryConditions are supported by the multiphase 2D solve (2D-VC) only");
    } else {
        static const int req_vc[] = {JR_F_txy_c, JR_F_txy_o_c, JR_F_pxx, JR_F_pyy, JR_F_pxy, JR_F_tII, JR_F_eta_vep, JR_F_e_vol_pl, JR_F_EII_pl, JR_F_EVol_pl};
        for (int q : req_vc) JR_REQUIRE(s->f[q] != nullptr, JR_ERR_SHAPE, "required field '%s' is NULL", jr_field_name(q));
        JR_REQUIRE(in && in->ph_center && in->ph_vertex, JR_ERR_SHAPE, "2D-VC needs phase ratios at centres and vertices");
        if (o->strain_increment || o->displacement_bcs) {
            JR_REQUIRE(F(Ux) && F(Uy), JR_ERR_SHAPE, "strain_increment / DisplacementBoundaryConditions need the displacement arrays U");
        }
        if (o->strain_increment) {
            JR_REQUIRE(F(dxx) && F(dyy) && F(dxy) && F(divU), JR_ERR_SHAPE, "strain_increment needs the Δε tensor and ∇U");
        }
        if (o->displacement_bcs) JR_REQUIRE(!(o->periodic[0] | o->periodic[1] | o->periodic[4] | o->periodic[5]), JR_ERR_UNSUPPORTED,
                       "periodic DisplacementBoundaryConditions are outside the supported subset");
    }
    return JR_OK;
}

static int plan2_begin(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, bool vc, const jr_vc_inputs *in, Plan2 *p)
{
    memset(p, 0, sizeof(*p));
    p->vc = vc;
    const int nx = p->nx = s->n[0], ny = p->ny = s->n[1];
    const size_t nc = (size_t)nx * ny * 8, nv = (size_t)(nx + 1) * (ny + 1) * 8;
    p->inc = vc && o->strain_increment;
    const size_t nVx = (size_t)(nx + 1) * (ny + 2) * 8, nVy = (size_t)(nx + 2) * (ny + 1) * 8;
    const size_t b[S_COUNT] = {nVx, nVy, nc, nc, nc, nv, nc, nc, nc, nv, nc, nv, p->inc ? nVx : 0, p->inc ? nVy : 0};
    size_t off[S_COUNT + 1], tot = 0;
    for (int q = 0; q < S_COUNT; q++) { p->bytes[q] = b[q]; off[q] = tot; tot += (b[q] + 255) & ~(size_t)255; }
    off[S_COUNT] = tot;
    void *base = nullptr;
    int st = jr_ctx_scratch(ctx, "stokes2d_sets", 2 * tot, &base);
    if (st) return st;
    for (int q = 0; q < S_COUNT; q++) {
        p->set[0][q] = (double *)((char *)base + off[q]);
        p->set[1][q] = (double *)((char *)base + tot + off[q]);
    }
    // set 0 aliases the caller's arrays where they exist (no packing); θ, λ, λv are solver-local (Stokes2D.jl:635-637)
    p->set[0][S_Vx] = F(Vx); p->set[0][S_Vy] = F(Vy); p->set[0][S_P] = F(P); p->set[0][S_txx] = F(txx); p->set[0][S_tyy] = F(tyy); p->set[0][S_txy] = F(txy);
    p->set[0][S_eta] = F(eta);
    if (p->inc) { p->set[0][S_Ux] = F(Ux); p->set[0][S_Uy] = F(Uy); }
    if (vc) {
        p->set[0][S_txyc] = F(txy_c);
        if (F(etav)) p->set[0][S_etav] = F(etav);
    } else
        p->set[1][S_eta] = F(eta);  // η is constant in V2
    for (int q = 0; q < 6; q++) { p->fs[q] = o->free_slip[q]; p->ns[q] = o->no_slip[q]; p->pe[q] = o->periodic[q]; }
    p->periodic = o->periodic[0] | o->periodic[1] | o->periodic[4] | o->periodic[5];

    K2 &k = p->k;
    k.nx = nx; k.ny = ny;
    k._dx = o->_di[0]; k._dy = o->_di[1]; k.dt = o->dt; k.r = o->r; k.th = o->theta_dtau; k.edt = o->eta_dtau;
    k.rel = o->lambda_relaxation; k.nu = o->viscosity_relaxation; k.cut_lo = o->visc_cutoff_lo; k.cut_hi = o->visc_cutoff_hi;
    k.fs_l = o->free_slip[0]; k.fs_r = o->free_slip[1]; k.fs_t = o->free_slip[4]; k.fs_b = o->free_slip[5];
    k.ns_l = o->no_slip[0]; k.ns_r = o->no_slip[1]; k.ns_t = o->no_slip[4]; k.ns_b = o->no_slip[5];
    k.dbc = vc && o->displacement_bcs;
    k.dT_ghosted = o->dT_ghosted;
    k.dxx = F(dxx); k.dyy = F(dyy); k.dxy = F(dxy); k.divU = F(divU);
    k.P0 = F(P0); k.Q = F(Q); k.K = F(K); k.G = F(G); k.etatau = F(etatau);
    k.txxo = F(txx_o); k.tyyo = F(tyy_o); k.txyo = F(txy_o); k.txyco = F(txy_o_c);
    k.rhogx = F(rhogx); k.rhogy = F(rhogy); k.T = F(T); k.Pargs = F(Pargs); k.dTargs = F(dTargs); k.EII = F(EII_pl);
    k.divV = F(divV); k.RP = F(RP); k.exx = F(exx); k.eyy = F(eyy); k.exy = F(exy); k.pxx = F(pxx); k.pyy = F(pyy); k.pxy = F(pxy);
    k.tII = F(tII); k.eta_vep = F(eta_vep); k.e_vol_pl = F(e_vol_pl); k.Ux = F(Ux); k.Uy = F(Uy); k.rhogx_w = F(rhogx); k.rhogy_w = F(rhogy);
    k.etatau_w = F(etatau);
    if (vc) {
        if ((st = jr_make_phase_tab(in, &p->pt))) return st;
        k.ph_c = in->ph_center; k.ph_v = in->ph_vertex;
    }
    p->rare = vc && (p->inc || k.dbc || k.dTargs != nullptr || p->pt.any_soft);
    if ((st = plan2_tile(ctx, p))) return st;
    p->rplan = V2ResPlan();
    if (!vc && !p->periodic && !(ctx->flags & JR_FLAG_DIAG_EVERY_ITER)) {
        V2ResArgs &r = p->rargs;
        r.nx = nx; r.ny = ny; r._dx = k._dx; r._dy = k._dy; r.dt = k.dt; r.r = k.r; r.th = k.th; r.edt = k.edt;
        r.fs_l = k.fs_l; r.fs_r = k.fs_r; r.fs_t = k.fs_t; r.fs_b = k.fs_b; r.ns_l = k.ns_l; r.ns_r = k.ns_r; r.ns_t = k.ns_t; r.ns_b = k.ns_b;
        for (int q = 0; q < 2; q++) {
            r.Vx[q] = p->set[q][S_Vx]; r.Vy[q] = p->set[q][S_Vy]; r.P[q] = p->set[q][S_P]; r.txx[q] = p->set[q][S_txx]; r.tyy[q] = p->set[q][S_tyy];
            r.txy[q] = p->set[q][S_txy];
        }
        r.eta = F(eta); r.etatau = F(etatau); r.rhogx = F(rhogx); r.rhogy = F(rhogy); r.G = F(G); r.K = F(K); r.Q = F(Q);
        if ((st = jr_v2_resident_plan(ctx, &r, &p->rplan))) return st;
    }
    return JR_OK;
}

// iterations it … it + n − 1, none of them observable: resident batch when the plan has one, else the fused kernel n times
static int plan2_iter(jr_context *ctx, Plan2 *p, int64_t it, bool diag);
static int plan2_batch(jr_context *ctx, Plan2 *p, int64_t it, int64_t n)
{
    if (n <= 0) return JR_OK;
    if (p->rplan.ok && n >= 2) {
        int st;
        for (int64_t done = 0; done < n;) {
            const int64_t m = n - done < (1 << 30) ? n - done : (1 << 30);
            if ((st = jr_v2_resident_run(ctx, &p->rargs, &p->rplan, it + done, (int)m))) return st;
            done += m;
        }
        return JR_OK;
    }
    for (int64_t q = 0; q < n; q++) {
        int st = plan2_iter(ctx, p, it + q, false);
        if (st) return st;
    }
    return JR_OK;
}

// one PT iteration: set (it & 1) → set ((it + 1) & 1)
static int plan2_iter(jr_context *ctx, Plan2 *p, int64_t it, bool diag)
{
    K2 k = p->k;
    double *const *I = p->set[it & 1], *const *O = p->set[(it + 1) & 1];
    k.Vx_i = I[S_Vx]; k.Vy_i = I[S_Vy]; k.P_i = I[S_P]; k.txx_i = I[S_txx]; k.tyy_i = I[S_tyy]; k.txy_i = I[S_txy]; k.th_i = I[S_th];
    k.txyc_i = I[S_txyc]; k.lam_i = I[S_lam]; k.lamv_i = I[S_lamv]; k.eta_i = I[S_eta]; k.etav_i = I[S_etav];
    k.Vx_o = O[S_Vx]; k.Vy_o = O[S_Vy]; k.P_o = O[S_P]; k.txx_o = O[S_txx]; k.tyy_o = O[S_tyy]; k.txy_o = O[S_txy]; k.th_o = O[S_th];
    k.txyc_o = O[S_txyc]; k.lam_o = O[S_lam]; k.lamv_o = O[S_lamv]; k.eta_o = O[S_eta]; k.etav_o = O[S_etav];
    k.Ux_i = I[S_Ux]; k.Uy_i = I[S_Uy]; k.Ux_o = O[S_Ux]; k.Uy_o = O[S_Uy];
    if (p->vc && k.Pargs == p->set[0][S_P]) k.Pargs = I[S_P];  // args.P aliases stokes.P in the reference's scripts
    const int ty = p->ty;
    dim3 blk(TX, ty), grd((p->nx + 1 + TX - 3) / (TX - 2), (p->ny + 1 + ty - 3) / (ty - 2));
    const size_t smem = (size_t)(p->vc ? 16 : 11) * TX * ty * 8;
#define X(T_)                                                                                              \
    if (ty == T_) {                                                                                        \
        if (p->inc) {                                                                                      \
            if (diag) k_stokes2d<true, true, T_, true, true><<<grd, blk, smem, ctx->stream>>>(k, p->pt);   \
            else k_stokes2d<true, false, T_, true, true><<<grd, blk, smem, ctx->stream>>>(k, p->pt);       \
        } else if (p->vc && p->rare) {                                                                     \
            if (diag) k_stokes2d<true, true, T_, false, true><<<grd, blk, smem, ctx->stream>>>(k, p->pt);  \
            else k_stokes2d<true, false, T_, false, true><<<grd, blk, smem, ctx->stream>>>(k, p->pt);      \
        } else if (p->vc) {                                                                                \
            if (diag) k_stokes2d<true, true, T_><<<grd, blk, smem, ctx->stream>>>(k, p->pt);               \
            else k_stokes2d<true, false, T_><<<grd, blk, smem, ctx->stream>>>(k, p->pt);                   \
        } else {                                                                                           \
            if (diag) k_stokes2d<false, true, T_><<<grd, blk, smem, ctx->stream>>>(k, p->pt);              \
            else k_stokes2d<false, false, T_><<<grd, blk, smem, ctx->stream>>>(k, p->pt);                  \
        }                                                                                                  \
    }
    K2_TILES(X)
#undef X
    ctx->launches++;
    JR_CHECK_LAUNCH();
    if (p->periodic) {
        const int32_t none[6] = {0, 0, 0, 0, 0, 0};
        int st = launch_flow_bcs2d(ctx, O[S_Vx], O[S_Vy], p->nx, p->ny, none, none, p->pe);
        if (st) return st;
    }
    return JR_OK;
}

// bring the final state (set niter & 1) back into the caller's arrays
static int plan2_finish(jr_context *ctx, const jr_fields *s, Plan2 *p, int64_t niter)
{
    double *const *Fin = p->set[niter & 1];
    cudaStream_t st = ctx->stream;
    if (niter & 1) {
        double *user[S_COUNT] = {F(Vx), F(Vy), F(P), F(txx), F(tyy), F(txy), nullptr, p->vc ? F(txy_c) : nullptr, nullptr, nullptr,
                                 p->vc ? F(eta) : nullptr, p->vc ? F(etav) : nullptr, p->inc ? F(Ux) : nullptr, p->inc ? F(Uy) : nullptr};
        for (int q = 0; q < S_COUNT; q++)
            if (user[q] && user[q] != Fin[q]) JR_CUDA(cudaMemcpyAsync(user[q], Fin[q], p->bytes[q], cudaMemcpyDeviceToDevice, st));
    }
    if (p->vc) {  // expose the solver-local λ, λv (the oracle does the same for parity checks)
        if (F(lam)) JR_CUDA(cudaMemcpyAsync(F(lam), Fin[S_lam], p->bytes[S_lam], cudaMemcpyDeviceToDevice, st));
        if (F(lamv)) JR_CUDA(cudaMemcpyAsync(F(lamv), Fin[S_lamv], p->bytes[S_lamv], cudaMemcpyDeviceToDevice, st));
    }
    return JR_OK;
}

static inline dim3 grid2(int nx, int ny) { return dim3((nx + 31) / 32, (ny + 7) / 8); }

static int launch_res2d(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, Plan2 *p, int64_t iter)
{
    double *const *C = p->set[iter & 1];
    k_res2d<<<grid2(p->nx, p->ny), dim3(32, 8), 0, ctx->stream>>>(p->nx, p->ny, o->_di[0], o->_di[1], p->vc ? 1 : 0, p->vc ? p->pt.fs : 0.0, C[S_P], C[S_txx],
                                                                  C[S_tyy], C[S_txy], F(rhogx), F(rhogy), C[S_Vy], F(Rx), F(Ry));
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

// errs = [norm_Rx, norm_Ry, norm_∇V]  Stokes2D.jl:273-300, 790-823 (‖R‖₂ / √N, quirk Q4)
static int norms2d(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, Plan2 *p, int64_t iter, double e[3])
{
    int st = launch_res2d(ctx, s, o, p, iter);
    if (st) return st;
    void *slots_v = nullptr;
    if ((st = jr_ctx_scratch(ctx, "norm_slots", 16 * sizeof(double), &slots_v))) return st;
    double *slots = (double *)slots_v;
    const int nx = p->nx, ny = p->ny;
    const int32_t nRx[3] = {nx - 1, ny, 1}, nRy[3] = {nx, ny - 1, 1}, nP[3] = {nx, ny, 1};
    if ((st = jr_launch_sumsq(ctx, F(Rx), nRx, 1, slots + 0))) return st;
    if ((st = jr_launch_sumsq(ctx, F(Ry), nRy, 1, slots + 1))) return st;
    if ((st = jr_launch_sumsq(ctx, F(RP), nP, 0, slots + 2))) return st;
    JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, slots, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    const double gx = o->n_g[0], gy = o->n_g[1];
    e[0] = sqrt(ctx->h_pinned[0]) / sqrt((gx - 2) * (gy - 1));
    e[1] = sqrt(ctx->h_pinned[1]) / sqrt((gx - 1) * (gy - 2));
    e[2] = sqrt(ctx->h_pinned[2]) / sqrt(gx * gy);
    return JR_OK;
}

// τ_o ← τ  (multi_copy!)
static int multi_copy2(jr_context *ctx, const jr_fields *s)
{
    const size_t nc = (size_t)s->n[0] * s->n[1] * 8, nv = (size_t)(s->n[0] + 1) * (s->n[1] + 1) * 8;
    cudaStream_t st = ctx->stream;
    JR_CUDA(cudaMemcpyAsync(F(txx_o), F(txx), nc, cudaMemcpyDeviceToDevice, st));
    JR_CUDA(cudaMemcpyAsync(F(tyy_o), F(tyy), nc, cudaMemcpyDeviceToDevice, st));
    JR_CUDA(cudaMemcpyAsync(F(txy_o), F(txy), nv, cudaMemcpyDeviceToDevice, st));
    if (F(txy_c) && F(txy_o_c)) JR_CUDA(cudaMemcpyAsync(F(txy_o_c), F(txy_c), nc, cudaMemcpyDeviceToDevice, st));
    return JR_OK;
}

static int pre_V2(jr_context *ctx, const jr_fields *s)
{
    const int32_t w[3] = {1, 1, 0}, n[3] = {s->n[0], s->n[1], 1};
    return jr_launch_maxloc3d(ctx, F(etatau), F(eta), n, w);  // ητ = deepcopy(η); compute_maxloc!  Stokes2D.jl:214-216
}

static inline double jr_inv_host(double x) { return 1.0 / x; }
__global__ void k_scale2(size_t n, double *__restrict__ dst, const double *__restrict__ src, double f)
{
    const size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (q < n) dst[q] = src[q] * f;
}

static int pre_VC(jr_context *ctx, const jr_fields *s, Plan2 *p)
{
    const size_t nc = p->bytes[S_P], nv = p->bytes[S_txy];
    cudaStream_t st = ctx->stream;
    if (p->k.dbc) {   // displacement2velocity!(stokes, dt, flow_bcs::DisplacementBoundaryConditions)  Stokes2D.jl:647 ; types/displacement.jl:33-70
        const size_t nVx = p->bytes[S_Vx] / 8, nVy = p->bytes[S_Vy] / 8;
        k_scale2<<<(unsigned)((nVx + 255) / 256), 256, 0, st>>>(nVx, F(Vx), F(Ux), jr_inv_host(p->k.dt));
        k_scale2<<<(unsigned)((nVy + 255) / 256), 256, 0, st>>>(nVy, F(Vy), F(Uy), jr_inv_host(p->k.dt));
        ctx->launches += 2;
    }
    JR_CUDA(cudaMemcpyAsync(F(P0), F(P), nc, cudaMemcpyDeviceToDevice, st));              // @copy stokes.P0 stokes.P   :609
    JR_CUDA(cudaMemcpyAsync(p->set[0][S_th], F(P), nc, cudaMemcpyDeviceToDevice, st));    // θ = deepcopy(stokes.P)     :635
    JR_CUDA(cudaMemsetAsync(p->set[0][S_lam], 0, nc, st));                                 // λ, λv = 0                  :636-637
    JR_CUDA(cudaMemsetAsync(p->set[0][S_lamv], 0, nv, st));
    if (!F(etav)) JR_CUDA(cudaMemsetAsync(p->set[0][S_etav], 0, nv, st));
    JR_CUDA(cudaMemsetAsync(F(pxx), 0, nc, st));                                           // @tensor_center(ε_pl) .= 0  :641-643
    JR_CUDA(cudaMemsetAsync(F(pyy), 0, nc, st));
    if (F(pxy_c)) JR_CUDA(cudaMemsetAsync(F(pxy_c), 0, nc, st));
    k_rhog2d<<<grid2(p->nx, p->ny), dim3(32, 8), 0, st>>>(p->nx, p->ny, p->pt, p->k.ph_c, F(T), F(Pargs), F(rhogx), F(rhogy));   // compute_ρg!  :646
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

// exit kernels of the VC solve  Stokes2D.jl:840-857
static int post_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o)
{
    const int nx = s->n[0], ny = s->n[1];
    cudaStream_t st = ctx->stream;
    const dim3 blk(32, 8);
    if (F(wxy)) { k_vorticity2d<<<grid2(nx + 1, ny + 1), blk, 0, st>>>(nx, ny, o->_di[0], o->_di[1], F(Vx), F(Vy), F(wxy)); ctx->launches++; }
    if (F(exy_c)) { k_shear2center2d<<<grid2(nx, ny), blk, 0, st>>>(nx, ny, F(exy_c), F(exy)); ctx->launches++; }
    if (F(pxy_c)) { k_shear2center2d<<<grid2(nx, ny), blk, 0, st>>>(nx, ny, F(pxy_c), F(pxy)); ctx->launches++; }
    if (F(dxy_c) && F(dxy)) { k_shear2center2d<<<grid2(nx, ny), blk, 0, st>>>(nx, ny, F(dxy_c), F(dxy)); ctx->launches++; }
    k_inv_stag2d<<<grid2(nx, ny), blk, 0, st>>>(nx, ny, F(EII_pl), F(pxx), F(pyy), F(pxy), 1, o->dt);   // accumulate_tensor!
    const size_t n = (size_t)nx * ny;
    k_axpy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, F(EVol_pl), F(e_vol_pl), o->dt);             // accumulate_vol!
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    return multi_copy2(ctx, s);
}

static void fill_result(jr_context *ctx, jr_stokes_result *res, int64_t iter, int64_t cont, double err)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    res->iter = iter; res->nhist = cont; res->err = err;
    res->time_s = ms * 1e-3;
    res->kernel_launches = ctx->launches;
}

extern "C" {

int jr_stokes2d_iterate_V2(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int64_t niter, jr_stokes_result *res)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    int st = check2d(s, o, false, nullptr);
    if (st) return st;
    if ((st = single_rank2d(ctx))) return st;
    JR_CUDA(cudaSetDevice(ctx->device));
    Plan2 p;
    ctx->launches = 0;
    if ((st = plan2_begin(ctx, s, o, false, nullptr, &p))) return st;
    if ((st = pre_V2(ctx, s))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) {
        for (int64_t it = 0; it < niter; it++)
            if ((st = plan2_iter(ctx, &p, it, true))) return st;
    } else if (niter >= 1) {
        if ((st = plan2_batch(ctx, &p, 0, niter - 1))) return st;
        if ((st = plan2_iter(ctx, &p, niter - 1, true))) return st;
    }
    if ((st = plan2_finish(ctx, s, &p, niter))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if ((st = launch_res2d(ctx, s, o, &p, 0))) return st;  // state is back in the caller's arrays (set 0)
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res) fill_result(ctx, res, niter, 0, NAN);
    return JR_OK;
}

int jr_stokes2d_solve_V2(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, jr_stokes_result *res)
{
    JR_REQUIRE(ctx && res, JR_ERR_ARG, "null context/result");
    int st = check2d(s, o, false, nullptr);
    if (st) return st;
    if ((st = single_rank2d(ctx))) return st;
    JR_REQUIRE(res->err_evo1 && res->err_evo2 && res->norm_Rx && res->norm_Ry && res->norm_divV, JR_ERR_ARG, "result history arrays must be provided");
    JR_CUDA(cudaSetDevice(ctx->device));
    Plan2 p;
    ctx->launches = 0;
    if ((st = plan2_begin(ctx, s, o, false, nullptr, &p))) return st;
    if ((st = pre_V2(ctx, s))) return st;
    double err_it1 = 1.0, err = 1.0;
    int64_t iter = 0, cont = 0;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    while (iter < 2 || (((err / err_it1) > o->eps_rel && err > o->eps_abs) && iter <= o->iterMax)) {  // Stokes2D.jl:222
        // the iterations up to the next sample (or iterMax) whose result nobody can observe: one resident batch.  The loop condition
        // cannot change in between (err is only updated at samples; iter stays ≤ iterMax)
        if (p.rplan.ok && iter >= 1) {
            const int64_t nb = std::min<int64_t>(o->nout - 1 - (iter % o->nout), o->iterMax - iter);
            if (nb >= 2) {
                if ((st = plan2_batch(ctx, &p, iter, nb))) return st;
                iter += nb;
                continue;
            }
        }
        const int64_t next = iter + 1;
        const bool diag = (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) || (next % o->nout == 0) || next > o->iterMax || next < 2;
        if ((st = plan2_iter(ctx, &p, iter, diag))) return st;
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            double e[3];
            if ((st = norms2d(ctx, s, o, &p, iter, e))) return st;
            res->norm_Rx[cont] = e[0]; res->norm_Ry[cont] = e[1]; res->norm_divV[cont] = e[2];
            err = fmax(fmax(e[0], e[1]), e[2]);
            if (std::isnan(e[0]) || std::isnan(e[1]) || std::isnan(e[2])) err = NAN;
            res->err_evo1[cont] = err; res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), res->norm_divV[0]);
            // no isnan test in this variant (Stokes2D.jl:222-311): a NaN error fails the loop condition, the loop ends quietly,
            // multi_copy! runs and the history (with the NaN) is returned
        }
    }
    if ((st = plan2_finish(ctx, s, &p, iter))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if ((st = multi_copy2(ctx, s))) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    fill_result(ctx, res, iter, cont, err);
    return JR_OK;
}

int jr_stokes2d_iterate_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, int64_t niter, int finish,
                           jr_stokes_result *res)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    int st = check2d(s, o, true, vc);
    if (st) return st;
    if ((st = single_rank2d(ctx))) return st;
    JR_CUDA(cudaSetDevice(ctx->device));
    Plan2 p;
    ctx->launches = 0;
    if ((st = plan2_begin(ctx, s, o, true, vc, &p))) return st;
    if ((st = pre_VC(ctx, s, &p))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int64_t it = 0; it < niter; it++)
        if ((st = plan2_iter(ctx, &p, it, (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) || it == niter - 1))) return st;
    if ((st = plan2_finish(ctx, s, &p, niter))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if ((st = launch_res2d(ctx, s, o, &p, 0))) return st;
    if (finish && (st = post_VC(ctx, s, o))) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res) fill_result(ctx, res, niter, 0, NAN);
    return JR_OK;
}

int jr_stokes2d_solve_VC(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, jr_stokes_result *res)
{
    JR_REQUIRE(ctx && res, JR_ERR_ARG, "null context/result");
    int st = check2d(s, o, true, vc);
    if (st) return st;
    if ((st = single_rank2d(ctx))) return st;
    JR_REQUIRE(res->err_evo1 && res->err_evo2 && res->norm_Rx && res->norm_Ry && res->norm_divV, JR_ERR_ARG, "result history arrays must be provided");
    JR_CUDA(cudaSetDevice(ctx->device));
    Plan2 p;
    ctx->launches = 0;
    if ((st = plan2_begin(ctx, s, o, true, vc, &p))) return st;
    if ((st = pre_VC(ctx, s, &p))) return st;
    double err_it1 = 1.0, err = 1.0;
    int64_t iter = 0, cont = 0;
    int status = JR_OK;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    while (iter <= o->iterMax) {  // Stokes2D.jl:649-652
        const bool conv = ((err / err_it1) < o->eps_rel || err < o->eps_abs);
        if (o->iterMin < iter && conv) break;
        const int64_t next = iter + 1;
        // the loop can end after a sample, beyond iterMax, or — once the tolerance is met — as soon as iter > iterMin
        const bool diag = (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) || (next % o->nout == 0) || next > o->iterMax || conv;
        if ((st = plan2_iter(ctx, &p, iter, diag))) return st;
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            double e[3];
            if ((st = norms2d(ctx, s, o, &p, iter, e))) return st;
            res->norm_Rx[cont] = e[0]; res->norm_Ry[cont] = e[1]; res->norm_divV[cont] = e[2];
            err = fmax(fmax(e[0], e[1]), e[2]);
            if (std::isnan(e[0]) || std::isnan(e[1]) || std::isnan(e[2])) err = NAN;
            res->err_evo1[cont] = err; res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), res->norm_divV[0]);
            if (std::isnan(err)) { status = JR_ERR_NAN; break; }  // isnan(err) && error("NaN(s)")  Stokes2D.jl:836
        }
    }
    if ((st = plan2_finish(ctx, s, &p, iter))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if (status == JR_OK && (st = post_VC(ctx, s, o))) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    fill_result(ctx, res, iter, cont, err);
    if (status) jr_set_error("NaN(s)");
    return status;
}

int jr_flow_bcs2d(jr_context *ctx, double *Ax, double *Ay, const int32_t n[3], const int32_t free_slip[6], const int32_t no_slip[6],
                  const int32_t periodic[6])
{
    JR_REQUIRE(ctx && Ax && Ay && n, JR_ERR_ARG, "jr_flow_bcs2d: null argument");
    JR_REQUIRE(n[0] >= 2 && n[1] >= 2, JR_ERR_SHAPE, "jr_flow_bcs2d: grid too small");
    int st = launch_flow_bcs2d(ctx, Ax, Ay, n[0], n[1], free_slip, no_slip, periodic);
    if (st) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_compute_viscosity2d(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, const jr_vc_inputs *vc, double nu)
{
    JR_REQUIRE(ctx && s && o && vc && F(eta) && vc->ph_center, JR_ERR_ARG, "jr_compute_viscosity2d: null argument");
    jr_phase_tab pt;
    int st = jr_make_phase_tab(vc, &pt);
    if (st) return st;
    const size_t nc = (size_t)s->n[0] * s->n[1], nv = (size_t)(s->n[0] + 1) * (s->n[1] + 1);
    k_viscosity2d<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(nc, pt, vc->ph_center, F(eta), nu, o->visc_cutoff_lo, o->visc_cutoff_hi);
    ctx->launches++;
    if (F(etav) && vc->ph_vertex) {
        k_viscosity2d<<<(unsigned)((nv + 255) / 256), 256, 0, ctx->stream>>>(nv, pt, vc->ph_vertex, F(etav), nu, o->visc_cutoff_lo, o->visc_cutoff_hi);
        ctx->launches++;
    }
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_compute_rhog2d(jr_context *ctx, const jr_fields *s, const jr_vc_inputs *vc)
{
    JR_REQUIRE(ctx && s && vc && F(rhogx) && F(rhogy) && vc->ph_center, JR_ERR_ARG, "jr_compute_rhog2d: null argument");
    jr_phase_tab pt;
    int st = jr_make_phase_tab(vc, &pt);
    if (st) return st;
    k_rhog2d<<<grid2(s->n[0], s->n[1]), dim3(32, 8), 0, ctx->stream>>>(s->n[0], s->n[1], pt, vc->ph_center, F(T), F(Pargs), F(rhogx), F(rhogy));
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_tensor_invariant2d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *xy, const int32_t n[3])
{
    JR_REQUIRE(ctx && II && xx && yy && xy && n, JR_ERR_ARG, "jr_tensor_invariant2d: null argument");
    k_inv_stag2d<<<grid2(n[0], n[1]), dim3(32, 8), 0, ctx->stream>>>(n[0], n[1], II, xx, yy, xy, 0, 0.0);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_accumulate_tensor2d(jr_context *ctx, double *II, const double *xx, const double *yy, const double *xy, const int32_t n[3], double dt)
{
    JR_REQUIRE(ctx && II && xx && yy && xy && n, JR_ERR_ARG, "jr_accumulate_tensor2d: null argument");
    k_inv_stag2d<<<grid2(n[0], n[1]), dim3(32, 8), 0, ctx->stream>>>(n[0], n[1], II, xx, yy, xy, 1, dt);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

}  // extern "C"
