// stokes2d_resident.cu — 2D-V2 PT iterations with the state RESIDENT IN SHARED MEMORY (sm_100a: 227 KB × 148 SMs = 33 MB on chip).
//
// What it replaces: the non-observable iterations of the 2D-V2 loop (src/stokes/Stokes2D.jl:222-311: compute_∇V! → compute_P! →
// compute_strain_rate! → compute_τ! → compute_V! → flow_bcs!) when the whole state fits on chip — config 2 (SolCx 511², 31 MB/iteration
// in the fused one-launch-per-iteration kernel of stokes2d.cu, which is bound by launch + tile latency: 14.4 µs per iteration).
//
// Design: the grid is cut into gx × gy ≤ 148 tiles, ONE persistent CTA per SM (cooperative launch).  A CTA loads its tile of P, τxx,
// τyy, τxy, Vx, Vy, η, ητ, ρg plus a one-cell rim ONCE, then runs all iterations of the batch out of shared memory:
//   A  new P, τxx, τyy for every cell of the tile + rim (the rim is recomputed redundantly, bit-identically, by both neighbours),
//      new τxy for every vertex of the tile — in place (a cell / vertex only reads its own old value and the velocities);
//   B  new Vx, Vy on the faces the tile owns (in place: a face only reads its own old value), boundary ghosts by the tile that owns
//      the adjacent interior face (free slip / no slip images; prescribed ghosts keep their value);
//   C  only VELOCITIES travel: the faces within two of a tile edge go to the dense global arrays (ping-pong by iteration parity), a
//      release flag per tile, and every tile waits for its ≤ 8 neighbours only (no grid-wide barrier), then reads its halo faces (L2);
//      the cells / vertices of the next iteration that read owned velocities only (≈ 85 % of a tile) are updated BEFORE the wait.
// Stresses and pressure never leave the SM until the batch ends.  Observable iterations (every nout and the last) run the regular
// fused kernel with diagnostics, so residuals / histories are the reference's.
// Arithmetic: operation for operation the V2 path of k_stokes2d for the case this kernel accepts — 1/(G·dt) = 1/(K·dt) = 0 and
// Q/dt = 0 everywhere (G = K = Inf as in SolCx / SolKz, or dt = Inf) — so results are bit-comparable with it and with the oracle.
#include "common.cuh"
#include "tma.cuh"
#include "stokes2d_resident.cuh"

#define RX 64     // threads in x = the widest window a tile may have (Vx: cx + 4 columns)
#define RYMAX 16  // threads in y (blockDim.y ≤ RYMAX)

struct V2R {
    int nx, ny, gx, gy, cx, cy;
    double _dx, _dy, r, th, edt;
    int fs_l, fs_r, fs_t, fs_b, ns_l, ns_r, ns_t, ns_b;
    double *Vx[2], *Vy[2], *P[2], *txx[2], *tyy[2], *txy[2];
    const double *eta, *ett, *rgx, *rgy;
    unsigned long long *flags;
    unsigned long long flag_base;
    long long it0;
    int niter;
};

// Thread (tx, ty) owns window COLUMN tx of every array and walks the rows ty, ty + blockDim.y, …: no index division, and every
// ownership / validity test splits into a per-thread column part (hoisted out of the iteration loop) and a cheap row part.
__global__ void __launch_bounds__(RX * RYMAX, 1) k_v2_resident(const __grid_constant__ V2R a)
{
    extern __shared__ double sm[];
    const int tx = threadIdx.x, ty = threadIdx.y, BR = blockDim.y, tid = ty * RX + tx, RT = RX * BR;
    const int nx = a.nx, ny = a.ny;
    const int bx = blockIdx.x % a.gx, by = blockIdx.x / a.gx;
    const int i0 = bx * a.cx + 1, i1 = min(i0 + a.cx - 1, nx), j0 = by * a.cy + 1, j1 = min(j0 + a.cy - 1, ny);
    // windows (fixed strides from the nominal tile size); all share the origin (i0 − 1, j0 − 1) except the vertices (i0, j0)
    const int WC = a.cx + 2, HC = a.cy + 2;  // cells  [i0−1, i1+1] × [j0−1, j1+1]
    const int WV = a.cx + 1, HV = a.cy + 1;  // vertices [i0, i1+1] × [j0, j1+1]
    const int WX = a.cx + 4, HX = a.cy + 2;  // Vx: faces [i0−1, i1+2] × cell rows [j0−1, j1+1] (rows 0 and ny+1 = ghost rows)
    const int WY = a.cx + 2, HY = a.cy + 4;  // Vy: cell columns [i0−1, i1+1] (0 and nx+1 = ghost columns) × faces [j0−1, j1+2]
    const int NC = WC * HC, NV = WV * HV, NX = WX * HX, NY = WY * HY;
    double *sP = sm, *sTxx = sP + NC, *sTyy = sTxx + NC, *sEta = sTyy + NC, *sEtt = sEta + NC, *sPsi = sEtt + NC, *sRgx = sPsi + NC,
           *sRgy = sRgx + NC, *sTxy = sRgy + NC, *sEtav = sTxy + NV, *sVx = sEtav + NV, *sVy = sVx + NX;
    const int in0 = (int)(a.it0 & 1);
    const bool last_x = i1 == nx, last_y = j1 == ny;

    // ---- load the tile (once) -------------------------------------------------------------------------------------------------
    for (int e = tid; e < NC; e += RT) {
        const int ci = i0 - 1 + e % WC, cj = j0 - 1 + e / WC;
        double p = 0, xx = 0, yy = 0, et = 1, tt = 1, fx = 0, fy = 0, psi = 0;
        if (ci >= 1 && ci <= nx && cj >= 1 && cj <= ny) {
            const size_t c = IX2(nx, ci, cj);
            p = a.P[in0][c]; xx = a.txx[in0][c]; yy = a.tyy[in0][c]; et = a.eta[c]; tt = a.ett[c]; fx = a.rgx[c]; fy = a.rgy[c];
            psi = jr_inv(jr_inv(tt) + 0.0) * a.r / a.th;  // compute_P! with ητ (quirk Q5), 1/(G dt) = 0  PressureKernels.jl:186-195
        }
        sP[e] = p; sTxx[e] = xx; sTyy[e] = yy; sEta[e] = et; sEtt[e] = tt; sPsi[e] = psi; sRgx[e] = fx; sRgy[e] = fy;
    }
    for (int e = tid; e < NV; e += RT) {
        const int vi = i0 + e % WV, vj = j0 + e / WV;
        double t = 0, ev = 1;
        if (vi <= nx + 1 && vj <= ny + 1) {
            t = a.txy[in0][IX2(nx + 1, vi, vj)];
            // _av_ai_clamped  MiniKernels.jl:76-80 (η is constant in this variant)
            const int ia = jr_clamp(vi - 1, 1, nx), ib = jr_clamp(vi, 1, nx), ja = jr_clamp(vj - 1, 1, ny), jb = jr_clamp(vj, 1, ny);
            ev = 0.25 * (a.eta[IX2(nx, ia, ja)] + a.eta[IX2(nx, ib, ja)] + a.eta[IX2(nx, ia, jb)] + a.eta[IX2(nx, ib, jb)]);
        }
        sTxy[e] = t; sEtav[e] = ev;
    }
    for (int e = tid; e < NX; e += RT) {
        const int fi = i0 - 1 + e % WX, rj = j0 - 1 + e / WX;
        sVx[e] = (fi >= 1 && fi <= nx + 1 && rj >= 0 && rj <= ny + 1) ? a.Vx[in0][IX2(nx + 1, fi, rj + 1)] : 0.0;
    }
    for (int e = tid; e < NY; e += RT) {
        const int ck = i0 - 1 + e % WY, fj = j0 - 1 + e / WY;
        sVy[e] = (ck >= 0 && ck <= nx + 1 && fj >= 1 && fj <= ny + 1) ? a.Vy[in0][IX2(nx + 2, ck + 1, fj)] : 0.0;
    }
    __syncthreads();

    const double _dx = a._dx, _dy = a._dy, edt = a.edt;
    const double inv3 = jr_inv(3.0);
    const double dtr = jr_inv(a.th + 1.0);  // compute_dτ_r with 1/(G dt) = 0: fma(η, 0, 1) = 1
    // neighbour tiles (threads 0..7 each watch one)
    int nb_tile = -1;
    if (tid < 8) {
        const int q = tid < 4 ? tid : tid + 1;  // skip the centre of the 3 × 3 neighbourhood
        const int tx2 = bx + q % 3 - 1, ty2 = by + q / 3 - 1;
        if (tx2 >= 0 && tx2 < a.gx && ty2 >= 0 && ty2 < a.gy) nb_tile = ty2 * a.gx + tx2;
    }
    // ownership of a velocity entry = (x part) && (y part)
    auto fxo = [&](int fi) { return (fi >= i0 && fi <= i1) || (fi == nx + 1 && last_x); };                                   // Vx faces
    auto cxo = [&](int ck) { return (ck >= i0 && ck <= i1) || (ck == 0 && i0 == 1) || (ck == nx + 1 && last_x); };         // Vy columns
    auto ryo = [&](int rj) { return (rj >= j0 && rj <= j1) || (rj == 0 && j0 == 1) || (rj == ny + 1 && last_y); };         // Vx rows
    auto fyo = [&](int fj) { return (fj >= j0 && fj <= j1) || (fj == ny + 1 && last_y); };                                   // Vy faces
    // ---- per-thread column facts (iteration-invariant) ----
    const int cw_ = i0 - 1 + tx;  // cell / Vx-face / Vy-column index of window column tx
    const bool c_val = tx < WC && cw_ >= 1 && cw_ <= nx && cw_ <= i1 + 1;                 // cell column
    const bool c_inx = fxo(cw_) && fxo(cw_ + 1) && cxo(cw_);                              // … reads owned velocities only (x part)
    const int vi = i0 + tx;
    const bool v_val = tx < WV && vi <= nx + 1 && vi <= i1 + 1;                           // vertex column
    const bool v_inx = fxo(vi) && cxo(vi) && cxo(vi - 1);
    const int nfx = (i1 - i0 + 1) + (last_x ? 1 : 0), nrow = j1 - j0 + 1, ncol = i1 - i0 + 1, nfy = (j1 - j0 + 1) + (last_y ? 1 : 0);
    const bool bx_val = tx < nfx, bx_int = vi >= 2 && vi <= nx, bx_strip = vi <= i0 + 1 || vi >= i1;   // owned Vx face column fi = vi
    const bool by_val = tx < ncol, by_strip = vi <= i0 || vi >= i1;                                     // owned Vy column ck = vi
    const bool hx_val = tx < WX && cw_ >= 1 && cw_ <= nx + 1 && cw_ <= i1 + 2, hx_own = fxo(cw_);      // Vx window column fi = cw_
    const bool hy_val = tx < WY && cw_ >= 0 && cw_ <= nx + 1 && cw_ <= i1 + 1, hy_own = cxo(cw_);      // Vy window column ck = cw_
    const bool nzero_x = (vi == 1 && a.ns_l) || (vi == nx + 1 && a.ns_r);  // no_slip! zeroes the boundary-normal face (all rows)

    // ---- CTA-uniform row ranges: the row part of every test, so that the loops below carry no per-element predicate ----
    const int c_lo = max(j0 - 1, 1), c_hi = min(j1 + 1, ny);          // cell rows of tile + rim
    const int ci_lo = j0, ci_hi = last_y ? j1 : j1 - 1;               // … whose four faces are owned (y part)
    const int v_hi = j1 + 1;                                           // vertex rows [j0, j1 + 1]
    const int vi_lo = j0 == 1 ? 1 : j0 + 1, vi_hi = last_y ? j1 + 1 : j1;   // … that read owned velocities only (y part)
    // A: pressure and normal stresses (cells of tile + rim), shear stress (vertices of the tile)
    auto cell_row = [&](int cj) {
        const int lj = cj - (j0 - 1);
        const int e = lj * WC + tx, lx = lj * WX + tx, ly = lj * WY + tx;
        const double dVx = (-sVx[lx] + sVx[lx + 1]) * _dx;
        const double dVy = (-sVy[ly] + sVy[ly + WY]) * _dy;
        const double divV = dVx + dVy;                     // compute_∇V!  VelocityKernels.jl:3-6
        const double dV = divV * inv3;
        const double exx = dVx - dV, eyy = dVy - dV;       // compute_strain_rate!  VelocityKernels.jl:10-44
        const double eta = sEta[e], txx = sTxx[e], tyy = sTyy[e];
        sP[e] = (-divV + 0.0) * sPsi[e] + sP[e];           // compute_P!: (P0/(K dt) − ∇V + Q/dt)·ψ + P, /(1 + ψ/(K dt)) = /1
        sTxx[e] = txx + dtr * fma(2.0 * eta, exx, -txx);   // compute_τ!  StressKernels.jl:63-91 with 1/(G dt) = 0
        sTyy[e] = tyy + dtr * fma(2.0 * eta, eyy, -tyy);
    };
    auto vert_row = [&](int vj) {
        const int lj = vj - j0;
        const int e = lj * WV + tx, lx = (lj + 1) * WX + tx + 1, ly = (lj + 1) * WY + tx + 1;   // Vx(vi, vj), Vy(vi, vj)
        const double exy = 0.5 * (_dy * (sVx[lx] - sVx[lx - WX]) + _dx * (sVy[ly] - sVy[ly - 1]));
        const double t0 = sTxy[e];
        sTxy[e] = t0 + dtr * fma(2.0 * sEtav[e], exy, -t0);
    };
    // part 1: the elements that read owned velocities only (they do not wait for the neighbours); part 2: the others
    auto phase_A_inner = [&]() {
        if (c_val && c_inx)
            for (int cj = ci_lo + ty; cj <= ci_hi; cj += BR) cell_row(cj);
        if (v_val && v_inx)
            for (int vj = vi_lo + ty; vj <= vi_hi; vj += BR) vert_row(vj);
    };
    auto phase_A_outer = [&]() {
        if (c_val) {
            if (c_inx) {   // an inner column: only the rows next to the tile's y edges are left (≤ 3)
                const int nlo = ci_lo - c_lo, nhi = c_hi - ci_hi;
                for (int q = ty; q < nlo + nhi; q += BR) cell_row(q < nlo ? c_lo + q : ci_hi + 1 + (q - nlo));
            } else
                for (int cj = c_lo + ty; cj <= c_hi; cj += BR) cell_row(cj);
        }
        if (v_val) {
            if (v_inx) {
                const int nlo = vi_lo - j0, nhi = v_hi - vi_hi;
                for (int q = ty; q < nlo + nhi; q += BR) vert_row(q < nlo ? j0 + q : vi_hi + 1 + (q - nlo));
            } else
                for (int vj = j0 + ty; vj <= v_hi; vj += BR) vert_row(vj);
        }
    };

    phase_A_inner();
    phase_A_outer();
    for (int it = 0; it < a.niter; ++it) {
        const int outq = (int)((a.it0 + it + 1) & 1);
        double *const Vxo = a.Vx[outq], *const Vyo = a.Vy[outq];
        __syncthreads();
        // ---- B: velocities on the owned faces + their boundary ghosts; faces near a tile edge also go to the dense arrays ----------
        if (bx_val)
            for (int lr = ty; lr < nrow; lr += BR) {
                const int rj = j0 + lr, fi = vi;
                const int l = (lr + 1) * WX + tx + 1;
                const double v0 = sVx[l];
                double vx;
                if (bx_int) {  // compute_V!  VelocityKernels.jl:108-131
                    const int cw = (lr + 1) * WC + tx, ce = cw + 1, lv = lr * WV + tx;
                    const double dP = (-sP[cw] + sP[ce]) * _dx, dt_xx = (-sTxx[cw] + sTxx[ce]) * _dx;
                    const double dt_xy = (-sTxy[lv] + sTxy[lv + WV]) * _dy;
                    const double avf = (sRgx[cw] + sRgx[ce]) * 0.5, ave = (sEtt[cw] + sEtt[ce]) * 0.5;
                    vx = v0 + jr_div_nr((-dP + dt_xx + dt_xy - avf) * edt, ave);
                } else
                    vx = ((fi == 1) ? a.ns_l : a.ns_r) ? 0.0 : v0;
                sVx[l] = vx;
                if (bx_strip || rj <= j0 || rj >= j1) Vxo[IX2(nx + 1, fi, rj + 1)] = vx;
                if (rj == 1) {   // flow_bcs! (no_slip! → free_slip!) as a gather: bottom ghost row
                    const int g = l - WX;
                    const double gv = a.fs_b ? vx : (a.ns_b ? -vx : (nzero_x ? 0.0 : sVx[g]));
                    sVx[g] = gv;
                    Vxo[IX2(nx + 1, fi, 1)] = gv;
                }
                if (rj == ny) {  // top ghost row
                    const int g = l + WX;
                    const double gv = a.fs_t ? vx : (a.ns_t ? -vx : (nzero_x ? 0.0 : sVx[g]));
                    sVx[g] = gv;
                    Vxo[IX2(nx + 1, fi, ny + 2)] = gv;
                }
            }
        if (by_val)
            for (int lf = ty; lf < nfy; lf += BR) {
                const int fj = j0 + lf, ck = vi;
                const int l = (lf + 1) * WY + tx + 1;
                const double v0 = sVy[l];
                double vy;
                if (fj >= 2 && fj <= ny) {
                    const int cs = lf * WC + tx + 1, cn = cs + WC, lv = lf * WV + tx;
                    const double dP = (-sP[cs] + sP[cn]) * _dy, dt_yy = (-sTyy[cs] + sTyy[cn]) * _dy;
                    const double dt_xy = (-sTxy[lv] + sTxy[lv + 1]) * _dx;
                    const double avf = (sRgy[cs] + sRgy[cn]) * 0.5, ave = (sEtt[cs] + sEtt[cn]) * 0.5;
                    vy = v0 + jr_div_nr((-dP + dt_yy + dt_xy - avf) * edt, ave);
                } else
                    vy = ((fj == 1) ? a.ns_b : a.ns_t) ? 0.0 : v0;
                sVy[l] = vy;
                if (by_strip || fj <= j0 + 1 || fj >= j1) Vyo[IX2(nx + 2, ck + 1, fj)] = vy;
                const bool nzero = (fj == 1 && a.ns_b) || (fj == ny + 1 && a.ns_t);
                if (ck == 1) {   // left ghost column
                    const int g = l - 1;
                    const double gv = a.fs_l ? vy : (a.ns_l ? -vy : (nzero ? 0.0 : sVy[g]));
                    sVy[g] = gv;
                    Vyo[IX2(nx + 2, 1, fj)] = gv;
                }
                if (ck == nx) {  // right ghost column
                    const int g = l + 1;
                    const double gv = a.fs_r ? vy : (a.ns_r ? -vy : (nzero ? 0.0 : sVy[g]));
                    sVy[g] = gv;
                    Vyo[IX2(nx + 2, nx + 2, fj)] = gv;
                }
            }
        if (it + 1 == a.niter) break;  // the final state is stored below; nobody needs this iteration's halo
        // ---- C: publish the strips, update everything that does not need the neighbours, then wait for them and read the halo faces ---
        __syncthreads();
        const unsigned long long stamp = a.flag_base + (unsigned long long)(it + 1);
        if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.flags + blockIdx.x), "l"(stamp) : "memory");
        phase_A_inner();
        if (nb_tile >= 0) {
            unsigned long long v;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.flags + nb_tile) : "memory");
            } while (v < stamp);
        }
        __syncthreads();
        if (hx_val) {   // Vx window column fi = cw_: a column of the neighbours entirely, or only the rows beyond the tile's y edges
            if (!hx_own) {
                for (int rj = max(j0 - 1, 0) + ty; rj <= min(j1 + 1, ny + 1); rj += BR) sVx[(rj - (j0 - 1)) * WX + tx] = __ldcg(Vxo + IX2(nx + 1, cw_, rj + 1));
            } else {
                if (ty == 0 && j0 > 1) sVx[tx] = __ldcg(Vxo + IX2(nx + 1, cw_, j0));
                if (ty == 1 && !last_y) sVx[(j1 + 2 - j0) * WX + tx] = __ldcg(Vxo + IX2(nx + 1, cw_, j1 + 2));
            }
        }
        if (hy_val) {   // Vy window column ck = cw_
            if (!hy_own) {
                for (int fj = max(j0 - 1, 1) + ty; fj <= min(j1 + 2, ny + 1); fj += BR) sVy[(fj - (j0 - 1)) * WY + tx] = __ldcg(Vyo + IX2(nx + 2, cw_ + 1, fj));
            } else {
                if (ty == 2 && j0 > 1) sVy[tx] = __ldcg(Vyo + IX2(nx + 2, cw_ + 1, j0 - 1));
                if (ty == 3 && !last_y) sVy[(j1 + 2 - j0) * WY + tx] = __ldcg(Vyo + IX2(nx + 2, cw_ + 1, j1 + 1));
                if (ty == 4 && !last_y && j1 + 2 <= ny + 1) sVy[(j1 + 3 - j0) * WY + tx] = __ldcg(Vyo + IX2(nx + 2, cw_ + 1, j1 + 2));
            }
        }
        __syncthreads();
        phase_A_outer();
    }
    __syncthreads();
    // ---- store the owned part of the final state (dense set of the batch's last iteration) ---------------------------------------------
    const int fin = (int)((a.it0 + a.niter) & 1);
    for (int e = tid; e < NC; e += RT) {
        const int ci = i0 - 1 + e % WC, cj = j0 - 1 + e / WC;
        if (ci < i0 || ci > i1 || cj < j0 || cj > j1) continue;
        const size_t c = IX2(nx, ci, cj);
        a.P[fin][c] = sP[e]; a.txx[fin][c] = sTxx[e]; a.tyy[fin][c] = sTyy[e];
    }
    for (int e = tid; e < NV; e += RT) {
        const int wi = i0 + e % WV, wj = j0 + e / WV;
        const bool own_i = wi <= i1 || (wi == nx + 1 && last_x), own_j = wj <= j1 || (wj == ny + 1 && last_y);
        if (own_i && own_j) a.txy[fin][IX2(nx + 1, wi, wj)] = sTxy[e];
    }
    for (int e = tid; e < NX; e += RT) {
        const int fi = i0 - 1 + e % WX, rj = j0 - 1 + e / WX;
        if (fxo(fi) && ryo(rj)) a.Vx[fin][IX2(nx + 1, fi, rj + 1)] = sVx[e];
    }
    for (int e = tid; e < NY; e += RT) {
        const int ck = i0 - 1 + e % WY, fj = j0 - 1 + e / WY;
        if (cxo(ck) && fyo(fj)) a.Vy[fin][IX2(nx + 2, ck + 1, fj)] = sVy[e];
    }
}

// does every cell satisfy 1/(G·dt) = 0, 1/(K·dt) = 0, Q·(1/dt) = 0 ?  (one pass per solve)
__global__ void k_v2_resident_check(const double *__restrict__ G, const double *__restrict__ K, const double *__restrict__ Q, double dt, size_t n,
                                    int *bad)
{
    const double _dt = jr_inv(dt);
    int b = 0;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x)
        if (jr_inv(G[q] * dt) != 0.0 || jr_inv(K[q] * dt) != 0.0 || Q[q] * _dt != 0.0) b = 1;
    if (__syncthreads_or(b) && threadIdx.x == 0) atomicOr(bad, 1);
}

static size_t tile_smem(int cx, int cy)
{
    return ((size_t)8 * (cx + 2) * (cy + 2) + (size_t)2 * (cx + 1) * (cy + 1) + (size_t)(cx + 4) * (cy + 2) + (size_t)(cx + 2) * (cy + 4)) * 8;
}

// tiles gx × gy ≤ SMs minimising the cells a CTA updates per iteration (tile + rim) among the tiles that fit the shared memory of
// one SM; false: the state does not fit on chip
static bool choose_tiles(int nx, int ny, int sms, size_t smem_max, int &gx, int &gy, int &cx, int &cy)
{
    long best = -1;
    for (int tx = 1; tx <= sms; tx++)
        for (int ty = 1; tx * ty <= sms; ty++) {
            const int c0 = (nx + tx - 1) / tx, c1 = (ny + ty - 1) / ty;
            if ((tx > 1 && c0 < 4) || (ty > 1 && c1 < 4)) continue;   // halo faces are two deep: no slivers
            if (c0 + 4 > RX) continue;                                // a thread per window column (Vx: c0 + 4 columns)
            if ((nx + c0 - 1) / c0 != tx || (ny + c1 - 1) / c1 != ty) continue;   // (the same tile size with fewer tiles was seen already)
            if (tile_smem(c0, c1) > smem_max) continue;
            const long work = (long)(c0 + 2) * (c1 + 2);
            if (best < 0 || work < best) { best = work; cx = c0; cy = c1; gx = tx; gy = ty; }
        }
    return best > 0;
}

int jr_v2_resident_plan(jr_context *ctx, const V2ResArgs *r, V2ResPlan *p)
{
    p->ok = false;
    if (const char *e = getenv("JRB200_2D_RESIDENT"))
        if (atoi(e) == 0) return JR_OK;
    int dev_smem = 0;
    JR_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    if (!choose_tiles(r->nx, r->ny, ctx->sm_count, (size_t)dev_smem, p->gx, p->gy, p->cx, p->cy)) return JR_OK;
    // the elastic / compressible terms must vanish identically (G = K = Inf or dt = Inf; Q/dt = 0)
    void *flag = nullptr;
    int st = jr_ctx_scratch(ctx, "v2res_flags", 4096, &flag);
    if (st) return st;
    int *bad = (int *)flag + 1000;  // last ints of the flag page
    JR_CUDA(cudaMemsetAsync(flag, 0, 4096, ctx->stream));
    k_v2_resident_check<<<296, 256, 0, ctx->stream>>>(r->G, r->K, r->Q, r->dt, (size_t)r->nx * r->ny, bad);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    int h = 1;
    JR_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h) return JR_OK;
    p->smem = tile_smem(p->cx, p->cy);
    JR_CUDA(cudaFuncSetAttribute(k_v2_resident, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    int nb = 0;
    p->rows = RYMAX;
    if (const char *e = getenv("JRB200_2D_RESIDENT_ROWS")) p->rows = atoi(e) < 8 ? 8 : (atoi(e) > RYMAX ? RYMAX : atoi(e));   // (the halo read hands rows to ty = 0 … 4)
    JR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_v2_resident, RX * p->rows, p->smem));
    if (nb < 1 || (long)nb * ctx->sm_count < (long)p->gx * p->gy) return JR_OK;
    p->flags = (unsigned long long *)flag;
    p->flag_base = 0;
    p->ok = true;
    if (getenv("JRB200_VERBOSE"))
        fprintf(stderr, "[jrb200] k_v2_resident: %d x %d grid on %d x %d tiles of %d x %d cells, %zu B shared memory per CTA\n", r->nx, r->ny, p->gx,
                p->gy, p->cx, p->cy, p->smem);
    return JR_OK;
}

// iterations it0 … it0 + niter − 1 (state in dense set it0 & 1 → dense set (it0 + niter) & 1), none of them observable
int jr_v2_resident_run(jr_context *ctx, const V2ResArgs *r, V2ResPlan *p, int64_t it0, int niter)
{
    JR_REQUIRE(p->ok && niter >= 1, JR_ERR_ARG, "resident 2D-V2 batch without a plan");
    V2R a;
    a.nx = r->nx; a.ny = r->ny; a.gx = p->gx; a.gy = p->gy; a.cx = p->cx; a.cy = p->cy;
    a._dx = r->_dx; a._dy = r->_dy; a.r = r->r; a.th = r->th; a.edt = r->edt;
    a.fs_l = r->fs_l; a.fs_r = r->fs_r; a.fs_t = r->fs_t; a.fs_b = r->fs_b; a.ns_l = r->ns_l; a.ns_r = r->ns_r; a.ns_t = r->ns_t; a.ns_b = r->ns_b;
    for (int q = 0; q < 2; q++) {
        a.Vx[q] = r->Vx[q]; a.Vy[q] = r->Vy[q]; a.P[q] = r->P[q]; a.txx[q] = r->txx[q]; a.tyy[q] = r->tyy[q]; a.txy[q] = r->txy[q];
    }
    a.eta = r->eta; a.ett = r->etatau; a.rgx = r->rhogx; a.rgy = r->rhogy;
    a.flags = p->flags; a.flag_base = p->flag_base;
    p->flag_base += (unsigned long long)niter;
    a.it0 = it0; a.niter = niter;
    void *args[1] = {(void *)&a};
    // cooperative: every tile spins on its neighbours' flags, so all gx·gy CTAs must be resident
    JR_CUDA(cudaLaunchCooperativeKernel((const void *)k_v2_resident, dim3(p->gx * p->gy, 1, 1), dim3(RX, p->rows, 1), args, p->smem, ctx->stream));
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}
