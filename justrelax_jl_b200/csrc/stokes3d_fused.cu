// stokes3d_fused.cu — ONE fused sm_100a kernel per PT iteration of the 3D visco-elastic Stokes solver
// (variant 3D-VA), plus the ping-pong boundary kernel.
//
// What it replaces (reference, one @parallel launch each, ≈75 array passes = 600 B/cell/iteration):
//   compute_∇V!  → compute_P!  → compute_strain_rate! → compute_τ! → compute_V! → velocity2displacement!
//   (src/stokes/Stokes3D.jl:78-119; kernels VelocityKernels.jl:3-6,59-104,182-242, PressureKernels.jl:10-15,
//   StressKernels.jl:149-230) and flow_bcs! (BoundaryConditions.jl:86-99).
//
// Design (B200: HBM-bound FP64 stencil, no tensor-core work):
//   * 2.5D marching: a CTA owns a (TX × TY) column tile of the (x,y) plane plus a one-cell halo ring
//     (BX = TX+2 = 32 lanes → a warp is one x-row: every global access is a coalesced 256-B row segment)
//     and marches a z-chunk.  Each z-plane of every input is read ONCE per CTA; neighbour values are
//     exchanged through shared memory; z-neighbours live in a register queue
//     (V(k), η-pair sums, τzz/P/ρgz/ητ of plane k−1, bottom-edge stresses, partial Rz).
//   * Jacobi-exact: V, P, τ are double-buffered (in → out), so halo cells can be recomputed by
//     neighbouring CTAs without races; the arithmetic per point is the reference's, operation for operation
//     (same fma placement, same summation order), hence results are bit-comparable with the CPU oracle.
//   * Diagnostics nobody reads inside the loop (∇V, RP, ε, R, U) are written only when `diag` is set
//     (nout samples / last iteration): algorithmic traffic is 25 passes = 200 B/cell/iteration.
//   * dt = Inf (SolVi, convection configs): 1/(G dt) = 1/(K dt) = 1/dt = 0 exactly, so G, K, P0, Q, τ_o are
//     not read at all (template FINITE_DT=false).
//   * The ghost layers / boundary faces of V_out are filled by k_bc_pingpong3 (O(surface)).
#include "common.cuh"

struct FusedArgs {
    // state in / out (ping-pong)
    const double *Vx_i, *Vy_i, *Vz_i, *P_i, *txx_i, *tyy_i, *tzz_i, *tyz_i, *txz_i, *txy_i;
    double *Vx_o, *Vy_o, *Vz_o, *P_o, *txx_o, *tyy_o, *tzz_o, *tyz_o, *txz_o, *txy_o;
    // read-only
    const double *eta, *etatau, *fx, *fy, *fz;
    const double *G, *K, *P0, *Q, *oxx, *oyy, *ozz, *oyz, *oxz, *oxy;  // τ_o (finite dt only)
    // diagnostics (written when diag != 0)
    double *divV, *RP, *exx, *eyy, *ezz, *eyz, *exz, *exy, *Rx, *Ry, *Rz, *Ux, *Uy, *Uz;
    int nx, ny, nz, kchunk, diag;
    double _dx, _dy, _dz, dt, r, theta_dtau, eta_dtau;
};

#define BX 32

template <int BY, bool FINITE_DT>
__global__ void __launch_bounds__(BX *BY) k_stokes3d_va_fused(const FusedArgs a)
{
    constexpr int TX = BX - 2, TY = BY - 2;
    // shared planes: arrival (Vx,Vy,Vz,η,G) and compute (τxx,τyy,P,τxy,τxz,τyz,fx,fy,ητ)
    __shared__ double sVx[BY][BX], sVy[BY][BX], sVz[BY][BX], sEta[BY][BX], sG[FINITE_DT ? BY : 1][BX];
    __shared__ double sTxx[BY][BX], sTyy[BY][BX], sP[BY][BX], sTxy[BY][BX], sTxz[BY][BX], sTyz[BY][BX];
    __shared__ double sFx[BY][BX], sFy[BY][BX], sEtt[BY][BX];

    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gi = blockIdx.x * TX + tx - 1, gj = blockIdx.y * TY + ty - 1;
    const int kb = blockIdx.z * a.kchunk;
    const int ke = min(kb + a.kchunk, nz);
    const int txW = max(tx - 1, 0), txE = min(tx + 1, BX - 1), tyS = max(ty - 1, 0), tyN = min(ty + 1, BY - 1);

    const bool own = tx >= 1 && tx <= BX - 2 && ty >= 1 && ty <= BY - 2 && gi <= nx && gj <= ny;
    const bool inx = gi >= -1 && gi <= nx, iny = gj >= -1 && gj <= ny;
    const bool ldVx = gi >= 0 && gi <= nx && iny;
    const bool ldVy = inx && gj >= 0 && gj <= ny;
    const bool ldVz = inx && iny;
    const bool cell = gi >= 0 && gi < nx && gj >= 0 && gj < ny;
    const bool vxy = gi >= 0 && gi <= nx && gj >= 0 && gj <= ny;  // xy edge exists
    const bool vxz = gi >= 0 && gi <= nx && gj >= 0 && gj < ny;   // xz edge exists
    const bool vyz = gi >= 0 && gi < nx && gj >= 0 && gj <= ny;   // yz edge exists
    const int ci = jr_clamp(gi, 0, nx - 1), cj = jr_clamp(gj, 0, ny - 1);

    // strides / offsets (0-based, x fastest)
    const size_t sVxk = (size_t)(nx + 1) * (ny + 2), sVyk = (size_t)(nx + 2) * (ny + 1), sVzk = (size_t)(nx + 2) * (ny + 2);
    const size_t sCk = (size_t)nx * ny, sXYk = (size_t)(nx + 1) * (ny + 1), sXZk = (size_t)(nx + 1) * ny, sYZk = (size_t)nx * (ny + 1);
    const size_t oVx = (size_t)(gj + 1) * (nx + 1) + gi;       // + kg * sVxk
    const size_t oVy = (size_t)gj * (nx + 2) + (gi + 1);       // + kg * sVyk
    const size_t oVz = (size_t)(gj + 1) * (nx + 2) + (gi + 1); // + kz * sVzk
    const size_t oC = (size_t)gj * nx + gi;                    // + k * sCk
    const size_t oCc = (size_t)cj * nx + ci;                   // clamped cell
    const size_t oXY = (size_t)gj * (nx + 1) + gi;             // + k * sXYk
    const size_t oXZ = (size_t)gj * (nx + 1) + gi;             // + kz * sXZk
    const size_t oYZ = (size_t)gj * nx + gi;                   // + kz * sYZk

    const double _dx = a._dx, _dy = a._dy, _dz = a._dz, dt = a.dt, th = a.theta_dtau;
    const double inv3 = jr_inv(3.0);
    const double dtr_inf = jr_inv(th + 1.0);  // compute_dτ_r with 1/(G dt) = 0: fma(η, 0, 1) = 1

    // ---- register queue (state of plane k carried along z) ----
    double vx0 = 0, vy0 = 0, vz0 = 0;          // V(k) own
    double dxx0 = 0, dyy0 = 0, exy0 = 0;       // ∂xVx, ∂yVy, ε_xy of plane k
    double eta0 = 1, etaxy0 = 1, sxz0 = 0, syz0 = 0;  // η(k), η̄xy(k), η pair sums of plane k
    double g0 = 1, gxy0 = 1, gsxz0 = 0, gsyz0 = 0;    // same for G (finite dt)
    double tzz_p = 0, P_p = 0, fz_p = 0, ett_p = 1;   // plane k−1 (new τzz, new P, ρgz, ητ)
    double txz_b = 0, tyz_b = 0, sRz = 0;             // new bottom-edge stresses (kz = k) and partial Rz

    // arrival of plane A (= V planes A, η plane A): loads → smem → neighbour-derived quantities.
    // Returns through references the "plane A" versions of the queue entries and the top-edge
    // (kz = A) strain rates / averaged viscosities.
    double nvx, nvy, nvz, neta, ng = 1;  // prefetched arrival values
    auto load_arrival = [&](int A) {
        // A ∈ [−1, nz]; V planes use ghosted index A+1; Vz face index A (valid 0..nz); η clamped
        const int kg = A + 1;
        nvx = ldVx ? a.Vx_i[oVx + (size_t)kg * sVxk] : 0.0;
        nvy = ldVy ? a.Vy_i[oVy + (size_t)kg * sVyk] : 0.0;
        nvz = (ldVz && A >= 0) ? a.Vz_i[oVz + (size_t)A * sVzk] : 0.0;
        const int ck = jr_clamp(A, 0, nz - 1);
        neta = a.eta[oCc + (size_t)ck * sCk];
        if (FINITE_DT) ng = a.G[oCc + (size_t)ck * sCk];
    };

    // compute-set prefetch for step k
    double o_txx, o_tyy, o_tzz, o_txy, o_txz, o_tyz, o_P, fx, fy, fz, ett;
    double q_txx = 0, q_tyy = 0, q_tzz = 0, q_txy = 0, q_txz = 0, q_tyz = 0, q_P0 = 0, q_K = 1, q_Q = 0;  // finite dt
    auto load_compute = [&](int k) {
        const bool ck = cell && k >= 0 && k < nz;
        const size_t c = oC + (size_t)max(k, 0) * sCk;
        o_txx = ck ? a.txx_i[c] : 0.0;
        o_tyy = ck ? a.tyy_i[c] : 0.0;
        o_tzz = ck ? a.tzz_i[c] : 0.0;
        o_P = ck ? a.P_i[c] : 0.0;
        fx = ck ? a.fx[c] : 0.0;
        fy = ck ? a.fy[c] : 0.0;
        fz = ck ? a.fz[c] : 0.0;
        ett = ck ? a.etatau[c] : 1.0;
        const bool kxy = vxy && k >= 0 && k < nz;
        o_txy = kxy ? a.txy_i[oXY + (size_t)max(k, 0) * sXYk] : 0.0;
        o_txz = vxz ? a.txz_i[oXZ + (size_t)(k + 1) * sXZk] : 0.0;  // kz = k+1 ∈ [0, nz]
        o_tyz = vyz ? a.tyz_i[oYZ + (size_t)(k + 1) * sYZk] : 0.0;
        if (FINITE_DT) {
            q_txx = ck ? a.oxx[c] : 0.0;
            q_tyy = ck ? a.oyy[c] : 0.0;
            q_tzz = ck ? a.ozz[c] : 0.0;
            q_P0 = ck ? a.P0[c] : 0.0;
            q_K = ck ? a.K[c] : 1.0;
            q_Q = ck ? a.Q[c] : 0.0;
            q_txy = kxy ? a.oxy[oXY + (size_t)max(k, 0) * sXYk] : 0.0;
            q_txz = vxz ? a.oxz[oXZ + (size_t)(k + 1) * sXZk] : 0.0;
            q_tyz = vyz ? a.oyz[oYZ + (size_t)(k + 1) * sYZk] : 0.0;
        }
    };

    // -------- prologue: arrival of plane kb−1 fills the queue --------
    load_arrival(kb - 1);
    sVx[ty][tx] = nvx; sVy[ty][tx] = nvy; sVz[ty][tx] = nvz; sEta[ty][tx] = neta;
    if (FINITE_DT) sG[ty][tx] = ng;
    __syncthreads();
    {
        vx0 = nvx; vy0 = nvy; vz0 = nvz; eta0 = neta;
        dxx0 = (-nvx + sVx[ty][txE]) * _dx;
        dyy0 = (-nvy + sVy[tyN][tx]) * _dy;
        exy0 = 0.5 * (_dy * (nvx - sVx[tyS][tx]) + _dx * (nvy - sVy[ty][txW]));
        const double eW = sEta[ty][txW], eS = sEta[tyS][tx], eSW = sEta[tyS][txW];
        etaxy0 = 0.25 * (eSW + eS + eW + neta);
        sxz0 = eW + neta;
        syz0 = eS + neta;
        if (FINITE_DT) {
            g0 = ng;
            const double gW = sG[ty][txW], gS = sG[tyS][tx], gSW = sG[tyS][txW];
            gxy0 = 0.25 * (gSW + gS + gW + ng);
            gsxz0 = gW + ng;
            gsyz0 = gS + ng;
        }
    }
    load_arrival(kb);       // arrival set of the first step (plane kb)
    load_compute(kb - 1);   // compute set of the warm-up step
    __syncthreads();

    for (int k = kb - 1; k < ke; ++k) {
        // ---- S1: publish arrival plane A = k+1 ----
        const double vx1 = nvx, vy1 = nvy, vz1 = nvz, eta1 = neta, g1 = ng;
        sVx[ty][tx] = vx1; sVy[ty][tx] = vy1; sVz[ty][tx] = vz1; sEta[ty][tx] = eta1;
        if (FINITE_DT) sG[ty][tx] = g1;
        // move this step's compute set out of the prefetch registers, then prefetch step k+1
        const double c_txx = o_txx, c_tyy = o_tyy, c_tzz = o_tzz, c_txy = o_txy, c_txz = o_txz, c_tyz = o_tyz, c_P = o_P;
        const double c_fx = fx, c_fy = fy, c_fz = fz, c_ett = ett;
        const double c_qxx = q_txx, c_qyy = q_tyy, c_qzz = q_tzz, c_qxy = q_txy, c_qxz = q_txz, c_qyz = q_tyz;
        const double c_P0 = q_P0, c_K = q_K, c_Q = q_Q;
        if (k + 1 < ke) {
            load_arrival(k + 2);
            load_compute(k + 1);
        }
        __syncthreads();

        // ---- R1: neighbour-derived quantities of plane A, top-edge strain rates, all new stresses ----
        const double dxx1 = (-vx1 + sVx[ty][txE]) * _dx;
        const double dyy1 = (-vy1 + sVy[tyN][tx]) * _dy;
        const double exy1 = 0.5 * (_dy * (vx1 - sVx[tyS][tx]) + _dx * (vy1 - sVy[ty][txW]));
        const double exz_t = 0.5 * (_dz * (vx1 - vx0) + _dx * (vz1 - sVz[ty][txW]));
        const double eyz_t = 0.5 * (_dz * (vy1 - vy0) + _dy * (vz1 - sVz[tyS][tx]));
        const double eW = sEta[ty][txW], eS = sEta[tyS][tx], eSW = sEta[tyS][txW];
        const double etaxy1 = 0.25 * (eSW + eS + eW + eta1);
        const double etaxz_t = 0.25 * (sxz0 + eW + eta1);
        const double etayz_t = 0.25 * (syz0 + eS + eta1);
        const double sxz1 = eW + eta1, syz1 = eS + eta1;
        double gxy1 = 1, gxz_t = 1, gyz_t = 1, gsxz1 = 0, gsyz1 = 0;
        if (FINITE_DT) {
            const double gW = sG[ty][txW], gS = sG[tyS][tx], gSW = sG[tyS][txW];
            gxy1 = 0.25 * (gSW + gS + gW + g1);
            gxz_t = 0.25 * (gsxz0 + gW + g1);
            gyz_t = 0.25 * (gsyz0 + gS + g1);
            gsxz1 = gW + g1;
            gsyz1 = gS + g1;
        }

        // centre of plane k
        const double dzz = (-vz0 + vz1) * _dz;
        const double divV = dxx0 + dyy0 + dzz;
        double RP, P_n = c_P;
        if (FINITE_DT) {
            jr_compute_P_point(RP, P_n, c_P0, divV, c_Q, eta0, c_K, g0, dt, a.r, th);
        } else {
            // _Kdt = _Gdt = _dt = 0 exactly (PressureKernels.jl:186-195 with dt = Inf)
            RP = -divV;
            const double psi = jr_inv(jr_inv(eta0)) * a.r / th;
            P_n = (-divV) * psi + c_P;
        }
        const double d3 = divV * inv3;
        const double exx = dxx0 - d3, eyy = dyy0 - d3, ezz = dzz - d3;
        double txx_n, tyy_n, tzz_n, txy_n, txz_n, tyz_n;
        if (FINITE_DT) {
            {
                const double _Gdt = jr_inv(g0 * dt), dtr = jr_dtau_r(th, eta0, _Gdt);
                txx_n = c_txx + jr_stress_increment(c_txx, c_qxx, eta0, exx, _Gdt, dtr);
                tyy_n = c_tyy + jr_stress_increment(c_tyy, c_qyy, eta0, eyy, _Gdt, dtr);
                tzz_n = c_tzz + jr_stress_increment(c_tzz, c_qzz, eta0, ezz, _Gdt, dtr);
            }
            {
                const double _Gdt = jr_inv(gxy0 * dt), dtr = jr_dtau_r(th, etaxy0, _Gdt);
                txy_n = c_txy + jr_stress_increment(c_txy, c_qxy, etaxy0, exy0, _Gdt, dtr);
            }
            {
                const double _Gdt = jr_inv(gxz_t * dt), dtr = jr_dtau_r(th, etaxz_t, _Gdt);
                txz_n = c_txz + jr_stress_increment(c_txz, c_qxz, etaxz_t, exz_t, _Gdt, dtr);
            }
            {
                const double _Gdt = jr_inv(gyz_t * dt), dtr = jr_dtau_r(th, etayz_t, _Gdt);
                tyz_n = c_tyz + jr_stress_increment(c_tyz, c_qyz, etayz_t, eyz_t, _Gdt, dtr);
            }
        } else {
            txx_n = c_txx + dtr_inf * fma(2.0 * eta0, exx, -c_txx);
            tyy_n = c_tyy + dtr_inf * fma(2.0 * eta0, eyy, -c_tyy);
            tzz_n = c_tzz + dtr_inf * fma(2.0 * eta0, ezz, -c_tzz);
            txy_n = c_txy + dtr_inf * fma(2.0 * etaxy0, exy0, -c_txy);
            txz_n = c_txz + dtr_inf * fma(2.0 * etaxz_t, exz_t, -c_txz);
            tyz_n = c_tyz + dtr_inf * fma(2.0 * etayz_t, eyz_t, -c_tyz);
        }

        // ---- S2: publish new stresses / pressure and the cell-centred forcing of plane k ----
        sTxx[ty][tx] = txx_n; sTyy[ty][tx] = tyy_n; sP[ty][tx] = P_n; sTxy[ty][tx] = txy_n;
        sTxz[ty][tx] = txz_n; sTyz[ty][tx] = tyz_n; sFx[ty][tx] = c_fx; sFy[ty][tx] = c_fy; sEtt[ty][tx] = c_ett;
        __syncthreads();

        // ---- R2: momentum residuals and velocity update of plane k, partial Rz of face k+1 ----
        const double sRz_next = _dx * (sTxz[ty][txE] - txz_n) + _dy * (sTyz[tyN][tx] - tyz_n);
        const bool kin = k >= kb;  // this chunk owns plane k (k = kb−1 is the warm-up plane)
        if (own && kin) {
            const size_t c = oC + (size_t)k * sCk;
            if (cell) {
                a.P_o[c] = P_n; a.txx_o[c] = txx_n; a.tyy_o[c] = tyy_n; a.tzz_o[c] = tzz_n;
                if (a.diag) {
                    a.divV[c] = divV; a.RP[c] = RP; a.exx[c] = exx; a.eyy[c] = eyy; a.ezz[c] = ezz;
                }
            }
            if (vxy) {
                a.txy_o[oXY + (size_t)k * sXYk] = txy_n;
                if (a.diag) a.exy[oXY + (size_t)k * sXYk] = exy0;
            }
            // x-momentum: face gi between cells gi−1 (W) and gi
            if (cell && gi >= 1) {
                const double R = (-sTxx[ty][txW] + txx_n) * _dx + _dy * (sTxy[tyN][tx] - txy_n) + _dz * (txz_n - txz_b) -
                                 (-sP[ty][txW] + P_n) * _dx - 0.5 * (sFx[ty][txW] + c_fx);
                const double vn = vx0 + R * a.eta_dtau / (0.5 * (sEtt[ty][txW] + c_ett));
                a.Vx_o[oVx + (size_t)(k + 1) * sVxk] = vn;
                if (a.diag) {
                    a.Rx[((size_t)k * ny + gj) * (nx - 1) + (gi - 1)] = R;
                    a.Ux[oVx + (size_t)(k + 1) * sVxk] = vn * dt;
                }
            }
            // y-momentum: face gj between cells gj−1 (S) and gj
            if (cell && gj >= 1) {
                const double R = _dx * (sTxy[ty][txE] - txy_n) + _dy * (tyy_n - sTyy[tyS][tx]) + _dz * (tyz_n - tyz_b) -
                                 (-sP[tyS][tx] + P_n) * _dy - 0.5 * (sFy[tyS][tx] + c_fy);
                const double vn = vy0 + R * a.eta_dtau / (0.5 * (sEtt[tyS][tx] + c_ett));
                a.Vy_o[oVy + (size_t)(k + 1) * sVyk] = vn;
                if (a.diag) {
                    a.Ry[((size_t)k * (ny - 1) + (gj - 1)) * nx + gi] = R;
                    a.Uy[oVy + (size_t)(k + 1) * sVyk] = vn * dt;
                }
            }
            // z-momentum: face k between planes k−1 and k
            if (cell && k >= 1) {
                const double R = sRz + (-tzz_p + tzz_n) * _dz - (-P_p + P_n) * _dz - 0.5 * (fz_p + c_fz);
                const double vn = vz0 + R * a.eta_dtau / (0.5 * (ett_p + c_ett));
                a.Vz_o[oVz + (size_t)k * sVzk] = vn;
                if (a.diag) {
                    a.Rz[((size_t)(k - 1) * ny + gj) * nx + gi] = R;
                    a.Uz[oVz + (size_t)k * sVzk] = vn * dt;
                }
            }
        }
        // top edges (kz = k+1) belong to the chunk that owns plane k; the kz = 0 edges to chunk 0's warm-up
        if (own && (kin || kb == 0)) {
            if (vxz) {
                a.txz_o[oXZ + (size_t)(k + 1) * sXZk] = txz_n;
                if (a.diag) a.exz[oXZ + (size_t)(k + 1) * sXZk] = exz_t;
            }
            if (vyz) {
                a.tyz_o[oYZ + (size_t)(k + 1) * sYZk] = tyz_n;
                if (a.diag) a.eyz[oYZ + (size_t)(k + 1) * sYZk] = eyz_t;
            }
        }

        // ---- rotate the register queue: plane k+1 becomes plane k ----
        vx0 = vx1; vy0 = vy1; vz0 = vz1; eta0 = eta1;
        dxx0 = dxx1; dyy0 = dyy1; exy0 = exy1; etaxy0 = etaxy1; sxz0 = sxz1; syz0 = syz1;
        if (FINITE_DT) { g0 = g1; gxy0 = gxy1; gsxz0 = gsxz1; gsyz0 = gsyz1; }
        tzz_p = tzz_n; P_p = P_n; fz_p = c_fz; ett_p = c_ett;
        txz_b = txz_n; tyz_b = tyz_n; sRz = sRz_next;
    }
}

// ------------------------------------------------------------------------------------------------------
// Ping-pong boundary kernel: fills every element of V_out the fused kernel does not compute, i.e. the
// tangential ghost layers and the boundary-normal faces, from (a) the freshly computed interior of V_out
// and (b) V_in for layers no boundary condition touches (prescribed values, or halo planes that the
// exchange overwrites afterwards).  Semantics = no_slip! → free_slip! of the reference applied as complete
// sweeps (no_slip.jl:21-54, free_slip.jl:15-70 incl. quirk Q2); every ghost value is gathered from its
// fully-clamped source with the product of the per-dimension signs, so ghost edges/corners are
// deterministic.  With diag set it also writes U = V·dt for those elements from V_in — the reference
// takes U before flow_bcs! (Stokes3D.jl:118-119).
struct BcArr {
    const double *in;
    double *out;
    double *U;
    int n[3];
    int normal;  // normal dimension of this component
};
struct BcArgs {
    BcArr A[3];
    int lo_fs[3], hi_fs[3], lo_ns[3], hi_ns[3];  // per dimension and side: free-slip / no-slip active
    int diag;
    double dt;
};

__global__ void k_bc_pingpong3(const BcArgs b)
{
    const int which = blockIdx.z / 6, plane = blockIdx.z % 6;  // component, (dim, lo/hi)
    const BcArr &A = b.A[which];
    const int d = plane >> 1, hi = plane & 1;
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    const int p = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y * blockDim.y + threadIdx.y;
    // fastest-varying free coordinate on threadIdx.x: for d = 0 planes use (dim1, dim2), else dim0 first
    int c[3];
    int u = (d == 0) ? 1 : 0, v = (d == 2) ? 1 : 2;
    if (p >= A.n[u] || q >= A.n[v]) return;
    (void)d1; (void)d2;
    c[d] = hi ? A.n[d] - 1 : 0;
    c[u] = p;
    c[v] = q;
    int s[3] = {c[0], c[1], c[2]};
    double sign = 1.0;
    bool zero = false;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const bool lo_e = c[e] == 0, hi_e = c[e] == A.n[e] - 1;
        if (!lo_e && !hi_e) continue;
        // no_slip! runs before free_slip! (BoundaryConditions.jl:86-99): on a side that carries both
        // (possible through quirk Q2) the free-slip copy wins for the tangential ghosts, the normal face stays 0
        const bool fsl = lo_e ? b.lo_fs[e] : b.hi_fs[e], nsl = lo_e ? b.lo_ns[e] : b.hi_ns[e];
        if (e == A.normal) {
            if (nsl) zero = true;
        } else if (fsl || nsl) {
            s[e] = lo_e ? 1 : A.n[e] - 2;
            if (!fsl) sign = -sign;
        }
    }
    const size_t ic = ((size_t)c[2] * A.n[1] + c[1]) * A.n[0] + c[0];
    const size_t is = ((size_t)s[2] * A.n[1] + s[1]) * A.n[0] + s[0];
    bool computed = true;
#pragma unroll
    for (int e = 0; e < 3; e++) computed = computed && s[e] >= 1 && s[e] <= A.n[e] - 2;
    double val;
    if (zero) val = 0.0;
    else val = sign * (computed ? A.out[is] : A.in[is]);
    A.out[ic] = val;
    if (b.diag) A.U[ic] = A.in[ic] * b.dt;
}

// ------------------------------------------------------------------------------------------------------
#define F(name) (s->f[JR_F_##name])

int jr_stokes3d_VA_fused_supported(const jr_fields *s, const jr_stokes_opts *o)
{
    for (int q = 0; q < 6; q++)
        if (o->periodic[q]) return JR_ERR_UNSUPPORTED;  // periodic wrap: reference-structured path
    if (s->n[0] < 3 || s->n[1] < 3 || s->n[2] < 3) return JR_ERR_UNSUPPORTED;
    return JR_OK;
}

static const char *k_pp_names[10] = {"pp_Vx", "pp_Vy", "pp_Vz", "pp_P", "pp_txx", "pp_tyy", "pp_tzz", "pp_tyz", "pp_txz", "pp_txy"};

// `parity` 0: user arrays → scratch, 1: scratch → user arrays.  The driver copies the scratch set back
// when a solve ends on an odd number of iterations (jr_stokes3d_VA_fused_finish).
int jr_stokes3d_VA_fused_iter(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int diag, int parity)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const size_t sz[10] = {(size_t)(nx + 1) * (ny + 2) * (nz + 2), (size_t)(nx + 2) * (ny + 1) * (nz + 2),
                           (size_t)(nx + 2) * (ny + 2) * (nz + 1), (size_t)nx * ny * nz, (size_t)nx * ny * nz,
                           (size_t)nx * ny * nz, (size_t)nx * ny * nz, (size_t)nx * (ny + 1) * (nz + 1),
                           (size_t)(nx + 1) * ny * (nz + 1), (size_t)(nx + 1) * (ny + 1) * nz};
    double *user[10] = {F(Vx), F(Vy), F(Vz), F(P), F(txx), F(tyy), F(tzz), F(tyz), F(txz), F(txy)};
    double *scr[10];
    for (int q = 0; q < 10; q++) {
        void *p = nullptr;
        int st = jr_ctx_scratch(ctx, k_pp_names[q], sz[q] * sizeof(double), &p);
        if (st) return st;
        scr[q] = (double *)p;
    }
    double **in = parity ? scr : user, **out = parity ? user : scr;

    FusedArgs a;
    a.Vx_i = in[0]; a.Vy_i = in[1]; a.Vz_i = in[2]; a.P_i = in[3]; a.txx_i = in[4]; a.tyy_i = in[5]; a.tzz_i = in[6];
    a.tyz_i = in[7]; a.txz_i = in[8]; a.txy_i = in[9];
    a.Vx_o = out[0]; a.Vy_o = out[1]; a.Vz_o = out[2]; a.P_o = out[3]; a.txx_o = out[4]; a.tyy_o = out[5]; a.tzz_o = out[6];
    a.tyz_o = out[7]; a.txz_o = out[8]; a.txy_o = out[9];
    a.eta = F(eta); a.etatau = F(etatau); a.fx = F(rhogx); a.fy = F(rhogy); a.fz = F(rhogz);
    a.G = F(G); a.K = F(K); a.P0 = F(P0); a.Q = F(Q);
    a.oxx = F(txx_o); a.oyy = F(tyy_o); a.ozz = F(tzz_o); a.oyz = F(tyz_o); a.oxz = F(txz_o); a.oxy = F(txy_o);
    a.divV = F(divV); a.RP = F(RP); a.exx = F(exx); a.eyy = F(eyy); a.ezz = F(ezz); a.eyz = F(eyz); a.exz = F(exz); a.exy = F(exy);
    a.Rx = F(Rx); a.Ry = F(Ry); a.Rz = F(Rz); a.Ux = F(Ux); a.Uy = F(Uy); a.Uz = F(Uz);
    a.nx = nx; a.ny = ny; a.nz = nz; a.diag = diag;
    a._dx = o->_di[0]; a._dy = o->_di[1]; a._dz = o->_di[2]; a.dt = o->dt; a.r = o->r; a.theta_dtau = o->theta_dtau;
    a.eta_dtau = o->eta_dtau;

    constexpr int BYc = 8;
    const int TX = BX - 2, TY = BYc - 2;
    const int gx = (nx + 1 + TX - 1) / TX, gy = (ny + 1 + TY - 1) / TY;
    // z-chunks: enough CTAs for >= 4 waves of resident CTAs, warm-up plane overhead <= ~3 %
    int nchunk = 1;
    const long want = (long)ctx->sm_count * 8;
    while ((long)gx * gy * nchunk < want && nz / (nchunk + 1) >= 32) nchunk++;
    a.kchunk = (nz + nchunk - 1) / nchunk;
    dim3 grid(gx, gy, (nz + a.kchunk - 1) / a.kchunk), block(BX, BYc, 1);
    const bool finite_dt = std::isfinite(o->dt);
    if (finite_dt) k_stokes3d_va_fused<BYc, true><<<grid, block, 0, ctx->stream>>>(a);
    else k_stokes3d_va_fused<BYc, false><<<grid, block, 0, ctx->stream>>>(a);

    BcArgs b;
    b.A[0] = BcArr{in[0], out[0], F(Ux), {nx + 1, ny + 2, nz + 2}, 0};
    b.A[1] = BcArr{in[1], out[1], F(Uy), {nx + 2, ny + 1, nz + 2}, 1};
    b.A[2] = BcArr{in[2], out[2], F(Uz), {nx + 2, ny + 2, nz + 1}, 2};
    // flags: left,right,front,back,top,bot.  no_slip: bot → z lo, top → z hi; free_slip (Q2): top → z lo, bot → z hi
    const int32_t *fs = o->free_slip, *ns = o->no_slip;
    b.lo_fs[0] = fs[0]; b.hi_fs[0] = fs[1]; b.lo_ns[0] = ns[0]; b.hi_ns[0] = ns[1];
    b.lo_fs[1] = fs[2]; b.hi_fs[1] = fs[3]; b.lo_ns[1] = ns[2]; b.hi_ns[1] = ns[3];
    b.lo_fs[2] = fs[4]; b.hi_fs[2] = fs[5]; b.lo_ns[2] = ns[5]; b.hi_ns[2] = ns[4];
    b.diag = diag; b.dt = o->dt;
    int m = nx > ny ? nx : ny;
    m = (m > nz ? m : nz) + 2;
    dim3 bgrid((m + 31) / 32, (m + 7) / 8, 18), bblock(32, 8, 1);
    k_bc_pingpong3<<<bgrid, bblock, 0, ctx->stream>>>(b);
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

// copy the scratch state back into the user's arrays (after an odd number of fused iterations)
int jr_stokes3d_VA_fused_finish(jr_context *ctx, const jr_fields *s)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const size_t sz[10] = {(size_t)(nx + 1) * (ny + 2) * (nz + 2), (size_t)(nx + 2) * (ny + 1) * (nz + 2),
                           (size_t)(nx + 2) * (ny + 2) * (nz + 1), (size_t)nx * ny * nz, (size_t)nx * ny * nz,
                           (size_t)nx * ny * nz, (size_t)nx * ny * nz, (size_t)nx * (ny + 1) * (nz + 1),
                           (size_t)(nx + 1) * ny * (nz + 1), (size_t)(nx + 1) * (ny + 1) * nz};
    double *user[10] = {F(Vx), F(Vy), F(Vz), F(P), F(txx), F(tyy), F(tzz), F(tyz), F(txz), F(txy)};
    for (int q = 0; q < 10; q++) {
        void *p = nullptr;
        int st = jr_ctx_scratch(ctx, k_pp_names[q], sz[q] * sizeof(double), &p);
        if (st) return st;
        JR_CUDA(cudaMemcpyAsync(user[q], p, sz[q] * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return JR_OK;
}
