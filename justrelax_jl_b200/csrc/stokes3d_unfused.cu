// stokes3d_unfused.cu — reference-structured GPU path for the 3D-VA PT iteration (one kernel per
// reference @parallel launch, same array traffic ≈ 600 B/cell/iter).  It exists (a) as the first
// parity-checked CUDA restatement, (b) as the on-device "what the reference's kernel split costs on a
// B200" comparison reported by bench.py next to the fused kernel.  Selected with JR_FLAG_UNFUSED.
//
// Reference: src/stokes/Stokes3D.jl:78-121; kernels cited per function.  Indices are 1-based (IX3).
#include "common.cuh"
#include "comm.cuh"

#define F(name) (s.f[JR_F_##name])

// compute_∇V!  VelocityKernels.jl:3-6 ; MiniKernels.jl:53-55,104-105
__global__ void k_divV3(jr_fields s, double _dx, double _dy, double _dz)
{
    const int nx = s.n[0], ny = s.n[1], nz = s.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const double *Vx = F(Vx), *Vy = F(Vy), *Vz = F(Vz);
    const double dx = (-Vx[IX3(nx + 1, ny + 2, i, j + 1, k + 1)] + Vx[IX3(nx + 1, ny + 2, i + 1, j + 1, k + 1)]) * _dx;
    const double dy = (-Vy[IX3(nx + 2, ny + 1, i + 1, j, k + 1)] + Vy[IX3(nx + 2, ny + 1, i + 1, j + 1, k + 1)]) * _dy;
    const double dz = (-Vz[IX3(nx + 2, ny + 2, i + 1, j + 1, k)] + Vz[IX3(nx + 2, ny + 2, i + 1, j + 1, k + 1)]) * _dz;
    F(divV)[IX3(nx, ny, i, j, k)] = dx + dy + dz;
}

// compute_P! (compressible, K/G arrays)  PressureKernels.jl:10-15,186-195
__global__ void k_P_VA(jr_fields s, const double *__restrict__ eta, double dt, double r, double theta_dtau, size_t N)
{
    const size_t I = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (I >= N) return;
    double RP, P = F(P)[I];
    jr_compute_P_point(RP, P, F(P0)[I], F(divV)[I], F(Q)[I], eta[I], F(K)[I], F(G)[I], dt, r, theta_dtau);
    F(RP)[I] = RP;
    F(P)[I] = P;
}

// compute_strain_rate! 3D  VelocityKernels.jl:59-104
__global__ void k_strain3(jr_fields s, double _dx, double _dy, double _dz, int p)
{
    const int nx = s.n[0], ny = s.n[1], nz = s.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx + p || j > ny + p || k > nz + p) return;
    const double *Vx = F(Vx), *Vy = F(Vy), *Vz = F(Vz);
#define VX(i, j, k) Vx[IX3(nx + 1, ny + 2, i, j, k)]
#define VY(i, j, k) Vy[IX3(nx + 2, ny + 1, i, j, k)]
#define VZ(i, j, k) Vz[IX3(nx + 2, ny + 2, i, j, k)]
    if (i <= nx && j <= ny && k <= nz) {
        const double d3 = F(divV)[IX3(nx, ny, i, j, k)] * jr_inv(3.0);
        F(exx)[IX3(nx, ny, i, j, k)] = (-VX(i, j + 1, k + 1) + VX(i + 1, j + 1, k + 1)) * _dx - d3;
        F(eyy)[IX3(nx, ny, i, j, k)] = (-VY(i + 1, j, k + 1) + VY(i + 1, j + 1, k + 1)) * _dy - d3;
        F(ezz)[IX3(nx, ny, i, j, k)] = (-VZ(i + 1, j + 1, k) + VZ(i + 1, j + 1, k + 1)) * _dz - d3;
    }
    if (i <= nx && j <= ny + 1 && k <= nz + 1)
        F(eyz)[IX3(nx, ny + 1, i, j, k)] =
            0.5 * (_dz * (VY(i + 1, j, k + 1) - VY(i + 1, j, k)) + _dy * (VZ(i + 1, j + 1, k) - VZ(i + 1, j, k)));
    if (i <= nx + 1 && j <= ny && k <= nz + 1)
        F(exz)[IX3(nx + 1, ny, i, j, k)] =
            0.5 * (_dz * (VX(i, j + 1, k + 1) - VX(i, j + 1, k)) + _dx * (VZ(i + 1, j + 1, k) - VZ(i, j + 1, k)));
    if (i <= nx + 1 && j <= ny + 1 && k <= nz)
        F(exy)[IX3(nx + 1, ny + 1, i, j, k)] =
            0.5 * (_dy * (VX(i, j + 1, k + 1) - VX(i, j, k + 1)) + _dx * (VY(i + 1, j, k + 1) - VY(i, j, k + 1)));
#undef VX
#undef VY
#undef VZ
}

// compute_τ! 3D visco-elastic  StressKernels.jl:149-230 ; MiniKernels.jl:133-147
__global__ void k_tau3_VE(jr_fields s, double dt, double theta_dtau)
{
    const int nx = s.n[0], ny = s.n[1], nz = s.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx + 1 || j > ny + 1 || k > nz + 1) return;
    const double *eta = F(eta), *G = F(G);
    const int i0 = jr_clamp(i - 1, 1, nx), i1 = jr_clamp(i, 1, nx);
    const int j0 = jr_clamp(j - 1, 1, ny), j1 = jr_clamp(j, 1, ny);
    const int k0 = jr_clamp(k - 1, 1, nz), k1 = jr_clamp(k, 1, nz);
#define C(A, i, j, k) A[IX3(nx, ny, i, j, k)]
    if (i <= nx && j <= ny && k <= nz) {
        const size_t I = IX3(nx, ny, i, j, k);
        const double _Gdt = jr_inv(G[I] * dt), e = eta[I];
        const double dtr = jr_dtau_r(theta_dtau, e, _Gdt);
        F(txx)[I] += jr_stress_increment(F(txx)[I], F(txx_o)[I], e, F(exx)[I], _Gdt, dtr);
        F(tyy)[I] += jr_stress_increment(F(tyy)[I], F(tyy_o)[I], e, F(eyy)[I], _Gdt, dtr);
        F(tzz)[I] += jr_stress_increment(F(tzz)[I], F(tzz_o)[I], e, F(ezz)[I], _Gdt, dtr);
    }
    if (i <= nx + 1 && j <= ny + 1 && k <= nz) {
        const size_t I = IX3(nx + 1, ny + 1, i, j, k);
        const double e = 0.25 * (C(eta, i0, j0, k) + C(eta, i1, j0, k) + C(eta, i0, j1, k) + C(eta, i1, j1, k));
        const double g = 0.25 * (C(G, i0, j0, k) + C(G, i1, j0, k) + C(G, i0, j1, k) + C(G, i1, j1, k));
        const double _Gdt = jr_inv(g * dt), dtr = jr_dtau_r(theta_dtau, e, _Gdt);
        F(txy)[I] += jr_stress_increment(F(txy)[I], F(txy_o)[I], e, F(exy)[I], _Gdt, dtr);
    }
    if (i <= nx + 1 && j <= ny && k <= nz + 1) {
        const size_t I = IX3(nx + 1, ny, i, j, k);
        const double e = 0.25 * (C(eta, i0, j, k0) + C(eta, i1, j, k0) + C(eta, i0, j, k1) + C(eta, i1, j, k1));
        const double g = 0.25 * (C(G, i0, j, k0) + C(G, i1, j, k0) + C(G, i0, j, k1) + C(G, i1, j, k1));
        const double _Gdt = jr_inv(g * dt), dtr = jr_dtau_r(theta_dtau, e, _Gdt);
        F(txz)[I] += jr_stress_increment(F(txz)[I], F(txz_o)[I], e, F(exz)[I], _Gdt, dtr);
    }
    if (i <= nx && j <= ny + 1 && k <= nz + 1) {
        const size_t I = IX3(nx, ny + 1, i, j, k);
        const double e = 0.25 * (C(eta, i, j0, k0) + C(eta, i, j1, k0) + C(eta, i, j0, k1) + C(eta, i, j1, k1));
        const double g = 0.25 * (C(G, i, j0, k0) + C(G, i, j1, k0) + C(G, i, j0, k1) + C(G, i, j1, k1));
        const double _Gdt = jr_inv(g * dt), dtr = jr_dtau_r(theta_dtau, e, _Gdt);
        F(tyz)[I] += jr_stress_increment(F(tyz)[I], F(tyz_o)[I], e, F(eyz)[I], _Gdt, dtr);
    }
#undef C
}

// compute_V! 3D  VelocityKernels.jl:182-242
__global__ void k_V3(jr_fields s, double eta_dtau, double _dx, double _dy, double _dz)
{
    const int nx = s.n[0], ny = s.n[1], nz = s.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const double *P = F(P), *fx = F(rhogx), *fy = F(rhogy), *fz = F(rhogz), *etatau = F(etatau);
    const double *txx = F(txx), *tyy = F(tyy), *tzz = F(tzz), *tyz = F(tyz), *txz = F(txz), *txy = F(txy);
#define C(A, i, j, k) A[IX3(nx, ny, i, j, k)]
#define XY(i, j, k) txy[IX3(nx + 1, ny + 1, i, j, k)]
#define XZ(i, j, k) txz[IX3(nx + 1, ny, i, j, k)]
#define YZ(i, j, k) tyz[IX3(nx, ny + 1, i, j, k)]
    if (i <= nx - 1) {
        const double R = (-C(txx, i, j, k) + C(txx, i + 1, j, k)) * _dx + _dy * (XY(i + 1, j + 1, k) - XY(i + 1, j, k)) +
                         _dz * (XZ(i + 1, j, k + 1) - XZ(i + 1, j, k)) - (-C(P, i, j, k) + C(P, i + 1, j, k)) * _dx -
                         0.5 * (C(fx, i, j, k) + C(fx, i + 1, j, k));
        F(Rx)[IX3(nx - 1, ny, i, j, k)] = R;
        F(Vx)[IX3(nx + 1, ny + 2, i + 1, j + 1, k + 1)] += R * eta_dtau / (0.5 * (C(etatau, i, j, k) + C(etatau, i + 1, j, k)));
    }
    if (j <= ny - 1) {
        const double R = _dx * (XY(i + 1, j + 1, k) - XY(i, j + 1, k)) + _dy * (C(tyy, i, j + 1, k) - C(tyy, i, j, k)) +
                         _dz * (YZ(i, j + 1, k + 1) - YZ(i, j + 1, k)) - (-C(P, i, j, k) + C(P, i, j + 1, k)) * _dy -
                         0.5 * (C(fy, i, j, k) + C(fy, i, j + 1, k));
        F(Ry)[IX3(nx, ny - 1, i, j, k)] = R;
        F(Vy)[IX3(nx + 2, ny + 1, i + 1, j + 1, k + 1)] += R * eta_dtau / (0.5 * (C(etatau, i, j, k) + C(etatau, i, j + 1, k)));
    }
    if (k <= nz - 1) {
        const double R = _dx * (XZ(i + 1, j, k + 1) - XZ(i, j, k + 1)) + _dy * (YZ(i, j + 1, k + 1) - YZ(i, j, k + 1)) +
                         (-C(tzz, i, j, k) + C(tzz, i, j, k + 1)) * _dz - (-C(P, i, j, k) + C(P, i, j, k + 1)) * _dz -
                         0.5 * (C(fz, i, j, k) + C(fz, i, j, k + 1));
        F(Rz)[IX3(nx, ny, i, j, k)] = R;
        F(Vz)[IX3(nx + 2, ny + 2, i + 1, j + 1, k + 1)] += R * eta_dtau / (0.5 * (C(etatau, i, j, k) + C(etatau, i, j, k + 1)));
    }
#undef C
#undef XY
#undef XZ
#undef YZ
}

// velocity2displacement!  types/displacement.jl:7-29 (all three components in one launch)
__global__ void k_v2u3(jr_fields s, double dt, size_t nVx, size_t nVy, size_t nVz)
{
    const size_t tot = nVx + nVy + nVz;
    for (size_t I = blockIdx.x * (size_t)blockDim.x + threadIdx.x; I < tot; I += (size_t)gridDim.x * blockDim.x) {
        if (I < nVx) F(Ux)[I] = F(Vx)[I] * dt;
        else if (I < nVx + nVy) F(Uy)[I - nVx] = F(Vy)[I - nVx] * dt;
        else F(Uz)[I - nVx - nVy] = F(Vz)[I - nVx - nVy] * dt;
    }
}

int jr_stokes3d_VA_unfused_iter(jr_context *ctx, const jr_fields *sp, const jr_stokes_opts *o)
{
    const jr_fields &s = *sp;
    const int nx = s.n[0], ny = s.n[1], nz = s.n[2];
    const size_t N = (size_t)nx * ny * nz;
    dim3 blk(32, 8, 1);
    dim3 g0((nx + 31) / 32, (ny + 7) / 8, nz), g1((nx + 32) / 32, (ny + 8) / 8, nz + 1);
    cudaStream_t st = ctx->stream;
    k_divV3<<<g0, blk, 0, st>>>(s, o->_di[0], o->_di[1], o->_di[2]);
    k_P_VA<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(s, F(eta), o->dt, o->r, o->theta_dtau, N);
    k_strain3<<<g1, blk, 0, st>>>(s, o->_di[0], o->_di[1], o->_di[2], 1);
    k_tau3_VE<<<g1, blk, 0, st>>>(s, o->dt, o->theta_dtau);
    k_V3<<<g0, blk, 0, st>>>(s, o->eta_dtau, o->_di[0], o->_di[1], o->_di[2]);
    const size_t nVx = (size_t)(nx + 1) * (ny + 2) * (nz + 2), nVy = (size_t)(nx + 2) * (ny + 1) * (nz + 2),
                 nVz = (size_t)(nx + 2) * (ny + 2) * (nz + 1);
    k_v2u3<<<ctx->sm_count * 8, 256, 0, st>>>(s, o->dt, nVx, nVy, nVz);
    ctx->launches += 6;
    JR_CHECK_LAUNCH();
    int rc = jr_launch_flow_bcs3d(ctx, F(Vx), F(Vy), F(Vz), s.n, o->free_slip, o->no_slip, o->periodic);
    if (rc) return rc;
    // update_halo!(Vx, Vy, Vz)  Stokes3D.jl:120
    const int32_t eVx[3] = {nx + 1, ny + 2, nz + 2}, eVy[3] = {nx + 2, ny + 1, nz + 2}, eVz[3] = {nx + 2, ny + 2, nz + 1};
    const jr_harr H[3] = {jr_harr_dense(F(Vx), eVx, s.n), jr_harr_dense(F(Vy), eVy, s.n), jr_harr_dense(F(Vz), eVz, s.n)};
    return jr_comm_halo(ctx, H, 3);
}
