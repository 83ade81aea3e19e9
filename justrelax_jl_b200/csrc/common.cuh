// common.cuh — shared internals of libjrb200 (sm_100a).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include "../../include/jrb200.h"

void jr_set_error(const char *fmt, ...);

#define JR_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            jr_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #call); \
            return JR_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define JR_CHECK_LAUNCH() JR_CUDA(cudaGetLastError())

#define JR_REQUIRE(cond, status, ...)                                                          \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            jr_set_error(__VA_ARGS__);                                                         \
            return (status);                                                                   \
        }                                                                                      \
    } while (0)

// A scratch buffer cache: solver temporaries (ping-pong copies, ητ, θ, λv*, reduction slots)
// are owned by the context and reused across solves of the same shape — the reference
// re-allocates them on every solve! call (Stokes3D.jl:55, 494-502).
struct jr_scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct jr_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint32_t flags = 0;
    int sm_count = 148;
    std::map<std::string, jr_scratch> scratch;
    double *h_pinned = nullptr;  // small pinned host buffer for norm read-back
    size_t h_pinned_count = 0;
    int64_t launches = 0;        // kernels launched since last reset
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    struct jr_comm *comm = nullptr;  // multi-GPU communicator (comm.cu); nullptr = single rank
};

int jr_ctx_scratch(jr_context *ctx, const char *key, size_t bytes, void **out);

// ---- 1-based column-major index helpers (mirror kernels keep the reference's indices) ----
#define IX3(n1, n2, i, j, k) ((size_t)((k) - 1) * (size_t)(n2) * (size_t)(n1) + (size_t)((j) - 1) * (size_t)(n1) + (size_t)((i) - 1))
#define IX2(n1, i, j) ((size_t)((j) - 1) * (size_t)(n1) + (size_t)((i) - 1))

__host__ __device__ __forceinline__ int jr_clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ double jr_inv(double x) { return 1.0 / x; }

// warp / block reductions (sum) — used by the residual norms (norm_mpi, src/Utils.jl:698-701)
__device__ __forceinline__ double jr_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// block sum; result valid in thread 0.  `sm` must hold >= 32 doubles.
__device__ __forceinline__ double jr_block_sum(double v, double *sm)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nth = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, wid = tid >> 5;
    v = jr_warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < (nth + 31) / 32) ? sm[lane] : 0.0;
        r = jr_warp_sum(r);
    }
    return r;
}

// Stokes helpers shared by the unfused and fused kernels -------------------------------------
// compute_dτ_r  src/rheology/StressUpdate.jl:70
__device__ __forceinline__ double jr_dtau_r(double theta_dtau, double eta, double _Gdt)
{
    return jr_inv(theta_dtau + fma(eta, _Gdt, 1.0));
}
// compute_stress_increment  src/stokes/StressKernels.jl:2-5
__device__ __forceinline__ double jr_stress_increment(double t, double t_o, double eta, double e, double _Gdt, double dtr)
{
    return dtr * fma(2.0 * eta, e, fma(-(t - t_o) * eta, _Gdt, -t));
}
// compute_stress_increment, strain-increment form  src/stokes/StressKernels.jl:19-22
__device__ __forceinline__ double jr_stress_increment_d(double t, double t_o, double eta, double de, double _G, double dtr, double dt)
{
    return dtr * fma(2.0 * eta, de, fma(-(t - t_o) * eta, _G, -t * dt));
}
// _compute_P!  src/stokes/PressureKernels.jl:186-195
__device__ __forceinline__ void jr_compute_P_point(double &RP, double &P, double P0, double divV, double Q, double eta,
                                                   double K, double G, double dt, double r, double theta_dtau)
{
    const double _Kdt = jr_inv(K * dt);
    const double _Gdt = jr_inv(G * dt);
    const double _dt = jr_inv(dt);
    const double Pc = P;
    RP = fma(-(Pc - P0), _Kdt, (-divV + (Q * _dt)));
    const double psi = jr_inv(jr_inv(eta) + _Gdt) * r / theta_dtau;
    P = ((fma(P0, _Kdt, (-divV + (Q * _dt)))) * psi + Pc) / (1 + _Kdt * psi);
}

// _compute_P! with thermal stresses  src/stokes/PressureKernels.jl:197-206 (Kiss et al. 2023)
__device__ __forceinline__ void jr_compute_P_point_dT(double &RP, double &P, double P0, double divV, double Q, double dT, double alpha,
                                                      double eta, double K, double G, double dt, double r, double theta_dtau)
{
    const double _Kdt = jr_inv(K * dt);
    const double _Gdt = jr_inv(G * dt);
    const double _dt = jr_inv(dt);
    const double Pc = P;
    RP = fma(-(Pc - P0), _Kdt, (-divV + (alpha * (dT * _dt)) + (Q * _dt)));
    const double psi = jr_inv(jr_inv(eta) + _Gdt) * r / theta_dtau;
    P = ((fma(P0, _Kdt, (-divV + (alpha * (dT * _dt)) + (Q * _dt)))) * psi + Pc) / (1 + _Kdt * psi);
}

// internal entry points shared between translation units
int jr_launch_flow_bcs3d(jr_context *ctx, double *Ax, double *Ay, double *Az, const int32_t n[3],
                         const int32_t free_slip[6], const int32_t no_slip[6], const int32_t periodic[6]);
int jr_launch_maxloc3d(jr_context *ctx, double *B, const double *A, const int32_t n[3], const int32_t w[3]);
int jr_launch_sumsq(jr_context *ctx, const double *A, const int32_t n[3], int interior, double *d_out_slot);
int jr_stokes3d_VA_unfused_iter(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o);
int jr_stokes3d_VA_fused_iter(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int write_diag, int parity);
int jr_stokes3d_VA_fused_multi_max(jr_context *ctx, const jr_stokes_opts *o);
int jr_stokes3d_VA_fused_multi(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int niter, int parity);
int jr_stokes3d_VA_fused_supported(const jr_fields *s, const jr_stokes_opts *o);
int jr_stokes3d_VA_fused_begin(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o);
int jr_stokes3d_VA_fused_finish(jr_context *ctx, const jr_fields *s, int64_t niter);
