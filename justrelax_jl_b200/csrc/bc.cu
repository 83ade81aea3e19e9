// bc.cu — velocity boundary-condition kernels (flow_bcs!) and compute_maxloc! for libjrb200.
//
// Reference: src/boundaryconditions/BoundaryConditions.jl:65-100 (_flow_bcs!: no_slip → free_slip → periodic),
// free_slip.jl:15-70, no_slip.jl:21-54, periodic.jl:56-98, src/Utils.jl:409-461.
//
// Flag order everywhere: left,right,front,back,top,bot.  Quirk Q2: in the reference's 3-D
// free_slip! `top` fills k=1 and `bot` fills k=end, while no_slip!/periodic use bot→k=1, top→k=end.
//
// free-slip is implemented as ONE race-free gather launch: every ghost element is read from the
// interior element that the reference's front/back → top/bot → left/right sequence ends up copying
// (ghost edges included), so the result is the deterministic fixed point of the reference kernel.
#include "common.cuh"

struct Arr3 {
    double *p;
    int n1, n2, n3;
};
__device__ __forceinline__ double &at(const Arr3 &A, int i, int j, int k) { return A.p[IX3(A.n1, A.n2, i, j, k)]; }

// map a (1-based) index onto its free-slip source index: ghost layer → adjacent interior layer
__device__ __forceinline__ int fs_src(int i, int n, bool lo, bool hi) { return (lo && i == 1) ? 2 : ((hi && i == n) ? n - 1 : i); }

// one array, ghost dims flagged by (lo1,hi1),(lo2,hi2),(lo3,hi3) — false for the normal direction
__device__ __forceinline__ void fs_fix(const Arr3 &A, int i, int j, int k, bool l1, bool h1, bool l2, bool h2, bool l3, bool h3)
{
    if (i > A.n1 || j > A.n2 || k > A.n3) return;
    const int si = fs_src(i, A.n1, l1, h1), sj = fs_src(j, A.n2, l2, h2), sk = fs_src(k, A.n3, l3, h3);
    if (si != i || sj != j || sk != k) at(A, i, j, k) = at(A, si, sj, sk);
}

__global__ void k_free_slip3(Arr3 Ax, Arr3 Ay, Arr3 Az, int left, int right, int front, int back, int top, int bot)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int b = blockIdx.y * blockDim.y + threadIdx.y + 1;
    // ghost dims: Ax: y (front/back), z (top→k=1 / bot→k=end);  Ay: x (left/right), z;  Az: x, y
    const bool zl = top, zh = bot;  // Q2
    // y-normal planes j ∈ {1,end}: (i=a, k=b)
    fs_fix(Ax, a, 1, b, false, false, front, back, zl, zh);
    fs_fix(Ax, a, Ax.n2, b, false, false, front, back, zl, zh);
    fs_fix(Az, a, 1, b, left, right, front, back, false, false);
    fs_fix(Az, a, Az.n2, b, left, right, front, back, false, false);
    // z-normal planes k ∈ {1,end}: (i=a, j=b)
    fs_fix(Ax, a, b, 1, false, false, front, back, zl, zh);
    fs_fix(Ax, a, b, Ax.n3, false, false, front, back, zl, zh);
    fs_fix(Ay, a, b, 1, left, right, false, false, zl, zh);
    fs_fix(Ay, a, b, Ay.n3, left, right, false, false, zl, zh);
    // x-normal planes i ∈ {1,end}: (j=a, k=b)
    fs_fix(Ay, 1, a, b, left, right, false, false, zl, zh);
    fs_fix(Ay, Ay.n1, a, b, left, right, false, false, zl, zh);
    fs_fix(Az, 1, a, b, left, right, front, back, false, false);
    fs_fix(Az, Az.n1, a, b, left, right, front, back, false, false);
}

// no_slip! — sequential broadcasts in the reference (left,right,front,back,bot,top); one launch per face.
// face: 0 left,1 right,2 front,3 back,4 top(k=end),5 bot(k=1)
__global__ void k_no_slip3_face(Arr3 Ax, Arr3 Ay, Arr3 Az, int face)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int b = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (face == 0 || face == 1) { // (j=a,k=b)
        const bool lo = face == 0;
        if (a <= Ax.n2 && b <= Ax.n3) at(Ax, lo ? 1 : Ax.n1, a, b) = 0.0;
        if (a <= Ay.n2 && b <= Ay.n3) at(Ay, lo ? 1 : Ay.n1, a, b) = -at(Ay, lo ? 2 : Ay.n1 - 1, a, b);
        if (a <= Az.n2 && b <= Az.n3) at(Az, lo ? 1 : Az.n1, a, b) = -at(Az, lo ? 2 : Az.n1 - 1, a, b);
    } else if (face == 2 || face == 3) { // (i=a,k=b)
        const bool lo = face == 2;
        if (a <= Ax.n1 && b <= Ax.n3) at(Ax, a, lo ? 1 : Ax.n2, b) = -at(Ax, a, lo ? 2 : Ax.n2 - 1, b);
        if (a <= Ay.n1 && b <= Ay.n3) at(Ay, a, lo ? 1 : Ay.n2, b) = 0.0;
        if (a <= Az.n1 && b <= Az.n3) at(Az, a, lo ? 1 : Az.n2, b) = -at(Az, a, lo ? 2 : Az.n2 - 1, b);
    } else { // (i=a,j=b); bot → k=1, top → k=end
        const bool lo = face == 5;
        if (a <= Ax.n1 && b <= Ax.n2) at(Ax, a, b, lo ? 1 : Ax.n3) = -at(Ax, a, b, lo ? 2 : Ax.n3 - 1);
        if (a <= Ay.n1 && b <= Ay.n2) at(Ay, a, b, lo ? 1 : Ay.n3) = -at(Ay, a, b, lo ? 2 : Ay.n3 - 1);
        if (a <= Az.n1 && b <= Az.n2) at(Az, a, b, lo ? 1 : Az.n3) = 0.0;
    }
}

// periodic_boundary! (Vx,Vy,Vz)  periodic.jl:56-98; three sweeps x, y, z (one launch each) so that
// ghost edges are well defined.
__global__ void k_periodic3_dim(Arr3 Ax, Arr3 Ay, Arr3 Az, int dim, int lo, int hi)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int b = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (dim == 0) { // left/right, (j=a,k=b)
        if (a <= Ax.n2 && b <= Ax.n3 && lo) at(Ax, 1, a, b) = at(Ax, Ax.n1, a, b);
        if (a <= Ay.n2 && b <= Ay.n3) {
            if (lo) at(Ay, 1, a, b) = at(Ay, Ay.n1 - 1, a, b);
            if (hi) at(Ay, Ay.n1, a, b) = at(Ay, 2, a, b);
        }
        if (a <= Az.n2 && b <= Az.n3) {
            if (lo) at(Az, 1, a, b) = at(Az, Az.n1 - 1, a, b);
            if (hi) at(Az, Az.n1, a, b) = at(Az, 2, a, b);
        }
    } else if (dim == 1) { // front/back, (i=a,k=b)
        if (a <= Ax.n1 && b <= Ax.n3) {
            if (lo) at(Ax, a, 1, b) = at(Ax, a, Ax.n2 - 1, b);
            if (hi) at(Ax, a, Ax.n2, b) = at(Ax, a, 2, b);
        }
        if (a <= Ay.n1 && b <= Ay.n3 && lo) at(Ay, a, 1, b) = at(Ay, a, Ay.n2, b);
        if (a <= Az.n1 && b <= Az.n3) {
            if (lo) at(Az, a, 1, b) = at(Az, a, Az.n2 - 1, b);
            if (hi) at(Az, a, Az.n2, b) = at(Az, a, 2, b);
        }
    } else { // bot (k=1) / top (k=end), (i=a,j=b)
        if (a <= Ax.n1 && b <= Ax.n2) {
            if (lo) at(Ax, a, b, 1) = at(Ax, a, b, Ax.n3 - 1);
            if (hi) at(Ax, a, b, Ax.n3) = at(Ax, a, b, 2);
        }
        if (a <= Ay.n1 && b <= Ay.n2) {
            if (lo) at(Ay, a, b, 1) = at(Ay, a, b, Ay.n3 - 1);
            if (hi) at(Ay, a, b, Ay.n3) = at(Ay, a, b, 2);
        }
        if (a <= Az.n1 && b <= Az.n2 && lo) at(Az, a, b, 1) = at(Az, a, b, Az.n3);
    }
}

static inline int any6(const int32_t b[6]) { return b[0] | b[1] | b[2] | b[3] | b[4] | b[5]; }

int jr_launch_flow_bcs3d(jr_context *ctx, double *Ax, double *Ay, double *Az, const int32_t n[3],
                         const int32_t fs[6], const int32_t ns[6], const int32_t pe[6])
{
    const int nx = n[0], ny = n[1], nz = n[2];
    Arr3 X{Ax, nx + 1, ny + 2, nz + 2}, Y{Ay, nx + 2, ny + 1, nz + 2}, Z{Az, nx + 2, ny + 2, nz + 1};
    int m = nx > ny ? nx : ny;
    m = (m > nz ? m : nz) + 2;
    dim3 blk(32, 8, 1), grd((m + 31) / 32, (m + 7) / 8, 1);
    if (any6(ns)) {
        const int order[6] = {0, 1, 2, 3, 5, 4};
        for (int q = 0; q < 6; q++) {
            const int f = order[q];
            if (!ns[f]) continue;
            k_no_slip3_face<<<grd, blk, 0, ctx->stream>>>(X, Y, Z, f);
            ctx->launches++;
        }
    }
    if (any6(fs)) {
        k_free_slip3<<<grd, blk, 0, ctx->stream>>>(X, Y, Z, fs[0], fs[1], fs[2], fs[3], fs[4], fs[5]);
        ctx->launches++;
    }
    if (any6(pe)) {
        if (pe[0] | pe[1]) { k_periodic3_dim<<<grd, blk, 0, ctx->stream>>>(X, Y, Z, 0, pe[0], pe[1]); ctx->launches++; }
        if (pe[2] | pe[3]) { k_periodic3_dim<<<grd, blk, 0, ctx->stream>>>(X, Y, Z, 1, pe[2], pe[3]); ctx->launches++; }
        if (pe[4] | pe[5]) { k_periodic3_dim<<<grd, blk, 0, ctx->stream>>>(X, Y, Z, 2, pe[5], pe[4]); ctx->launches++; }
    }
    JR_CHECK_LAUNCH();
    return JR_OK;
}

// compute_maxloc! / _maxloc_window_clamped  src/Utils.jl:409-461
__global__ void k_maxloc3(double *__restrict__ B, const double *__restrict__ A, int nx, int ny, int nz, int wx, int wy, int wz)
{
    const int I = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int J = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int K = blockIdx.z + 1;
    if (I > nx || J > ny || K > nz) return;
    double x = -INFINITY;
    for (int k = K - wz; k <= K + wz; k++) {
        const int kk = jr_clamp(k, 1, nz);
        for (int j = J - wy; j <= J + wy; j++) {
            const int jj = jr_clamp(j, 1, ny);
            for (int i = I - wx; i <= I + wx; i++) {
                const int ii = jr_clamp(i, 1, nx);
                const double a = A[IX3(nx, ny, ii, jj, kk)];
                if (a > x) x = a;
            }
        }
    }
    B[IX3(nx, ny, I, J, K)] = x;
}

// window (1,1,1), the one the solvers use: z-marching, the clamped 3×3 in-plane maxima of planes k−1, k, k+1 roll through registers
// (9 loads per cell instead of 27; the max of maxima is the same value as the reference's 27-point scan: max is exact)
#define MAXLOC_KCH 32
__global__ void __launch_bounds__(256) k_maxloc3_w1(double *__restrict__ B, const double *__restrict__ A, int nx, int ny, int nz)
{
    const int I = blockIdx.x * blockDim.x + threadIdx.x + 1, J = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (I > nx || J > ny) return;
    const int k0 = blockIdx.z * MAXLOC_KCH + 1, k1 = min(k0 + MAXLOC_KCH - 1, nz);
    const int i0 = jr_clamp(I - 1, 1, nx), i1 = jr_clamp(I + 1, 1, nx), j0 = jr_clamp(J - 1, 1, ny), j1 = jr_clamp(J + 1, 1, ny);
    auto plane_max = [&](int k) {
        const int kk = jr_clamp(k, 1, nz);
        const double *p = A + (size_t)(kk - 1) * nx * ny;
        const double a0 = p[(size_t)(j0 - 1) * nx + i0 - 1], a1 = p[(size_t)(j0 - 1) * nx + I - 1], a2 = p[(size_t)(j0 - 1) * nx + i1 - 1];
        const double b0 = p[(size_t)(J - 1) * nx + i0 - 1], b1 = p[(size_t)(J - 1) * nx + I - 1], b2 = p[(size_t)(J - 1) * nx + i1 - 1];
        const double c0 = p[(size_t)(j1 - 1) * nx + i0 - 1], c1 = p[(size_t)(j1 - 1) * nx + I - 1], c2 = p[(size_t)(j1 - 1) * nx + i1 - 1];
        return fmax(fmax(fmax(a0, a1), fmax(a2, b0)), fmax(fmax(b1, b2), fmax(fmax(c0, c1), c2)));
    };
    double lo = plane_max(k0 - 1), mid = plane_max(k0);
    for (int k = k0; k <= k1; k++) {
        const double hi = plane_max(k + 1);
        B[IX3(nx, ny, I, J, k)] = fmax(fmax(lo, mid), hi);
        lo = mid; mid = hi;
    }
}

int jr_launch_maxloc3d(jr_context *ctx, double *B, const double *A, const int32_t n[3], const int32_t w[3])
{
    if (w[0] == 1 && w[1] == 1 && w[2] == 1) {
        dim3 blk1(32, 8, 1), grd1((n[0] + 31) / 32, (n[1] + 7) / 8, (n[2] + MAXLOC_KCH - 1) / MAXLOC_KCH);
        k_maxloc3_w1<<<grd1, blk1, 0, ctx->stream>>>(B, A, n[0], n[1], n[2]);
        ctx->launches++;
        JR_CHECK_LAUNCH();
        return JR_OK;
    }
    dim3 blk(32, 8, 1), grd((n[0] + 31) / 32, (n[1] + 7) / 8, n[2]);
    k_maxloc3<<<grd, blk, 0, ctx->stream>>>(B, A, n[0], n[1], n[2], w[0], w[1], w[2]);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

extern "C" {

int jr_flow_bcs3d(jr_context *ctx, double *Ax, double *Ay, double *Az, const int32_t n[3],
                  const int32_t free_slip[6], const int32_t no_slip[6], const int32_t periodic[6])
{
    JR_REQUIRE(ctx && Ax && Ay && Az && n, JR_ERR_ARG, "jr_flow_bcs3d: null argument");
    JR_REQUIRE(n[0] >= 2 && n[1] >= 2 && n[2] >= 2, JR_ERR_SHAPE, "jr_flow_bcs3d: grid too small");
    int st = jr_launch_flow_bcs3d(ctx, Ax, Ay, Az, n, free_slip, no_slip, periodic);
    if (st) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_maxloc3d(jr_context *ctx, double *B, const double *A, const int32_t n[3], const int32_t window[3])
{
    JR_REQUIRE(ctx && A && B && n && window, JR_ERR_ARG, "jr_maxloc3d: null argument");
    JR_REQUIRE(A != B, JR_ERR_ARG, "jr_maxloc3d: in-place not allowed");
    int st = jr_launch_maxloc3d(ctx, B, A, n, window);
    if (st) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

} // extern "C"
