// phase_ratios.cu — grid-based phase ratios on the B200: update_phase_ratios_{2,3}D! (src/phases/PhaseRatios.jl:21-78, GPU method
// src/ext/CUDA/3D.jl:519-539).  One kernel per staggered location family, operation for operation the reference's kernels:
//   centres   :90-117    vertices :134-212    faces (Vx, Vy, Vz) :232-296    edge midpoints (xy, yz, xz) :328-396
// Input: N phase arrays (nx, ny[, nz]) with values in [0, 1]; output ratios laid out [phase][node] (the CellArray flattened), the layout
// the Stokes / thermal kernels read.  Bit-exact against the oracle (the north star asks bit-exact phase arrays).
#include "common.cuh"

#define PR_MAXP 8
struct PrArgs {
    int nd, nx, ny, nz, N;
    const double *ph[PR_MAXP];
    const double *xc[3], *xv[3];   // device coordinate vectors
    double dx, dy, dz;             // JustPIC.compute_dx(xvi)
    double *out;
    int off[3];                    // staggering of the output location per dimension
    int kind;                      // 0 centre, 1 vertex, 2 face, 3 midpoint
};

__device__ __forceinline__ void pr_finish(double *w, int N, double tw, double *out, size_t stride, size_t idx)
{
    for (int k = 0; k < N; k++) w[k] /= tw;
    for (int k = 0; k < N; k++) w[k] = fmin(fmax(w[k], 0.0), 1.0);
    double total = 0.0;
    for (int k = 0; k < N; k++) { w[k] = w[k] < 1.0e-5 ? 0.0 : w[k]; total += w[k]; }
    for (int k = 0; k < N; k++) out[(size_t)k * stride + idx] = w[k] / total;
}

__global__ void k_phase_ratios(const __grid_constant__ PrArgs a)
{
    const int nx = a.nx, ny = a.ny, nz = a.nz, N = a.N, nd = a.nd;
    const int e0 = nx + a.off[0], e1 = ny + a.off[1], e2 = nz + (nd == 3 ? a.off[2] : 0);
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, l = blockIdx.z + 1;
    if (i > e0 || j > e1 || l > e2) return;
    const size_t st = (size_t)e0 * e1 * e2, idx = IX3(e0, e1, i, j, l);
    double w[PR_MAXP];
#pragma unroll
    for (int k = 0; k < PR_MAXP; k++) w[k] = 0.0;
#define PH(k, ic, jc, lc) a.ph[k][IX3(nx, ny, ic, jc, lc)]
    if (a.kind == 0) {
        double total = 0.0;
#pragma unroll
        for (int k = 0; k < PR_MAXP; k++)
            if (k < N) { w[k] = PH(k, i, j, l); total = k == 0 ? w[k] : total + w[k]; }
#pragma unroll
        for (int k = 0; k < PR_MAXP; k++)
            if (k < N) { double v = w[k] / total; v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); w[k] = v < 1.0e-5 ? 0.0 : v; }
        double ft = 0.0;
#pragma unroll
        for (int k = 0; k < PR_MAXP; k++)
            if (k < N) ft = k == 0 ? w[k] : ft + w[k];
#pragma unroll
        for (int k = 0; k < PR_MAXP; k++)
            if (k < N) a.out[(size_t)k * st + idx] = w[k] / ft;
        return;
    }
    double tw = 0.0;
    if (a.kind == 1) {
        for (int o1 = -1; o1 <= 0; o1++)
            for (int o2 = -1; o2 <= 0; o2++)
                for (int o3 = (nd == 3 ? -1 : 0); o3 <= 0; o3++) {
                    const int ic = i + o1, jc = j + o2, lc = nd == 3 ? l + o3 : 1;
                    if (!(1 <= ic && ic <= nx && 1 <= jc && jc <= ny && 1 <= lc && lc <= nz)) continue;
                    double weight;
                    if (nd == 2) {
                        const double wx = fma(-fabs(a.xv[0][i - 1] - a.xc[0][ic - 1]), 1.0 / a.dx, 1.0);
                        const double wy = fma(-fabs(a.xv[1][j - 1] - a.xc[1][jc - 1]), 1.0 / a.dy, 1.0);
                        weight = wx * wy;
                    } else {
                        weight = 1.0;
                        weight *= (1.0 - fabs(a.xv[0][i - 1] - a.xc[0][ic - 1]) * (1.0 / a.dx));
                        weight *= (1.0 - fabs(a.xv[1][j - 1] - a.xc[1][jc - 1]) * (1.0 / a.dy));
                        weight *= (1.0 - fabs(a.xv[2][l - 1] - a.xc[2][lc - 1]) * (1.0 / a.dz));
                    }
                    tw += weight;
#pragma unroll
                    for (int k = 0; k < PR_MAXP; k++)
                        if (k < N) w[k] += weight * PH(k, ic, jc, lc);
                }
    } else if (a.kind == 2) {
        for (int side = 0; side <= 1; side++) {
            const int ic = a.off[0] ? i - 1 + side : i, jc = a.off[1] ? j - 1 + side : j, lc = (nd == 3 && a.off[2]) ? l - 1 + side : l;
            if (!(1 <= ic && ic <= nx && 1 <= jc && jc <= ny && 1 <= lc && lc <= nz)) continue;
            tw += 0.5;
#pragma unroll
            for (int k = 0; k < PR_MAXP; k++)
                if (k < N) w[k] += 0.5 * PH(k, ic, jc, lc);
        }
    } else {
        for (int corner = 1; corner <= 4; corner++) {
            const int first = corner <= 2 ? -1 : 0, second = (corner & 1) ? -1 : 0;
            const int ic = i + a.off[0] * first, jo = a.off[0] == 1 ? second : first, jc = j + a.off[1] * jo, lc = l + a.off[2] * second;
            if (!(1 <= ic && ic <= nx && 1 <= jc && jc <= ny && 1 <= lc && lc <= nz)) continue;
            tw += 0.25;
#pragma unroll
            for (int k = 0; k < PR_MAXP; k++)
                if (k < N) w[k] += 0.25 * PH(k, ic, jc, lc);
        }
    }
#undef PH
    pr_finish(w, N, tw, a.out, st, idx);
}

extern "C" int jr_phase_ratios_from_arrays(jr_context *ctx, int32_t ndim, const int32_t n[3], int32_t nphase, const double *const *phase_arrays,
                                           const double *const *xci_host, const double *const *xvi_host, double *center, double *vertex, double *Vx,
                                           double *Vy, double *Vz, double *xy, double *yz, double *xz)
{
    JR_REQUIRE(ctx && n && phase_arrays && xci_host && xvi_host, JR_ERR_ARG, "jr_phase_ratios_from_arrays: null argument");
    JR_REQUIRE(ndim == 2 || ndim == 3, JR_ERR_SHAPE, "ndim must be 2 or 3");
    JR_REQUIRE(nphase >= 1 && nphase <= PR_MAXP, JR_ERR_UNSUPPORTED, "number of phases %d outside 1..%d", nphase, PR_MAXP);
    JR_CUDA(cudaSetDevice(ctx->device));
    PrArgs a;
    memset(&a, 0, sizeof(a));
    a.nd = ndim; a.nx = n[0]; a.ny = n[1]; a.nz = ndim == 3 ? n[2] : 1; a.N = nphase;
    for (int k = 0; k < nphase; k++) { JR_REQUIRE(phase_arrays[k], JR_ERR_ARG, "phase array %d is NULL", k); a.ph[k] = phase_arrays[k]; }
    // coordinates: host vectors → one scratch buffer [xc0 | xc1 | xc2 | xv0 | xv1 | xv2]
    size_t tot = 0;
    for (int d = 0; d < ndim; d++) tot += (size_t)(2 * n[d] + 1);
    void *buf = nullptr;
    int st = jr_ctx_scratch(ctx, "phase_ratio_coords", tot * sizeof(double), &buf);
    if (st) return st;
    double *p = (double *)buf;
    for (int d = 0; d < ndim; d++) {
        JR_REQUIRE(xci_host[d] && xvi_host[d] && n[d] >= 1, JR_ERR_ARG, "coordinate vector %d is NULL", d);
        JR_CUDA(cudaMemcpyAsync(p, xci_host[d], n[d] * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        a.xc[d] = p; p += n[d];
        JR_CUDA(cudaMemcpyAsync(p, xvi_host[d], (n[d] + 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        a.xv[d] = p; p += n[d] + 1;
    }
    JR_CUDA(cudaStreamSynchronize(ctx->stream));  // the host vectors may be temporaries of the caller
    a.dx = xvi_host[0][1] - xvi_host[0][0];
    a.dy = xvi_host[1][1] - xvi_host[1][0];
    a.dz = ndim == 3 ? xvi_host[2][1] - xvi_host[2][0] : 1.0;
    struct Job { double *out; int kind, o0, o1, o2; };
    const Job jobs[8] = {{center, 0, 0, 0, 0}, {vertex, 1, 1, 1, 1}, {Vx, 2, 1, 0, 0}, {Vy, 2, 0, 1, 0}, {ndim == 3 ? Vz : nullptr, 2, 0, 0, 1},
                         {ndim == 3 ? xy : nullptr, 3, 1, 1, 0}, {ndim == 3 ? yz : nullptr, 3, 0, 1, 1}, {ndim == 3 ? xz : nullptr, 3, 1, 0, 1}};
    for (const Job &jb : jobs) {
        if (!jb.out) continue;
        a.out = jb.out; a.kind = jb.kind; a.off[0] = jb.o0; a.off[1] = jb.o1; a.off[2] = jb.o2;
        const int e0 = a.nx + jb.o0, e1 = a.ny + jb.o1, e2 = a.nz + (ndim == 3 ? jb.o2 : 0);
        k_phase_ratios<<<dim3((e0 + 31) / 32, (e1 + 7) / 8, e2), dim3(32, 8, 1), 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}
