// api.cu — context, memory helpers, error reporting, reductions of libjrb200 (sm_100a).
#include "common.cuh"
#include "comm.cuh"
#include <cstdarg>

static thread_local char g_err[1024] = "";

void jr_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static const char *k_field_names[] = {
#define X(n) #n,
    JR_STOKES_FIELDS(X)
#undef X
};

extern "C" {

const char *jr_last_error(void) { return g_err; }
int jr_abi_version(void) { return JRB200_ABI_VERSION; }
int jr_field_count(void) { return JR_F_COUNT; }
const char *jr_field_name(int i) { return (i >= 0 && i < JR_F_COUNT) ? k_field_names[i] : nullptr; }

int jr_context_create(int device, void *stream, jr_context **out)
{
    JR_REQUIRE(out != nullptr, JR_ERR_ARG, "jr_context_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        jr_set_error("jr_context_create: no CUDA device available (%s); libjrb200 has no CPU fallback",
                     cudaGetErrorString(e));
        return JR_ERR_CUDA;
    }
    JR_REQUIRE(device >= 0 && device < ndev, JR_ERR_ARG, "jr_context_create: device %d out of range (%d devices)", device, ndev);
    JR_CUDA(cudaSetDevice(device));
    jr_context *ctx = new jr_context();
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
        ctx->own_stream = false;
    } else {
        JR_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    cudaDeviceProp prop;
    JR_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->h_pinned_count = 64;
    JR_CUDA(cudaMallocHost((void **)&ctx->h_pinned, ctx->h_pinned_count * sizeof(double)));
    JR_CUDA(cudaEventCreate(&ctx->ev0));
    JR_CUDA(cudaEventCreate(&ctx->ev1));
    *out = ctx;
    return JR_OK;
}

int jr_context_destroy(jr_context *ctx)
{
    if (!ctx) return JR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->scratch)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return JR_OK;
}

int jr_context_set_flags(jr_context *ctx, uint32_t flags)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    ctx->flags = flags;
    return JR_OK;
}

int jr_context_synchronize(jr_context *ctx)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_malloc(jr_context *ctx, size_t bytes, void **dptr)
{
    JR_REQUIRE(ctx && dptr, JR_ERR_ARG, "jr_malloc: null argument");
    JR_CUDA(cudaSetDevice(ctx->device));
    JR_CUDA(cudaMalloc(dptr, bytes ? bytes : 8));
    return JR_OK;
}
int jr_free(jr_context *ctx, void *dptr)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    JR_CUDA(cudaSetDevice(ctx->device));
    JR_CUDA(cudaFree(dptr));
    return JR_OK;
}
int jr_memcpy_h2d(jr_context *ctx, void *dst, const void *src, size_t bytes)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    JR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}
int jr_memcpy_d2h(jr_context *ctx, void *dst, const void *src, size_t bytes)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    JR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}
int jr_memcpy_d2d(jr_context *ctx, void *dst, const void *src, size_t bytes)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    JR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return JR_OK;
}

} // extern "C"

int jr_ctx_scratch(jr_context *ctx, const char *key, size_t bytes, void **out)
{
    jr_scratch &s = ctx->scratch[key];
    if (s.bytes < bytes) {
        if (s.ptr) JR_CUDA(cudaFree(s.ptr));
        s.ptr = nullptr;
        s.bytes = 0;
        JR_CUDA(cudaMalloc(&s.ptr, bytes));
        s.bytes = bytes;
    }
    *out = s.ptr;
    return JR_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_fill(double *__restrict__ p, double v, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// velocity2displacement! / displacement2velocity!  src/types/displacement.jl:17-60 (dst = src * factor)
__global__ void k_scale_copy(double *__restrict__ dst, const double *__restrict__ src, double f, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i] * f;
}

// Σ A[2:end-1,2:end-1,2:end-1]^2 — deterministic two-stage reduction.
//   stage 1: one block per (j,k) row group, fixed block→data mapping, per-block partial
//   stage 2: one block adds the partials in index order
#define SUMSQ_BLOCKS 1024
__global__ void k_sumsq_stage1(const double *__restrict__ A, int n1, int n2, int n3, int o, int ko, double *__restrict__ part)
{
    __shared__ double sm[32];
    const int m1 = n1 - 2 * o, m2 = n2 - 2 * o, m3 = n3 - 2 * ko;
    const size_t rows = (size_t)m2 * m3;
    double acc = 0.0;
    for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const int j = (int)(r % m2) + o, k = (int)(r / m2) + ko;
        const double *row = A + ((size_t)k * n2 + j) * n1 + o;
        for (int i = threadIdx.x; i < m1; i += blockDim.x) {
            const double a = row[i];
            acc += a * a;
        }
    }
    const double s = jr_block_sum(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void k_sumsq_stage2(const double *__restrict__ part, int nparts, double *__restrict__ out)
{
    __shared__ double sm[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += part[i];
    const double s = jr_block_sum(acc, sm);
    if (threadIdx.x == 0) *out = s;
}


// maximum(abs.(A)) — the reduction behind compute_dt (src/Utils.jl:492-519); NaN propagates like Julia's maximum
__global__ void k_absmax_stage1(const double *__restrict__ A, size_t n, double *__restrict__ part)
{
    __shared__ double sm[32];
    double m = 0.0;
    bool nan = false;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = fabs(A[i]);
        if (v != v) nan = true;
        m = fmax(m, v);
    }
    if (nan) m = NAN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_down_sync(0xffffffffu, m, o);
        m = (m != m || w != w) ? NAN : fmax(m, w);
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = sm[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = (r != r || sm[w] != sm[w]) ? NAN : fmax(r, sm[w]);
        part[blockIdx.x] = r;
    }
}
__global__ void k_absmax_stage2(const double *__restrict__ part, int nparts, double *__restrict__ out)
{
    if (threadIdx.x == 0) {
        double r = part[0];
        for (int b = 1; b < nparts; b++) r = (r != r || part[b] != part[b]) ? NAN : fmax(r, part[b]);
        *out = r;
    }
}

int jr_launch_sumsq(jr_context *ctx, const double *A, const int32_t n[3], int interior, double *d_out_slot)
{
    const int o = interior ? 1 : 0;
    const int ko = (n[2] > 1) ? o : 0;
    void *part = nullptr;
    int st = jr_ctx_scratch(ctx, "sumsq_part", SUMSQ_BLOCKS * sizeof(double), &part);
    if (st) return st;
    const long m1 = n[0] - 2 * o, m2 = n[1] - 2 * o, m3 = n[2] - 2 * ko;
    if (m1 <= 0 || m2 <= 0 || m3 <= 0) {
        JR_CUDA(cudaMemsetAsync(d_out_slot, 0, sizeof(double), ctx->stream));
        return JR_OK;
    }
    k_sumsq_stage1<<<SUMSQ_BLOCKS, 256, 0, ctx->stream>>>(A, n[0], n[1], n[2], o, ko, (double *)part);
    k_sumsq_stage2<<<1, 256, 0, ctx->stream>>>((const double *)part, SUMSQ_BLOCKS, d_out_slot);
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

extern "C" {

int jr_fill_f64(jr_context *ctx, double *dptr, double value, size_t count)
{
    JR_REQUIRE(ctx && dptr, JR_ERR_ARG, "jr_fill_f64: null argument");
    if (!count) return JR_OK;
    int blocks = (int)((count + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_fill<<<blocks, 256, 0, ctx->stream>>>(dptr, value, count);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

int jr_scale_copy(jr_context *ctx, double *dst, const double *src, double factor, size_t count)
{
    JR_REQUIRE(ctx && dst && src, JR_ERR_ARG, "jr_scale_copy: null argument");
    if (!count) return JR_OK;
    int blocks = (int)((count + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_scale_copy<<<blocks, 256, 0, ctx->stream>>>(dst, src, factor, count);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

int jr_sumsq(jr_context *ctx, const double *A, const int32_t n[3], int interior, double *out_host)
{
    JR_REQUIRE(ctx && A && out_host, JR_ERR_ARG, "jr_sumsq: null argument");
    void *slot = nullptr;
    int st = jr_ctx_scratch(ctx, "norm_slots", 16 * sizeof(double), &slot);
    if (st) return st;
    st = jr_launch_sumsq(ctx, A, n, interior, (double *)slot);
    if (st) return st;
    JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_host = ctx->h_pinned[0];
    return JR_OK;
}


int jr_absmax(jr_context *ctx, const double *A, size_t count, int allreduce, double *out_host)
{
    JR_REQUIRE(ctx && A && out_host && count > 0, JR_ERR_ARG, "jr_absmax: null argument");
    JR_CUDA(cudaSetDevice(ctx->device));
    void *buf = nullptr;
    const int nb = 512;
    int st = jr_ctx_scratch(ctx, "absmax_part", (nb + 16) * sizeof(double), &buf);
    if (st) return st;
    double *part = (double *)buf, *out = part + nb;
    k_absmax_stage1<<<nb, 256, 0, ctx->stream>>>(A, count, part);
    k_absmax_stage2<<<1, 32, 0, ctx->stream>>>(part, nb, out);
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    if (allreduce && (st = jr_comm_allreduce_dev(ctx, out, 1, 1))) return st;   // maximum_mpi
    JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_host = ctx->h_pinned[0];
    return JR_OK;
}

} // extern "C"
