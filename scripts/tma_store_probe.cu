// tma_store_probe.cu — which 4-D bulk tensor STORE shapes does the sm_100a TMA unit accept?  (development probe for k_va_tma's
// TSTORE path; not part of the library.)  One configuration per process: an illegal configuration poisons the context.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_store_probe tma_store_probe.cu -lcudart
//   ./tma_store_probe BOXX BOXY NBOX X0 Y0 A0 Z0 HINT SMEM_OFF_BYTES
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Args {
    CUtensorMap m;
    int x0, y0, a0, z0, hint, off, n, cx;
};

__global__ void k_probe(const __grid_constant__ Args a)
{
    extern __shared__ __align__(128) unsigned char raw[];
    double *src = reinterpret_cast<double *>(raw + a.off);
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) src[i] = 1000.0 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(src);
        if (a.hint) {
            uint64_t p;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;" ::"l"(
                             (uint64_t)&a.m),
                         "r"(s), "r"(a.cx), "r"(a.y0), "r"(a.a0), "r"(a.z0), "l"(p)
                         : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)&a.m), "r"(s),
                         "r"(a.cx), "r"(a.y0), "r"(a.a0), "r"(a.z0)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void k_probe_load(const __grid_constant__ Args a, double *out)
{
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    double *dst = reinterpret_cast<double *>(raw + a.off);
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(a.n * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(dst)),
                     "l"((uint64_t)&a.m), "r"(a.cx), "r"(a.y0), "r"(a.a0), "r"(a.z0), "r"(b)
                     : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) out[i] = dst[i];
}

int main(int argc, char **argv)
{
    if (argc < 10) return 2;
    const int bx = atoi(argv[1]), by = atoi(argv[2]), nb = atoi(argv[3]);
    Args a;
    a.x0 = atoi(argv[4]); a.y0 = atoi(argv[5]); a.a0 = atoi(argv[6]); a.z0 = atoi(argv[7]); a.hint = atoi(argv[8]); a.off = atoi(argv[9]);
    a.n = bx * by * nb;
    const int PX = 68, PY = 37, NA = 10, PZ = 6;
    const size_t n = (size_t)PX * PY * NA * PZ;
    double *d;
    cudaMalloc(&d, n * 8);
    cudaMemset(d, 0, n * 8);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    const bool as32 = argc > 11 && argv[11][0] == '4';   // describe the f64 tensor as 32-bit words (inner extent, box and coordinate doubled)
    cuuint64_t gd[4] = {(cuuint64_t)(as32 ? 2 * PX : PX), PY, NA, PZ}, gs[3] = {PX * 8ull, PX * PY * 8ull, (cuuint64_t)NA * PX * PY * 8ull};
    cuuint32_t bd[4] = {(cuuint32_t)(as32 ? 2 * bx : bx), (cuuint32_t)by, (cuuint32_t)nb, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((PFN_encodeTiled)p)(&a.m, as32 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    a.cx = as32 ? 2 * a.x0 : a.x0;
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int smem = a.off + a.n * 8;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (argc > 10 && argv[10][0] == 'L') {
        // load mode: the tensor holds f(x,y,a,z); the box (OOB elements read as 0) must arrive dense in shared memory
        std::vector<double> hh(n);
        for (size_t i = 0; i < n; i++) hh[i] = 1.0 + (double)i;
        cudaMemcpy(d, hh.data(), n * 8, cudaMemcpyHostToDevice);
        double *o;
        cudaMalloc(&o, a.n * 8);
        cudaFuncSetAttribute(k_probe_load, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k_probe_load<<<1, 128, smem>>>(a, o);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e2 != cudaSuccess) { printf("FAIL(load) %s\n", cudaGetErrorString(e2)); return 1; }
        std::vector<double> ho(a.n);
        cudaMemcpy(ho.data(), o, a.n * 8, cudaMemcpyDeviceToHost);
        long bad2 = 0;
        for (int la = 0; la < nb; la++)
            for (int ly = 0; ly < by; ly++)
                for (int lx = 0; lx < bx; lx++) {
                    const int x = a.x0 + lx, y = a.y0 + ly, aa = a.a0 + la, z = a.z0;
                    const bool in = x >= 0 && x < PX && y >= 0 && y < PY && aa >= 0 && aa < NA && z >= 0 && z < PZ;
                    const double want = in ? hh[(((size_t)z * NA + aa) * PY + y) * PX + x] : 0.0;
                    if (ho[(la * by + ly) * bx + lx] != want) bad2++;
                }
        printf("OK load; %d elements, %ld mismatches\n", a.n, bad2);
        return bad2 != 0;
    }
    k_probe<<<1, 128, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAIL %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<double> h(n);
    cudaMemcpy(h.data(), d, n * 8, cudaMemcpyDeviceToHost);
    long bad = 0, hit = 0;
    for (int z = 0; z < PZ; z++)
        for (int aa = 0; aa < NA; aa++)
            for (int y = 0; y < PY; y++)
                for (int x = 0; x < PX; x++) {
                    const double v = h[(((size_t)z * NA + aa) * PY + y) * PX + x];
                    const int lx = x - a.x0, ly = y - a.y0, la = aa - a.a0;
                    const bool in = z == a.z0 && lx >= 0 && lx < bx && ly >= 0 && ly < by && la >= 0 && la < nb;
                    const double want = in ? 1000.0 + ((la * by + ly) * bx + lx) : 0.0;
                    if (v != want) bad++;
                    if (in) hit++;
                }
    printf("OK launch; %ld in-box elements, %ld mismatches\n", hit, bad);
    return bad != 0;
}
