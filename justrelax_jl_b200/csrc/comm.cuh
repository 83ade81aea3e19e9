// comm.cuh — internal types of the multi-GPU layer (comm.cu).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/jrb200.h"

#define JR_COMM_MAX_RANKS 32
#define JR_COMM_RED_SLOTS 16
#define JR_HALO_MAX_ARRAYS 8

// one per rank, in device memory, CUDA-IPC mapped by every other rank
struct jr_comm_sig {
    unsigned long long flags[JR_COMM_MAX_RANKS];                // flags[r] = last epoch rank r has reached (written by r)
    unsigned long long push_flags[JR_COMM_MAX_RANKS];           // push protocol of the fused 3D-VA kernel: push_flags[r] = r has finished
                                                                // (kernel + BC pushes) every iteration before its launch number …
    double red[2][JR_COMM_MAX_RANKS][JR_COMM_RED_SLOTS];        // all-reduce slots, double-buffered
};

// what kernels need (passed by value as a __grid_constant__ parameter)
struct jr_comm_dev {
    int rank, nranks;
    int nbr[27];                                  // rank at coordinate offset (dx,dy,dz): [(dz+1)*9 + (dy+1)*3 + dx+1], −1 = none
    jr_comm_sig *sig[JR_COMM_MAX_RANKS];          // sig[r]: rank r's signal page (peer mapping; sig[rank] = own)
    double *stage[JR_COMM_MAX_RANKS][2];          // stage[r][b]: rank r's halo staging buffer b
};

struct jr_comm {
    int rank = 0, nranks = 1;
    int dims[3] = {1, 1, 1}, coords[3] = {0, 0, 0};
    int periods[3] = {0, 0, 0};                   // IGG's periodx/y/z: the grid of ranks wraps around in that dimension
    bool has_lo[3] = {false, false, false}, has_hi[3] = {false, false, false};   // a neighbour (possibly this rank itself) on that side
    bool active = false;                          // update_halo! moves data: more than one rank, or a periodic dimension
    jr_allgather_fn allgather = nullptr;
    void *user = nullptr;
    jr_comm_dev dev;
    jr_comm_sig *sig_mine = nullptr;
    void *stage_mine[2] = {nullptr, nullptr};
    size_t stage_cap = 0;                         // doubles per staging buffer
    unsigned long long epoch = 0, red_count = 0, push_epoch = 0, halo_count = 0;
    size_t halo_bytes = 0;
    std::map<std::string, void *> ipc_open;       // opened peer handles (handle bytes + rank → mapped pointer)
    std::vector<void *> retired;                  // outgrown staging buffers (freed at destroy)
};

// an array taking part in a halo update: dense user array or one array of a box set
struct jr_harr {
    double *p;
    long sy, sz;   // strides (elements) of the 2nd and 3rd index
    int n[3];      // extents of the array itself
    int o[3];      // offset of element (0,0,0) inside the addressed space (box layout), 0 for dense arrays
    int ol[3];     // IGG overlap of this array per dimension: 2 + (n[d] − ncell[d]); < 2 = not exchanged
};

struct jr_context;
jr_harr jr_harr_dense(double *p, const int32_t ext[3], const int32_t ncell[3]);
// update_halo!(arrs...) on ctx's communicator (no-op without one / on a single non-periodic rank); asynchronous on ctx->stream
int jr_comm_halo(jr_context *ctx, const jr_harr *arrs, int narr);
size_t jr_comm_halo_bytes(const jr_comm *cm, const jr_harr *arrs, int narr);
// in-place all-reduce of n ≤ 16 device doubles (op 0 sum, 1 max, 2 min); asynchronous on ctx->stream
int jr_comm_allreduce_dev(jr_context *ctx, double *d_vals, int n, int op);
// CUDA-IPC share `mine` (a cudaMalloc base pointer) with every rank: out[r] = rank r's buffer mapped here (out[rank] = mine).
// Collective (one host all-gather); mappings are cached per handle.
int jr_comm_share(jr_context *ctx, void *mine, void **out);
// make sure every rank's staging buffers hold at least `doubles` elements (collective: all ranks call it with the same size)
int jr_comm_reserve_stage(jr_context *ctx, size_t doubles);

// staging layout of one array: planes (d, side) in the order (0,0),(0,1),(1,0),(1,1),(2,0),(2,1); side 0 = plane ol−1
// (wanted by the low neighbour), side 1 = plane n−ol (wanted by the high neighbour)
__host__ __device__ inline long jr_stage_plane_size(const int n[3], int d) { return d == 0 ? (long)n[1] * n[2] : d == 1 ? (long)n[0] * n[2] : (long)n[0] * n[1]; }
__host__ __device__ inline long jr_stage_plane_off(const int n[3], int d, int side)
{
    long off = 0;
    for (int e = 0; e < d; e++) off += 2 * jr_stage_plane_size(n, e);
    return off + side * jr_stage_plane_size(n, d);
}
__host__ __device__ inline long jr_stage_elem(const int n[3], int d, const int s[3])
{
    return d == 0 ? (long)s[2] * n[1] + s[1] : d == 1 ? (long)s[2] * n[0] + s[0] : (long)s[1] * n[0] + s[0];
}

struct HaloArgs {
    jr_harr A[JR_HALO_MAX_ARRAYS];
    long stage_off[JR_HALO_MAX_ARRAYS];  // offset (doubles) of each array's planes in the staging buffer
    int narr;
    int buf;                             // staging buffer 0/1 (alternates per exchange)
    bool has_lo[3], has_hi[3];
};

// exchange in two parts for a caller that packs the send planes itself (see comm.cu)
int jr_comm_halo_begin(jr_context *ctx, const jr_harr *arrs, int narr, HaloArgs *h, double **stage_mine);
int jr_comm_halo_pull(jr_context *ctx, const HaloArgs *h);
// update_halo!(arrs...) split along the slowest index (see comm.cu: k_halo_pull_z)
int jr_comm_halo_z(jr_context *ctx, const jr_harr *arrs, int narr, int pz, int head, int chunk, int rest_ctas, cudaStream_t side, cudaEvent_t ev_head,
                   cudaEvent_t ev_rest, unsigned long long *prog, unsigned long long prog_base);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------
// device-side barrier over all ranks of the communicator (called by every block of a kernel; block 0 signals)
__device__ __forceinline__ void jr_comm_barrier_dev(const jr_comm_dev &cd, unsigned long long epoch)
{
    const int t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    if (t < cd.nranks && t != cd.rank) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            __threadfence_system();
            unsigned long long *f = &cd.sig[t]->flags[cd.rank];
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        }
        const unsigned long long *mine = &cd.sig[cd.rank]->flags[t];
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        } while (v < epoch);
    }
    __syncthreads();
}

#endif
