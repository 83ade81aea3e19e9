// comm.cu — multi-GPU plumbing of libjrb200: ImplicitGlobalGrid-compatible halo exchange and reductions over
// CUDA-IPC peer memory (NVLink 5 / NVSwitch), one process per GPU.
//
// Replaces, on the hot path, `update_halo!(A...)` of ImplicitGlobalGrid (call sites src/stokes/Stokes3D.jl:57,120,
// 515,578-580,596; Stokes2D.jl:655,757,784; src/thermal_diffusion/DiffusionPT_solver.jl:110,261) and the
// `MPI.Allreduce` behind norm_mpi / maximum_mpi (src/Utils.jl:688-730).
//
// IGG semantics reproduced: Cartesian grid of ranks `dims`, every rank holds arrays of the same local
// extents; per array and dimension ol = overlap + (size(A,d) − n_d) with overlap = 2; arrays with ol < 2 are not
// exchanged in that dimension; halo width 1; a rank sends plane `ol` (1-based) to its low neighbour and plane
// `size−ol+1` to its high neighbour and receives into planes 1 and `size`; the dimensions are processed x → y → z
// so ghost edges/corners carry the diagonal neighbours' values; physical boundaries have no neighbour.
//
// Implementation: NO per-dimension message rounds.  Because the sequence x → y → z only ever forwards values that
// were local to some rank before the exchange, the final value of every ghost element is the pre-exchange value of
// ONE element on ONE (possibly diagonal) neighbour — found by walking the dimensions in reverse order
// (jr_halo_chase).  So an exchange is
//     1. k_halo_pack : copy my six "send planes" (indices ol−1 and n−ol, whole planes incl. their ghost rows) of every
//                      array into my staging buffer (CUDA-IPC exported, double-buffered by epoch parity);
//     2. k_halo_pull : device-side flag barrier with all ranks (release/acquire at system scope on peer memory),
//                      then every ghost element is read straight from the owning peer's staging buffer over NVLink.
// Two small kernels per exchange, no host involvement, no NCCL call; bootstrap needs one host all-gather of the IPC
// handles, done through a callback the host language provides (MPI.Allgather in Julia, torch.distributed in Python).
#include "common.cuh"
#include "comm.cuh"

// ---------------------------------------------------------------------------------------------------------------
// host + device: where does ghost element `c` of an array with extents n[3] get its value from?
// returns true if the element is overwritten by the exchange; dr = rank-coordinate offset of the source rank,
// s = source index (0-based) there, first = first dimension walked (the staging plane the element sits in).
__host__ __device__ bool jr_halo_chase(const int n[3], const int ol[3], const bool has_lo[3], const bool has_hi[3],
                                       const int c[3], int dr[3], int s[3], int &first, int &first_side)
{
    bool moved = false;
    first = -1; first_side = 0;
    for (int d = 0; d < 3; d++) { dr[d] = 0; s[d] = c[d]; }
    for (int d = 2; d >= 0; d--) {
        if (ol[d] < 2) continue;
        if (c[d] == 0 && has_lo[d]) {
            dr[d] = -1; s[d] = n[d] - ol[d];          // low neighbour's plane size−ol+1 (1-based)
            if (!moved) { first = d; first_side = 1; }
            moved = true;
        } else if (c[d] == n[d] - 1 && has_hi[d]) {
            dr[d] = +1; s[d] = ol[d] - 1;             // high neighbour's plane ol (1-based)
            if (!moved) { first = d; first_side = 0; }
            moved = true;
        }
    }
    return moved;
}

static long stage_array_size(const int n[3]) { return 2 * (jr_stage_plane_size(n, 0) + jr_stage_plane_size(n, 1) + jr_stage_plane_size(n, 2)); }

__global__ void k_comm_barrier(const __grid_constant__ jr_comm_dev cd, unsigned long long epoch) { jr_comm_barrier_dev(cd, epoch); }

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t harr_idx(const jr_harr &A, const int c[3])
{
    return (size_t)(c[2] + A.o[2]) * A.sz + (size_t)(c[1] + A.o[1]) * A.sy + (size_t)(c[0] + A.o[0]);
}

// plane p = d·2 + side of array q: blockIdx.z = q·6 + p; (u, v) = the two free coordinates, fastest first
__device__ __forceinline__ bool plane_coords(const jr_harr &A, int d, int &u, int &v, int c[3])
{
    const int du = (d == 0) ? 1 : 0, dv = (d == 2) ? 1 : 2;
    u = blockIdx.x * blockDim.x + threadIdx.x;
    v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= A.n[du] || v >= A.n[dv]) return false;
    c[du] = u; c[dv] = v;
    return true;
}

__global__ void k_halo_pack(const __grid_constant__ HaloArgs h, double *__restrict__ stage)
{
    const int q = blockIdx.z / 6, p = blockIdx.z % 6, d = p >> 1, side = p & 1;
    const jr_harr &A = h.A[q];
    if (A.ol[d] < 2) return;
    // side 0 (plane ol−1) is read by the low neighbour, side 1 (plane n−ol) by the high neighbour
    if (side == 0 ? !h.has_lo[d] : !h.has_hi[d]) return;
    int u, v, c[3];
    if (!plane_coords(A, d, u, v, c)) return;
    c[d] = side == 0 ? A.ol[d] - 1 : A.n[d] - A.ol[d];
    stage[h.stage_off[q] + jr_stage_plane_off(A.n, d, side) + jr_stage_elem(A.n, d, c)] = A.p[harr_idx(A, c)];
}

__global__ void k_halo_pull(const __grid_constant__ HaloArgs h, const __grid_constant__ jr_comm_dev cd, unsigned long long epoch)
{
    jr_comm_barrier_dev(cd, epoch);
    const int q = blockIdx.z / 6, p = blockIdx.z % 6, d = p >> 1, side = p & 1;
    const jr_harr &A = h.A[q];
    if (A.ol[d] < 2) return;
    if (side == 0 ? !h.has_lo[d] : !h.has_hi[d]) return;
    int u, v, c[3];
    if (!plane_coords(A, d, u, v, c)) return;
    c[d] = side == 0 ? 0 : A.n[d] - 1;
    int dr[3], s[3], first, fside;
    if (!jr_halo_chase(A.n, A.ol, h.has_lo, h.has_hi, c, dr, s, first, fside)) return;
    const int peer = cd.nbr[(dr[2] + 1) * 9 + (dr[1] + 1) * 3 + (dr[0] + 1)];
    const double *src = cd.stage[peer][h.buf] + h.stage_off[q] + jr_stage_plane_off(A.n, first, fside) + jr_stage_elem(A.n, first, s);
    A.p[harr_idx(A, c)] = *src;
}

// ---------------------------------------------------------------------------------------------------------------
// z-ordered pull: the same exchange as k_halo_pull, but split along the slowest index.  A "head" launch (with the flag barrier, full
// grid) takes the first planes; a "rest" launch of a few CTAs on a second stream walks the remaining planes in z order and publishes
// its progress, so that a consumer which itself marches in z (the fused 3D-VA kernel of the next iteration) runs on top of it — the
// @hide_communication of the reference (Stokes3D.jl:104-121).  Sources are the peers' packed staging buffers: reading the x faces
// straight from the peers' arrays was measured at ≈ 20 µs per dependent access (every plane of a 255^3 set lies in another page of
// the peer mapping), the packed planes are a handful of pages.
struct ZArgs {
    HaloArgs h;
    int z0, z1, chunk;                   // planes [z0, z1) of the addressed space (index 2 + offset), published every `chunk` planes
    unsigned long long *prog;            // [gridDim.x] progress slots: prog_base + planes complete (nullptr: nothing published)
    unsigned long long prog_base;
};

__device__ __forceinline__ bool z_elem(const HaloArgs &h, const jr_comm_dev &cd, int q, const int c[3], const double *&src, double *&dst)
{
    const jr_harr &A = h.A[q];
    int dr[3], s[3], first, fside;
    if (!jr_halo_chase(A.n, A.ol, h.has_lo, h.has_hi, c, dr, s, first, fside)) return false;
    const int peer = cd.nbr[(dr[2] + 1) * 9 + (dr[1] + 1) * 3 + (dr[0] + 1)];
    src = cd.stage[peer][h.buf] + h.stage_off[q] + jr_stage_plane_off(A.n, first, fside) + jr_stage_elem(A.n, first, s);
    dst = A.p + harr_idx(A, c);
    return true;
}
// f(e, src, dst) → does element e exist; eight independent peer reads in flight per thread
template <class F>
__device__ __forceinline__ void z_batch(int total, int gt, int nthr, F f)
{
    for (int e0 = gt; e0 < total; e0 += 8 * nthr) {
        const double *src[8];
        double *dst[8];
        double v[8];
        bool ok[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + u * nthr;
            src[u] = nullptr; dst[u] = nullptr;
            ok[u] = e < total && f(e, src[u], dst[u]);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = ok[u] ? __ldcg(src[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (ok[u]) *dst[u] = v[u];
    }
}

__global__ void __launch_bounds__(256) k_halo_pull_z(const __grid_constant__ ZArgs z, const __grid_constant__ jr_comm_dev cd, unsigned long long epoch)
{
    const HaloArgs &h = z.h;
    if (epoch) jr_comm_barrier_dev(cd, epoch);   // every rank has packed (and nobody still reads the buffer packed two exchanges ago)
    const int nthr = gridDim.x * blockDim.x, gt = blockIdx.x * blockDim.x + threadIdx.x;
    // x / y faces with a neighbour, per array: only elements the exchange really moves are enumerated, so that every one of a thread's
    // eight slots is a peer read in flight (the launch is pure latency)
    int W = 0;
    for (int q = 0; q < h.narr; q++)
        for (int f = 0; f < 4; f++) {
            const int d = f >> 1;
            if (h.A[q].ol[d] >= 2 && ((f & 1) ? h.has_hi[d] : h.has_lo[d])) W += h.A[q].n[1 - d];
        }
    for (int zc = z.z0; zc < z.z1; zc += z.chunk) {
        const int ze = min(zc + z.chunk, z.z1);
        // (1) the x / y ghost faces of the planes of this chunk
        z_batch((ze - zc) * W, gt, nthr, [&](int e, const double *&src, double *&dst) {
            int w = e % W;
            const int Z = zc + e / W;
            for (int q = 0; q < h.narr; q++) {
                const jr_harr &A = h.A[q];
                for (int f = 0; f < 4; f++) {
                    const int d = f >> 1;
                    if (!(A.ol[d] >= 2 && ((f & 1) ? h.has_hi[d] : h.has_lo[d]))) continue;
                    const int L = A.n[1 - d];
                    if (w >= L) { w -= L; continue; }
                    int c[3];
                    c[2] = Z - A.o[2];
                    if (c[2] < 0 || c[2] >= A.n[2]) return false;
                    if (A.ol[2] >= 2 && ((c[2] == 0 && h.has_lo[2]) || (c[2] == A.n[2] - 1 && h.has_hi[2]))) return false;  // whole plane: part (2)
                    c[d] = (f & 1) ? A.n[d] - 1 : 0;
                    c[1 - d] = w;
                    return z_elem(h, cd, q, c, src, dst);
                }
            }
            return false;
        });
        // (2) whole ghost planes of the z faces that fall into this chunk
        for (int q = 0; q < h.narr; q++) {
            const jr_harr &A = h.A[q];
            if (A.ol[2] < 2) continue;
            for (int side = 0; side < 2; side++) {
                if (side == 0 ? !h.has_lo[2] : !h.has_hi[2]) continue;
                const int c2 = side == 0 ? 0 : A.n[2] - 1, Z = c2 + A.o[2];
                if (Z < zc || Z >= ze) continue;
                z_batch(A.n[0] * A.n[1], gt, nthr, [&](int e, const double *&src, double *&dst) {
                    const int c[3] = {e % A.n[0], e / A.n[0], c2};
                    return z_elem(h, cd, q, c, src, dst);
                });
            }
        }
        if (z.prog) {
            __syncthreads();
            if (threadIdx.x == 0)
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(z.prog + blockIdx.x), "l"(z.prog_base + (unsigned long long)ze) : "memory");
        }
    }
}

// deterministic all-reduce of n ≤ 16 doubles: every rank writes its partials into its slot on every rank, barrier,
// then every rank combines the slots in rank order (bit-identical result everywhere).  op: 0 sum, 1 max, 2 min
__global__ void k_allreduce(const __grid_constant__ jr_comm_dev cd, unsigned long long epoch, int buf, double *vals, int n, int op)
{
    const int t = threadIdx.x;
    if (t < n) {
        const double v = vals[t];
        for (int r = 0; r < cd.nranks; r++) cd.sig[r]->red[buf][cd.rank][t] = v;
    }
    __syncthreads();
    jr_comm_barrier_dev(cd, epoch);
    if (t < n) {
        const jr_comm_sig *me = cd.sig[cd.rank];
        double acc = me->red[buf][0][t];
        for (int r = 1; r < cd.nranks; r++) {
            const double v = me->red[buf][r][t];
            acc = op == 0 ? acc + v : op == 1 ? fmax(acc, v) : fmin(acc, v);
            if (op != 0 && v != v) acc = v;
        }
        vals[t] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
static int comm_allgather(jr_comm *cm, const void *send, void *recv, size_t bytes)
{
    JR_REQUIRE(cm->allgather, JR_ERR_ARG, "communicator has no all-gather callback");
    const int rc = cm->allgather(send, recv, bytes, cm->user);
    JR_REQUIRE(rc == 0, JR_ERR_NCCL, "host all-gather callback failed with code %d", rc);
    return JR_OK;
}

// export `ptr` (a cudaMalloc base pointer), all-gather the handles, open the peers' → out[r] (out[rank] = ptr)
static int comm_share(jr_comm *cm, void *ptr, std::vector<void *> &out)
{
    if (cm->nranks == 1) { out.assign(1, ptr); return JR_OK; }   // a periodic single rank is its own neighbour
    cudaIpcMemHandle_t mine;
    JR_CUDA(cudaIpcGetMemHandle(&mine, ptr));
    std::vector<cudaIpcMemHandle_t> all(cm->nranks);
    int st = comm_allgather(cm, &mine, all.data(), sizeof(mine));
    if (st) return st;
    out.assign(cm->nranks, nullptr);
    for (int r = 0; r < cm->nranks; r++) {
        if (r == cm->rank) { out[r] = ptr; continue; }
        std::string key((const char *)&all[r], sizeof(cudaIpcMemHandle_t));
        key += (char)r;
        auto it = cm->ipc_open.find(key);
        if (it == cm->ipc_open.end()) {
            void *p = nullptr;
            JR_CUDA(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess));
            it = cm->ipc_open.emplace(key, p).first;
        }
        out[r] = it->second;
    }
    return JR_OK;
}

static int comm_ensure_stage(jr_context *ctx, jr_comm *cm, size_t doubles)
{
    if (doubles <= cm->stage_cap) return JR_OK;
    // collective (all ranks hold arrays of the same extents, so all of them grow in the same call)
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    size_t cap = doubles + doubles / 4 + 1024;
    for (int b = 0; b < 2; b++) {
        // old buffers stay allocated: a slow peer may still be pulling from them
        if (cm->stage_mine[b]) cm->retired.push_back(cm->stage_mine[b]);
        JR_CUDA(cudaMalloc(&cm->stage_mine[b], cap * sizeof(double)));
        std::vector<void *> peers;
        int st = comm_share(cm, cm->stage_mine[b], peers);
        if (st) return st;
        for (int r = 0; r < cm->nranks; r++) cm->dev.stage[r][b] = (double *)peers[r];
    }
    cm->stage_cap = cap;
    return JR_OK;
}

int jr_comm_share(jr_context *ctx, void *mine, void **out)
{
    jr_comm *cm = ctx->comm;
    JR_REQUIRE(cm && cm->nranks > 1, JR_ERR_ARG, "no multi-rank communicator attached to this context");
    std::vector<void *> peers;
    int st = comm_share(cm, mine, peers);
    if (st) return st;
    for (int r = 0; r < cm->nranks; r++) out[r] = peers[r];
    return JR_OK;
}

int jr_comm_reserve_stage(jr_context *ctx, size_t doubles)
{
    jr_comm *cm = ctx->comm;
    JR_REQUIRE(cm && cm->active, JR_ERR_ARG, "no multi-rank communicator attached to this context");
    return comm_ensure_stage(ctx, cm, doubles);
}


static int halo_args(jr_context *ctx, jr_comm *cm, const jr_harr *arrs, int narr, HaloArgs &h, int &maxu, int &maxv)
{
    JR_REQUIRE(narr <= JR_HALO_MAX_ARRAYS, JR_ERR_ARG, "at most %d arrays per halo update", JR_HALO_MAX_ARRAYS);
    h.narr = narr;
    long off = 0;
    maxu = 1; maxv = 1;
    for (int q = 0; q < narr; q++) {
        h.A[q] = arrs[q];
        h.stage_off[q] = off;
        off += stage_array_size(arrs[q].n);
        for (int d = 0; d < 3; d++) {
            const int du = (d == 0) ? 1 : 0, dv = (d == 2) ? 1 : 2;
            if (arrs[q].n[du] > maxu) maxu = arrs[q].n[du];
            if (arrs[q].n[dv] > maxv) maxv = arrs[q].n[dv];
            JR_REQUIRE(arrs[q].ol[d] < 2 || arrs[q].n[d] >= 2 * arrs[q].ol[d], JR_ERR_SHAPE,
                       "array too small for its overlap in dimension %d (n = %d, ol = %d)", d, arrs[q].n[d], arrs[q].ol[d]);
        }
    }
    int st = comm_ensure_stage(ctx, cm, (size_t)off);
    if (st) return st;
    for (int d = 0; d < 3; d++) { h.has_lo[d] = cm->has_lo[d]; h.has_hi[d] = cm->has_hi[d]; }
    // staging buffers alternate strictly from one exchange to the next (a peer may still be pulling from the previous one)
    h.buf = (int)(cm->halo_count++ & 1);
    return JR_OK;
}

int jr_comm_halo(jr_context *ctx, const jr_harr *arrs, int narr)
{
    jr_comm *cm = ctx->comm;
    if (!cm || !cm->active || narr == 0) return JR_OK;
    HaloArgs h;
    int maxu, maxv;
    int st = halo_args(ctx, cm, arrs, narr, h, maxu, maxv);
    if (st) return st;
    const unsigned long long epoch = ++cm->epoch;
    dim3 block(32, 8, 1), grid((maxu + 31) / 32, (maxv + 7) / 8, narr * 6);
    k_halo_pack<<<grid, block, 0, ctx->stream>>>(h, (double *)cm->stage_mine[h.buf]);
    k_halo_pull<<<grid, block, 0, ctx->stream>>>(h, cm->dev, epoch);
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

// update_halo!(arrs...) in two parts, for a caller whose own kernel packs the send planes (k_bc_box3 of the fused 3D-VA iteration: one
// launch less per exchange): begin → layout + this rank's staging buffer; pull → flag barrier + pull
int jr_comm_halo_begin(jr_context *ctx, const jr_harr *arrs, int narr, HaloArgs *h, double **stage_mine)
{
    jr_comm *cm = ctx->comm;
    JR_REQUIRE(cm && cm->active && narr >= 1, JR_ERR_ARG, "split exchange without a multi-rank / periodic communicator");
    int maxu, maxv;
    int st = halo_args(ctx, cm, arrs, narr, *h, maxu, maxv);
    if (st) return st;
    *stage_mine = (double *)cm->stage_mine[h->buf];
    return JR_OK;
}
int jr_comm_halo_pull(jr_context *ctx, const HaloArgs *h)
{
    jr_comm *cm = ctx->comm;
    int maxu = 1, maxv = 1;
    for (int q = 0; q < h->narr; q++)
        for (int d = 0; d < 3; d++) {
            const int du = (d == 0) ? 1 : 0, dv = (d == 2) ? 1 : 2;
            if (h->A[q].n[du] > maxu) maxu = h->A[q].n[du];
            if (h->A[q].n[dv] > maxv) maxv = h->A[q].n[dv];
        }
    const unsigned long long epoch = ++cm->epoch;
    dim3 block(32, 8, 1), grid((maxu + 31) / 32, (maxv + 7) / 8, h->narr * 6);
    k_halo_pull<<<grid, block, 0, ctx->stream>>>(*h, cm->dev, epoch);
    ctx->launches += 1;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

// update_halo!(arrs...) split along the slowest index of the addressed space (pz planes): pack + flag barrier + planes [0, head) on
// ctx->stream (ev_head recorded), planes [head, pz) by `rest_ctas` CTAs on `side`, which publish prog_base + planes complete into
// prog[0 … rest_ctas) every `chunk` planes (ev_rest recorded).  The caller joins ev_rest before the next exchange.
int jr_comm_halo_z(jr_context *ctx, const jr_harr *arrs, int narr, int pz, int head, int chunk, int rest_ctas, cudaStream_t side, cudaEvent_t ev_head,
                   cudaEvent_t ev_rest, unsigned long long *prog, unsigned long long prog_base)
{
    jr_comm *cm = ctx->comm;
    JR_REQUIRE(cm && cm->active && narr >= 1, JR_ERR_ARG, "z-ordered exchange without a multi-rank / periodic communicator");
    ZArgs z;
    memset(&z, 0, sizeof(z));
    int maxu, maxv;
    int st = halo_args(ctx, cm, arrs, narr, z.h, maxu, maxv);
    if (st) return st;
    const unsigned long long epoch = ++cm->epoch;
    dim3 block(32, 8, 1), grid((maxu + 31) / 32, (maxv + 7) / 8, narr * 6);
    k_halo_pack<<<grid, block, 0, ctx->stream>>>(z.h, (double *)cm->stage_mine[z.h.buf]);
    if (head > pz) head = pz;
    z.z0 = 0; z.z1 = head; z.chunk = head;
    z.prog = nullptr; z.prog_base = 0;
    k_halo_pull_z<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(z, cm->dev, epoch);
    ctx->launches += 2;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaEventRecord(ev_head, ctx->stream));
    JR_CUDA(cudaStreamWaitEvent(side, ev_head, 0));
    z.z0 = head; z.z1 = pz; z.chunk = chunk < 1 ? 1 : chunk;
    z.prog = prog; z.prog_base = prog_base;
    k_halo_pull_z<<<rest_ctas, 256, 0, side>>>(z, cm->dev, 0ull);
    ctx->launches += 1;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaEventRecord(ev_rest, side));
    return JR_OK;
}

// bytes a rank receives in one exchange of `arrs` (reporting: halo bytes per iteration vs NVLink bandwidth)
size_t jr_comm_halo_bytes(const jr_comm *cm, const jr_harr *arrs, int narr)
{
    if (!cm || !cm->active) return 0;
    size_t b = 0;
    for (int q = 0; q < narr; q++)
        for (int d = 0; d < 3; d++) {
            if (arrs[q].ol[d] < 2) continue;
            const int nb = (int)cm->has_lo[d] + (int)cm->has_hi[d];
            b += (size_t)nb * jr_stage_plane_size(arrs[q].n, d) * sizeof(double);
        }
    return b;
}

jr_harr jr_harr_dense(double *p, const int32_t ext[3], const int32_t ncell[3])
{
    jr_harr A;
    A.p = p; A.sy = ext[0]; A.sz = (long)ext[0] * ext[1];
    for (int d = 0; d < 3; d++) { A.n[d] = ext[d]; A.o[d] = 0; A.ol[d] = ext[d] > 1 ? 2 + (ext[d] - ncell[d]) : 0; }
    return A;
}

int jr_comm_allreduce_dev(jr_context *ctx, double *d_vals, int n, int op)
{
    jr_comm *cm = ctx->comm;
    if (!cm || cm->nranks == 1) return JR_OK;
    JR_REQUIRE(n >= 1 && n <= JR_COMM_RED_SLOTS, JR_ERR_ARG, "all-reduce of %d values (max %d)", n, JR_COMM_RED_SLOTS);
    const unsigned long long epoch = ++cm->epoch;
    const int buf = (int)(cm->red_count++ & 1);
    k_allreduce<<<1, 64, 0, ctx->stream>>>(cm->dev, epoch, buf, d_vals, n, op);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

extern "C" {

int jr_comm_create_periodic(jr_context *ctx, int rank, int nranks, const int32_t dims[3], const int32_t coords[3], const int32_t periods[3],
                            jr_allgather_fn allgather, void *user, jr_comm **out)
{
    JR_REQUIRE(ctx && out && dims && coords, JR_ERR_ARG, "jr_comm_create: null argument");
    JR_REQUIRE(nranks >= 1 && nranks <= JR_COMM_MAX_RANKS && rank >= 0 && rank < nranks, JR_ERR_ARG,
               "jr_comm_create: rank %d of %d (max %d ranks: one NVSwitch domain)", rank, nranks, JR_COMM_MAX_RANKS);
    JR_REQUIRE((long)dims[0] * dims[1] * dims[2] == nranks, JR_ERR_ARG, "jr_comm_create: dims %dx%dx%d != %d ranks", dims[0], dims[1],
               dims[2], nranks);
    for (int d = 0; d < 3; d++)
        JR_REQUIRE(coords[d] >= 0 && coords[d] < dims[d], JR_ERR_ARG, "jr_comm_create: coordinate %d out of range in dimension %d", coords[d], d);
    JR_REQUIRE(nranks == 1 || allgather, JR_ERR_ARG, "jr_comm_create: an all-gather callback is required for more than one rank");
    JR_CUDA(cudaSetDevice(ctx->device));
    jr_comm *cm = new jr_comm();
    cm->rank = rank; cm->nranks = nranks; cm->allgather = allgather; cm->user = user;
    cm->active = nranks > 1;
    for (int d = 0; d < 3; d++) {
        cm->dims[d] = dims[d]; cm->coords[d] = coords[d];
        cm->periods[d] = periods && periods[d] ? 1 : 0;
        // ImplicitGlobalGrid: a periodic dimension wraps around (with one rank in it the rank is its own neighbour)
        cm->has_lo[d] = coords[d] > 0 || cm->periods[d];
        cm->has_hi[d] = coords[d] < dims[d] - 1 || cm->periods[d];
        if (cm->periods[d]) cm->active = true;
    }
    memset(&cm->dev, 0, sizeof(cm->dev));
    cm->dev.rank = rank; cm->dev.nranks = nranks;
    for (int q = 0; q < 27; q++) cm->dev.nbr[q] = -1;
    if (cm->active) {
        // rank of every coordinate triple
        std::vector<int32_t> allc(3 * nranks);
        int st = JR_OK;
        if (nranks > 1) st = comm_allgather(cm, coords, allc.data(), 3 * sizeof(int32_t));
        else for (int d = 0; d < 3; d++) allc[d] = coords[d];
        if (st) { delete cm; return st; }
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int c[3] = {coords[0] + dx, coords[1] + dy, coords[2] + dz};
                    bool exists = true;
                    for (int d = 0; d < 3; d++) {
                        if (c[d] >= 0 && c[d] < dims[d]) continue;
                        if (cm->periods[d]) c[d] = (c[d] + dims[d]) % dims[d];
                        else exists = false;
                    }
                    if (!exists) continue;
                    for (int r = 0; r < nranks; r++)
                        if (allc[3 * r] == c[0] && allc[3 * r + 1] == c[1] && allc[3 * r + 2] == c[2]) cm->dev.nbr[(dz + 1) * 9 + (dy + 1) * 3 + dx + 1] = r;
                }
        JR_REQUIRE(cm->dev.nbr[13] == rank, JR_ERR_ARG, "jr_comm_create: coordinates gathered from the ranks are inconsistent");
        // signal page
        JR_CUDA(cudaMalloc((void **)&cm->sig_mine, sizeof(jr_comm_sig)));
        JR_CUDA(cudaMemset(cm->sig_mine, 0, sizeof(jr_comm_sig)));
        JR_CUDA(cudaDeviceSynchronize());
        std::vector<void *> peers;
        st = comm_share(cm, cm->sig_mine, peers);
        if (st) { delete cm; return st; }
        for (int r = 0; r < nranks; r++) cm->dev.sig[r] = (jr_comm_sig *)peers[r];
        if (nranks > 1) {
            // everybody has zeroed its page before anybody signals: one more host collective as a barrier
            int32_t token = rank;
            std::vector<int32_t> toks(nranks);
            if ((st = comm_allgather(cm, &token, toks.data(), sizeof(token)))) { delete cm; return st; }
        }
    }
    *out = cm;
    return JR_OK;
}

int jr_comm_create(jr_context *ctx, int rank, int nranks, const int32_t dims[3], const int32_t coords[3], jr_allgather_fn allgather,
                   void *user, jr_comm **out)
{
    return jr_comm_create_periodic(ctx, rank, nranks, dims, coords, nullptr, allgather, user, out);
}

int jr_comm_destroy(jr_comm *cm)
{
    if (!cm) return JR_OK;
    cudaDeviceSynchronize();
    // a slower peer's last pull / all-reduce may still be reading this rank's exported buffers: every rank drains its own
    // device first, then one host token round makes sure ALL ranks have done so before anybody unmaps or frees
    if (cm->nranks > 1 && cm->allgather) {
        char tok = 1;
        std::vector<char> all((size_t)cm->nranks);
        cm->allgather(&tok, all.data(), 1, cm->user);
    }
    for (auto &kv : cm->ipc_open) cudaIpcCloseMemHandle(kv.second);
    for (void *p : cm->retired) cudaFree(p);
    for (int b = 0; b < 2; b++)
        if (cm->stage_mine[b]) cudaFree(cm->stage_mine[b]);
    if (cm->sig_mine) cudaFree(cm->sig_mine);
    delete cm;
    return JR_OK;
}

int jr_context_set_comm(jr_context *ctx, jr_comm *comm)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    ctx->comm = comm;
    return JR_OK;
}

int jr_comm_barrier(jr_context *ctx)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    jr_comm *cm = ctx->comm;
    if (!cm || cm->nranks == 1) return JR_OK;
    k_comm_barrier<<<1, 64, 0, ctx->stream>>>(cm->dev, ++cm->epoch);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_update_halo3d(jr_context *ctx, int narrays, double *const *arrays, const int32_t *extents, const int32_t ncell[3])
{
    JR_REQUIRE(ctx && arrays && extents && ncell, JR_ERR_ARG, "jr_update_halo3d: null argument");
    JR_REQUIRE(narrays >= 0 && narrays <= JR_HALO_MAX_ARRAYS, JR_ERR_ARG, "jr_update_halo3d: at most %d arrays per call", JR_HALO_MAX_ARRAYS);
    jr_harr A[JR_HALO_MAX_ARRAYS];
    for (int q = 0; q < narrays; q++) {
        JR_REQUIRE(arrays[q], JR_ERR_ARG, "jr_update_halo3d: array %d is NULL", q);
        A[q] = jr_harr_dense(arrays[q], extents + 3 * q, ncell);
    }
    int st = jr_comm_halo(ctx, A, narrays);
    if (st) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_allreduce_f64(jr_context *ctx, double *vals_host, int n, int op)
{
    JR_REQUIRE(ctx && vals_host, JR_ERR_ARG, "jr_allreduce_f64: null argument");
    JR_REQUIRE(n >= 1 && n <= JR_COMM_RED_SLOTS && op >= 0 && op <= 2, JR_ERR_ARG, "jr_allreduce_f64: n = %d (max %d), op = %d", n,
               JR_COMM_RED_SLOTS, op);
    jr_comm *cm = ctx->comm;
    if (!cm || cm->nranks == 1) return JR_OK;
    void *slot = nullptr;
    int st = jr_ctx_scratch(ctx, "allreduce_slots", JR_COMM_RED_SLOTS * sizeof(double), &slot);
    if (st) return st;
    JR_CUDA(cudaMemcpyAsync(slot, vals_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if ((st = jr_comm_allreduce_dev(ctx, (double *)slot, n, op))) return st;
    JR_CUDA(cudaMemcpyAsync(vals_host, slot, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

// pure host arithmetic (no GPU): source of element idx of an array after update_halo! — the index logic of
// k_halo_pull, exported so that it can be tested against a literal x → y → z message exchange on CPU ranks.
int jr_halo_source_periodic(const int32_t dims[3], const int32_t periods[3], const int32_t coords[3], const int32_t ext[3], const int32_t ncell[3],
                            const int32_t idx[3], int32_t src_coords[3], int32_t src_idx[3])
{
    JR_REQUIRE(dims && coords && ext && ncell && idx && src_coords && src_idx, JR_ERR_ARG, "jr_halo_source: null argument");
    int n[3], ol[3], c[3], dr[3], s[3], first, fside;
    bool lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
        n[d] = ext[d]; c[d] = idx[d];
        ol[d] = ext[d] > 1 ? 2 + (ext[d] - ncell[d]) : 0;
        const bool per = periods && periods[d];
        lo[d] = coords[d] > 0 || per; hi[d] = coords[d] < dims[d] - 1 || per;
    }
    const bool moved = jr_halo_chase(n, ol, lo, hi, c, dr, s, first, fside);
    for (int d = 0; d < 3; d++) { src_coords[d] = ((coords[d] + dr[d]) % dims[d] + dims[d]) % dims[d]; src_idx[d] = s[d]; }
    return moved ? 1 : 0;
}
int jr_halo_source(const int32_t dims[3], const int32_t coords[3], const int32_t ext[3], const int32_t ncell[3], const int32_t idx[3],
                   int32_t src_coords[3], int32_t src_idx[3])
{
    return jr_halo_source_periodic(dims, nullptr, coords, ext, ncell, idx, src_coords, src_idx);
}

} // extern "C"
