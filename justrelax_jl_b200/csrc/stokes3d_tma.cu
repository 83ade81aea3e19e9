// stokes3d_tma.cu — the fused sm_100a kernel of the 3D visco-elastic Stokes PT iteration (variant 3D-VA):
// ONE kernel per PT iteration, inputs staged by TMA (cp.async.bulk.tensor + mbarrier ring), 2.5D z-marching.
//
// What it replaces (reference, one @parallel launch each, ≈75 array passes = 600 B/cell/iteration):
//   compute_∇V! → compute_P! → compute_strain_rate! → compute_τ! → compute_V! → velocity2displacement!
//   (src/stokes/Stokes3D.jl:78-119; kernels VelocityKernels.jl:3-6,59-104,182-242, PressureKernels.jl:10-15,
//   StressKernels.jl:149-230) and flow_bcs! (BoundaryConditions.jl:86-99).
//
// Data layout in HBM ("box sets", owned by the library for the duration of a solve):
//   every staggered array of the iteration is stored in ONE common index space, the box
//   (X,Y,Z) ∈ [0,PX)×[0,PY)×[0,PZ), PX = nx+2 rounded to 32 B, PY = ny+2, PZ = nz+2, with
//        cell centre (i,j,k)      → (i+1, j+1, k+1)      P, τxx, τyy, τzz, η, ητ, ρg, K, G, P0, Q
//        Vx(I, J, K) (I face)     → (I+1, J,   K  )      J,K ghosted indices of the reference
//        Vy(I, J, K)              → (I,   J+1, K  )
//        Vz(I, J, K)              → (I,   J,   K+1)
//        τxy(I, J, k) / τxz(I, j, K) / τyz(i, J, K) (edges) → (I+1, J+1, k+1) etc. (edge = low corner of its cell)
//   so ONE offset addresses all 25 arrays at a thread's position.  Arrays of a set are interleaved plane-wise:
//   element (a, X, Y, Z) of a set with NA arrays lives at ((Z·NA + a)·PY + Y)·PX + X — a 4-D tensor that one
//   CUtensorMap describes, with 16-B aligned pitches whatever nx is (255, 511 … are odd).  η and G carry their
//   clamped-index ghost copies in the box, everything else is zero outside its own extent.
//   Sets: S0/S1 (state ping-pong: Vx,Vy,Vz,P,τxx,τyy,τzz,τyz,τxz,τxy), C (η,ητ,ρgx,ρgy,ρgz), D (finite dt only:
//   G,K,P0,Q,τ_o×6).  The user's dense arrays are packed at solve entry and unpacked at exit.
//
// Kernel: a CTA owns a (30 × BY−2) column tile (+1 halo ring = 32 × BY threads, a warp = one x-row) and marches a
// z-chunk.  Per z-step ONE elected thread issues 15 (25) TMA box loads (32 × BY × 1 doubles each) into the next
// slot of a 3-deep shared-memory ring; completion arrives on the slot's mbarrier.  Threads read their own and
// their neighbours' values from the slot, keep the z-direction state in a register queue, publish the new
// stresses IN PLACE in the slot (one __syncthreads per step), then form the momentum residuals and the new
// velocities and store the 10 outputs straight from registers (coalesced 240-B rows).  No global load
// instruction, no address arithmetic and no bounds predicate is left on the load side.
// Jacobi-exact (in → out sets), arithmetic identical operation for operation to the reference
// (same fma placement, same summation order) ⇒ bit-comparable with the CPU oracle.
#include "common.cuh"
#include "tma.cuh"
#include "comm.cuh"

// set orders: arrays that are loaded at the same box plane are adjacent, so ONE TMA box (32 × BY × n × 1) brings n tiles
enum { S_tzz = 0, S_P, S_txx, S_tyy, S_txy, /* plane k+1 */ S_Vx, S_Vy, S_Vz, S_tyz, S_txz, /* plane k+2 */ S_N };
enum { C_eta = 0, /* plane k+2 */ C_ett, C_fx, C_fy, C_fz, /* plane k+1 */ C_N };
enum { D_G = 0, D_oyz, D_oxz, /* plane k+2 */ D_K, D_P0, D_Q, D_oxx, D_oyy, D_ozz, D_oxy, /* plane k+1 */ D_N };
// tiles of one ring slot.  The first (τzz) and the last (ρgz) tile are only ever read at a thread's own position:
// the ±1 / ±32 / −33 neighbour reads of rim lanes then stay inside the slot without any index clamping.
enum { T_tzz = 0, T_P, T_txx, T_tyy, T_txy, T_Vx, T_Vy, T_Vz, T_tyz, T_txz, T_eta, T_NEXT };
// RHOG = false: the three body-force arrays are spatially constant (ρg ≡ 0 in SolVi, gravity-free benchmarks) and
// are not streamed at all — their value travels as a kernel argument.
template <bool FIN, bool RHOG> struct SlotMap {
    static constexpr int G = T_NEXT, oyz = G + 1, oxz = G + 2, K = G + 3, P0 = G + 4, Q = G + 5, oxx = G + 6, oyy = G + 7, ozz = G + 8,
                         oxy = G + 9;
    static constexpr int ett = FIN ? G + 10 : T_NEXT, fx = ett + 1, fy = ett + 2, fz = ett + 3, NARR = ett + (RHOG ? 4 : 1);
};

// push protocol (see comm.cuh: push_flags).  Offsets index the 27 neighbour directions [(dz+1)·9 + (dy+1)·3 + dx+1].
struct PushArgs {
    double *peer_out[27];              // neighbours' out-set base (the set with the same parity as ours); nullptr = no such neighbour
    jr_comm_sig *sig_peer[27];         // their signal pages (nullptr = none) and ours
    jr_comm_sig *sig_mine;
    int nbr_rank[27];
    int rank;
    int has_lo[3], has_hi[3];
    long delta[3];                     // (n_d − 2) · stride_d in elements: own element + Σ −f_d·delta_d = its ghost image on the neighbour
    unsigned long long epoch;          // launch number of this iteration (monotonic over the communicator's life)
};

// store v, produced at out-set offset `off` of a V component whose normal dimension is `nrm`, into every neighbour whose ghost planes
// hold it: g = (gi, gj, k) cell / face indices; send planes are index 2 on the low side and n − 2 (normal) / n − 3 (tangential) on
// the high side — IGG's planes ol and size − ol + 1 of the staggered array (overlap 2)
__device__ __forceinline__ void jr_push_v(const PushArgs &pu, int nrm, int gi, int gj, int k, int nx, int ny, int nz, size_t off, double v)
{
    const int f0 = (gi == 2 && pu.has_lo[0]) ? -1 : ((gi == (nrm == 0 ? nx - 2 : nx - 3) && pu.has_hi[0]) ? 1 : 0);
    const int f1 = (gj == 2 && pu.has_lo[1]) ? -1 : ((gj == (nrm == 1 ? ny - 2 : ny - 3) && pu.has_hi[1]) ? 1 : 0);
    const int f2 = (k == 2 && pu.has_lo[2]) ? -1 : ((k == (nrm == 2 ? nz - 2 : nz - 3) && pu.has_hi[2]) ? 1 : 0);
    if (!(f0 | f1 | f2)) return;
#pragma unroll
    for (int m = 1; m < 8; m++) {
        const int d0 = (m & 1) ? f0 : 0, d1 = (m & 2) ? f1 : 0, d2 = (m & 4) ? f2 : 0;
        if (((m & 1) && !f0) || ((m & 2) && !f1) || ((m & 4) && !f2)) continue;
        double *base = pu.peer_out[(d2 + 1) * 9 + (d1 + 1) * 3 + d0 + 1];
        if (base) base[(long)off - d0 * pu.delta[0] - d1 * pu.delta[1] - d2 * pu.delta[2]] = v;
    }
}

// TRAIL: flow_bcs! inside the iteration launch, OFF the critical path.  The persistent grid gets `ntrail` extra CTAs (the CTA slots the
// column tiles leave free: 296 − 288 at 255^3) that trail the z-march: main CTA c publishes the number of z-steps whose stores are
// complete in done[c]; the trailing CTAs poll the minimum and fill, plane group by plane group, the "ring" of every velocity plane
// (elements whose x or y index is a ghost / boundary-normal one) with the same gather as k_bc_box3.  The interior of the two z ghost
// planes and of the z boundary-normal faces is written by the thread that holds its source, in the first / last z-steps (CTA-uniform
// branches).  One launch per iteration instead of two, and the boundary work overlaps the z-march.
struct TrailArgs {
    unsigned long long *done;        // [main CTAs] done_base + complete z-steps of this launch
    unsigned long long done_base;
    const double *in;                // in-set base (generic loads of values no boundary condition touches)
    int ntrail;                      // trailing CTAs = gridDim.x − main CTAs
    int nb, nb_tail, tail_planes;    // plane groups per batch far from / within tail_planes of the end of the z-march
    int sig_every;                   // main CTAs publish every sig_every-th step
    int fs_lo[3], fs_hi[3], ns_lo[3], ns_hi[3];  // per dimension and side: free-slip / no-slip active (k_bc_box3's flags)
};

struct alignas(64) VaArgs {
    CUtensorMap mS5, mC1, mC4, mD1, mD2, mD7;  // in-state set (5-array boxes), const set, finite-dt set
    double *out;                               // out-state set base
    double *divV, *RP, *exx, *eyy, *ezz, *eyz, *exz, *exy, *Rx, *Ry, *Rz, *Ux, *Uy, *Uz;  // dense user arrays (DIAG)
    double *dVx, *dVy, *dVz, *dP, *dtxx, *dtyy, *dtzz, *dtyz, *dtxz, *dtxy;  // dense state (DIAG: an observable iteration leaves
                                                                            // the user's arrays current, no unpack pass)
    unsigned long long *progress;              // grid-wide step counter (soft lock-step of the resident CTAs)
    unsigned long long progress_base;          // counter value at launch
    int nx, ny, nz, PX, PY, kchunk;
    int ntx, nty, nchunk;                      // work items: ntx · nty column tiles × nchunk z-chunks
    int slack;                                 // a CTA may run at most DEPTH + slack z-steps ahead of the slowest one
    int stagger_ns;                            // >= 0: the compute-plane half of a step's loads is issued by the northern halo warp
                                               // this many ns after the barrier (smooths the lock-stepped request bursts); < 0: off
    int pol_ld, pol_st;  // L2 eviction policy of the TMA loads / output stores (0 normal, 1 evict_first, 2 evict_last)
    double _dx, _dy, _dz, dt, r, theta_dtau, eta_dtau;
    double fxc, fyc, fzc;  // constant body force (RHOG = false)
    // ---- MULTI (several PT iterations in one launch, boundary conditions applied by the kernel itself) ----
    CUtensorMap mS5b;                  // the other state set (in-set of odd iterations of this launch)
    double *out2;                      // out-set of odd iterations (= the set mS5 describes)
    unsigned long long *gbar;          // grid barrier counter between iterations
    unsigned long long gbar_base;      // its value at launch
    int niter;                         // iterations of this launch
    int bc_nsn[6];                     // x-lo, x-hi, y-lo, y-hi, z-lo, z-hi: boundary-normal face is zeroed (no_slip!)
    double bc_sg[6];                   // same sides: tangential ghost = bc_sg · interior (+1 free slip, −1 no slip)
    // ---- PUSH (multi-GPU): the halo exchange of update_halo!(Vx, Vy, Vz) (Stokes3D.jl:120) inside the iteration.  Every V element
    // on one of this rank's send planes is stored a second (… eighth) time straight into the ghost planes of the neighbours'
    // out-sets over the CUDA-IPC mapping, as it is produced — the transfer rides on NVLink under the z-march.
    PushArgs push;
    TrailArgs tr;
    // ---- overlapped halo exchange (multi-GPU): the direct exchange of the PREVIOUS iteration's velocities (comm.cu: k_halo_direct on a
    // second stream) walks the planes of this launch's in-set in z order while this launch marches behind it: the producer lane waits
    // until the arrival plane it is about to load has been exchanged
    const unsigned long long *halo_prog;  // progress slots of the exchange in flight (nullptr: the in-set is complete)
    unsigned long long halo_base;         // slot value = halo_base + planes complete
    int halo_nslots, halo_head, halo_pz;  // planes < halo_head were exchanged before this launch; halo_pz = planes of the set
    int halo_faces;                       // bit 2d + side: this rank has a neighbour there (only tiles on such a face wait)
    int spin_cap;                         // > 0: polls after which the soft lock-step gives up (non-cooperative launches)
};

#define TXW 30  // owned columns per tile
#ifndef JR_VA_MINB8
#define JR_VA_MINB8 3  // dt = Inf, 8-row tiles: 80 registers, three CTAs per SM (12 tiles x 2 KB x 3 stages = 72 KB each)
#endif

__device__ __forceinline__ bool jr_elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}


// ---- the trailing CTAs of a TRAIL launch (see TrailArgs) -------------------------------------------------------------------------
// ring of plane K of velocity component q (extents n0 × n1 in x, y): the two x-rows first (coalesced), then the column pairs
__device__ __forceinline__ void jr_ring_decode(int r, int n0, int n1, int &c0, int &c1)
{
    if (r < n0) { c0 = r; c1 = 0; }
    else if (r < 2 * n0) { c0 = r - n0; c1 = n1 - 1; }
    else { const int t = r - 2 * n0; c1 = 1 + (t >> 1); c0 = (t & 1) ? n0 - 1 : 0; }
}
// flow_bcs! value of element c of component q: the gather of k_bc_box3 (no_slip! → free_slip! as complete sweeps; every ghost value from
// its fully clamped source with the product of the per-dimension signs).  Returns the out-set offsets of the element and of its source.
__device__ __forceinline__ void jr_trail_gather(const VaArgs &a, int q, const int c[3], size_t &ic, size_t &is, double &sign, bool &zero,
                                                bool &computed)
{
    const int nc[3] = {a.nx, a.ny, a.nz};
    int n[3], s[3];
#pragma unroll
    for (int e = 0; e < 3; e++) { n[e] = nc[e] + (e == q ? 1 : 2); s[e] = c[e]; }
    sign = 1.0; zero = false;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const bool lo_e = c[e] == 0, hi_e = c[e] == n[e] - 1;
        if (!lo_e && !hi_e) continue;
        const bool fsl = lo_e ? a.tr.fs_lo[e] : a.tr.fs_hi[e], nsl = lo_e ? a.tr.ns_lo[e] : a.tr.ns_hi[e];
        if (e == q) {
            if (nsl) zero = true;
        } else if (fsl || nsl) {
            s[e] = lo_e ? 1 : n[e] - 2;
            if (!fsl) sign = -sign;
        }
    }
    computed = true;
#pragma unroll
    for (int e = 0; e < 3; e++) computed = computed && s[e] >= 1 && s[e] <= n[e] - 2;
    const size_t pxy = (size_t)a.PX * a.PY;
    const int ox = q == 0, oy = q == 1, oz = q == 2;
    ic = ((size_t)(c[2] + oz) * S_N + (S_Vx + q)) * pxy + (size_t)(c[1] + oy) * a.PX + (c[0] + ox);
    is = ((size_t)(s[2] + oz) * S_N + (S_Vx + q)) * pxy + (size_t)(s[1] + oy) * a.PX + (s[0] + ox);
}

#define TRAIL_UNROLL 8
// Latency is the trailing CTAs' only cost (a poll and two dependent memory round trips per batch, on a memory system the z-march keeps
// saturated), so a thread owns ring POSITIONS and takes all planes of a batch at once: TRAIL_UNROLL independent gathers in flight.
// Batches are a.tr.nb plane groups while the z-march is far from its end and a.tr.nb_tail near it (the last batch — the groups
// that wait for the final z-step — is what the launch pays on top of the z-march).
template <int NTHREADS>
__device__ __noinline__ void jr_va_trailer(const VaArgs &a, int t, int tid)
{
    __shared__ unsigned long long s_min[NTHREADS / 32];
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int Gm = (int)gridDim.x - a.tr.ntrail;
    const int R0 = 2 * (nx + 1) + 2 * ny, R1 = 2 * (nx + 2) + 2 * (ny - 1), R2 = 2 * (nx + 2) + 2 * ny;
    const int per = R0 + R1 + R2;
    const int nZ = nz + 2;  // plane groups: box planes Z = 0 … nz+1 (Vx, Vy: K = Z; Vz: K = Z − 1)
    double *const outset = a.out;
    const int stride = a.tr.ntrail * NTHREADS;
    for (int z0 = 0; z0 < nZ;) {
        // groups nz and nz+1 gather from plane nz (the final z-step): they form the last batch
        int z1 = z0 + (z0 + a.tr.nb + a.tr.tail_planes <= nZ ? a.tr.nb : a.tr.nb_tail);
        if (z1 > nZ - 2 && z0 < nZ - 2) z1 = nZ - 2;
        if (z1 > nZ) z1 = nZ;
        // box plane Z is stored in z-step Z + 1 (Z = 1 … nz); group 0 gathers from plane 1, group nz+1 from plane nz
        const int zs = min(max(z1 - 1, 1), nz);
        const unsigned long long need = a.tr.done_base + (unsigned long long)(zs + 2);
        for (;;) {
            unsigned long long m = ~0ull;
            for (int c = tid; c < Gm; c += NTHREADS) {
                unsigned long long v;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.tr.done + c) : "memory");
                m = v < m ? v : m;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long w = __shfl_xor_sync(0xffffffffu, m, o);
                m = w < m ? w : m;
            }
            if ((tid & 31) == 0) s_min[tid >> 5] = m;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < NTHREADS / 32; w++) m = s_min[w] < m ? s_min[w] : m;
            __syncthreads();
            if (m >= need) break;
            __nanosleep(100);
        }
        // (ld.acquire above + the CTA barriers: the main CTAs' stores behind the published step counts are visible now)
        for (int r0 = t * NTHREADS + tid; r0 < per; r0 += stride) {
            int q, c[3];
            if (r0 < R0) { q = 0; jr_ring_decode(r0, nx + 1, ny + 2, c[0], c[1]); }
            else if (r0 < R0 + R1) { q = 1; jr_ring_decode(r0 - R0, nx + 2, ny + 1, c[0], c[1]); }
            else { q = 2; jr_ring_decode(r0 - R0 - R1, nx + 2, ny + 2, c[0], c[1]); }
            const int oz = q == 2 ? 1 : 0;
            for (int zb = z0; zb < z1; zb += TRAIL_UNROLL) {
                size_t ic[TRAIL_UNROLL], is[TRAIL_UNROLL];
                double sg[TRAIL_UNROLL], val[TRAIL_UNROLL];
                bool ok[TRAIL_UNROLL], zero[TRAIL_UNROLL], comp[TRAIL_UNROLL];
#pragma unroll
                for (int u = 0; u < TRAIL_UNROLL; u++) {
                    c[2] = zb + u - oz;
                    ok[u] = zb + u < z1 && c[2] >= 0;  // Vz has no plane below box plane 1
                    ic[u] = is[u] = 0; sg[u] = 1.0; zero[u] = true; comp[u] = false;
                    if (ok[u]) jr_trail_gather(a, q, c, ic[u], is[u], sg[u], zero[u], comp[u]);
                }
#pragma unroll
                for (int u = 0; u < TRAIL_UNROLL; u++) {
                    val[u] = 0.0;
                    if (ok[u] && !zero[u]) val[u] = comp[u] ? __ldcg(outset + is[u]) : __ldg(a.tr.in + is[u]);
                }
#pragma unroll
                for (int u = 0; u < TRAIL_UNROLL; u++)
                    if (ok[u]) outset[ic[u]] = zero[u] ? 0.0 : sg[u] * val[u];
            }
        }
        z0 = z1;
    }
}

// Persistent kernel.  grid = min(resident CTA slots, work items), launched cooperatively so that every CTA is
// resident.  Work item = (column tile, z-chunk), numbered z-chunk-major / y / x-fastest; CTA c takes items
// c, c+G, c+2G, …; every item is kchunk+2 z-steps (step 0 fills the register queue, step 1 is the warm-up plane).
//   producer = one elected lane of warp 0 (the southern halo row: it has no momentum/store work after the
//       barrier, so the TMA issue rides in its idle time).  It issues the loads of z-step g+DEPTH into the ring slot
//       the barrier has just freed, but only once every CTA of the grid has reached step g − slack (grid-wide
//       progress counter, polled before the barrier so the L2 round trip is hidden): neighbouring tiles then load
//       their shared halo rows within a few steps of each other and the second reader hits L2 instead of HBM.
//   all warps: wait on the slot's full mbarrier, update, ONE __syncthreads per step, store.
//   MULTI: the launch runs a.niter iterations (ping-pong S_in ↔ S_out, a grid-wide barrier with generic→async proxy
//       fences between them) and applies flow_bcs! itself: every thread that holds the source of a ghost / boundary
//       value (no_slip! → free_slip! as complete sweeps, same gather as k_bc_box3) also stores its images.
template <int BY, bool FINITE_DT, bool DIAG, int NST, bool RHOG, bool MULTI, bool PUSH = false, bool TRAIL = false>
__global__ void __launch_bounds__(32 * BY, (BY <= 8 ? (FINITE_DT ? 2 : JR_VA_MINB8) : BY <= 10 ? 2 : 1)) k_va_tma(const __grid_constant__ VaArgs a)
{
    using M = SlotMap<FINITE_DT, RHOG>;
    constexpr int TY = BY - 2, TILE = 32 * BY;
    constexpr int NARR = M::NARR, SLOT = NARR * TILE;
    constexpr uint32_t TILE_BYTES = TILE * 8;
    constexpr int DEPTH = NST - 1;  // prefetch distance in z-steps
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NST];
    double *const sm = reinterpret_cast<double *>(smem_raw);

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int G = TRAIL ? (int)gridDim.x - a.tr.ntrail : (int)gridDim.x, cta = blockIdx.x;
    if (TRAIL && cta >= G) {
        jr_va_trailer<32 * BY>(a, cta - G, tid);
        return;
    }
    const int nstep = a.kchunk + 2;
    const int ntile = a.ntx * a.nty, nitem = ntile * a.nchunk;
    const int my_rounds = cta < nitem ? (nitem - cta + G - 1) / G : 0;
    const int my_steps = my_rounds * nstep;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++) jr_mbar_init(&full_bar[s], 1);
        jr_fence_mbar_init();
    }
    __syncthreads();
    if (PUSH) {
        // everything this rank launched before this kernel (previous iteration + its BC kernel, or the layout entry) is complete:
        // tell the neighbours, then wait until they say the same — their pushes into our in-set are then visible, and they no longer
        // read the set our pushes of THIS iteration go to
        if (tid < 27) {
            jr_comm_sig *pe = a.push.sig_peer[tid];
            if (pe) {
                if (cta == 0) {
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&pe->push_flags[a.push.rank]), "l"(a.push.epoch) : "memory");
                }
                const unsigned long long *mine = &a.push.sig_mine->push_flags[a.push.nbr_rank[tid]];
                unsigned long long v;
                do {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
                } while (v < a.push.epoch);
            }
        }
        __syncthreads();
        asm volatile("fence.proxy.async.global;" ::: "memory");
    }

    // ---- producer state (CTA-uniform): the next z-step to load ----
    int p_g = 0, p_r = 0, p_l = 0, p_slot = 0, p_x0 = 0, p_y0 = 0, p_kb = 0;
    const CUtensorMap *mS = &a.mS5;                // in-set of the current iteration
    unsigned long long pbase = a.progress_base;    // progress counter value at the start of the current iteration
    auto p_decode = [&]() {
        const int item = p_r * G + cta;
        const int chunk = item / ntile, t = item - chunk * ntile;
        const int by = t / a.ntx, bx = t - by * a.ntx;
        p_x0 = bx * TXW; p_y0 = by * TY; p_kb = chunk * a.kchunk;
    };
    // loads of step p_g (one elected lane).  Step 0 of an item only needs V, η[, G] of plane kb−1 (queue fill).
    // part 0: everything; 1: the arrival-plane boxes (+ the expect_tx of the whole step); 2: the compute-plane boxes only
    auto p_issue = [&](int part) {
        const uint64_t pld = jr_l2_policy(a.pol_ld);
        const int k = p_kb - 2 + p_l;
        const bool full = p_l > 0;
        uint64_t *bar = &full_bar[p_slot];
        double *d = sm + (size_t)p_slot * SLOT;
        const int za = k + 2, zc = k + 1;  // arrival plane (V, η, top edges) / compute plane
        if (part != 2) {
            uint32_t bytes = (5 + 1 + (FINITE_DT ? 1 : 0)) * TILE_BYTES;
            if (full) bytes = NARR * TILE_BYTES;
            jr_mbar_arrive_expect_tx(bar, bytes);
            jr_tma_load_4d_hint(d + T_Vx * TILE, mS, p_x0, p_y0, S_Vx, za, bar, pld);
            jr_tma_load_4d_hint(d + T_eta * TILE, &a.mC1, p_x0, p_y0, C_eta, za, bar, pld);
            if (FINITE_DT) jr_tma_load_4d_hint(d + M::G * TILE, &a.mD1, p_x0, p_y0, D_G, za, bar, pld);
        }
        if (full && part != 1) {
            jr_tma_load_4d_hint(d + T_tzz * TILE, mS, p_x0, p_y0, S_tzz, zc, bar, pld);
            if (FINITE_DT) {
                jr_tma_load_4d_hint(d + M::oyz * TILE, &a.mD2, p_x0, p_y0, D_oyz, za, bar, pld);
                jr_tma_load_4d_hint(d + M::K * TILE, &a.mD7, p_x0, p_y0, D_K, zc, bar, pld);
            }
            if (RHOG) jr_tma_load_4d_hint(d + M::ett * TILE, &a.mC4, p_x0, p_y0, C_ett, zc, bar, pld);
            else jr_tma_load_4d_hint(d + M::ett * TILE, &a.mC1, p_x0, p_y0, C_ett, zc, bar, pld);
        }
    };
    // overlapped halo exchange: the V planes of the in-set arrive in z order (see VaArgs::halo_prog)
    unsigned long long halo_seen = 0;
    auto p_halo_wait = [&]() {
        if (a.halo_prog == nullptr) return;
        const int za = p_kb + p_l;  // arrival plane of this step
        if (za < a.halo_head) return;
        // only tiles whose box touches a ghost face with a neighbour read exchanged values (the top z ghost plane: every tile)
        const int hf = a.halo_faces;
        const bool touch = ((hf & 1) && p_x0 == 0) || ((hf & 2) && p_x0 + 32 > nx + 1) || ((hf & 4) && p_y0 == 0) || ((hf & 8) && p_y0 + BY > ny + 1) ||
                           ((hf & 32) && za >= nz + 1);
        if (!touch) return;
        const unsigned long long need = a.halo_base + (unsigned long long)min(za + 1, a.halo_pz);
        if (halo_seen >= need) return;
        for (;;) {
            unsigned long long m = ~0ull;
            for (int q = 0; q < a.halo_nslots; q++) {
                unsigned long long v;
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.halo_prog + q) : "memory");
                m = v < m ? v : m;
            }
            if (m >= need) { halo_seen = m; break; }
            __nanosleep(1000);   // polite: up to a few hundred lanes poll the same line
        }
        {
            unsigned long long v;  // acquire (the values polled above were relaxed)
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.halo_prog) : "memory");
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");  // generic-proxy stores of the exchange → our TMA loads
    };
    auto p_advance = [&]() {
        ++p_g;
        if (++p_slot == NST) p_slot = 0;
        if (++p_l == nstep) {
            p_l = 0;
            ++p_r;
            if (p_r < my_rounds) p_decode();
        }
    };
    // progress-counter value that says "every active CTA has reached global step gt"
    auto p_target = [&](int gt) -> unsigned long long {
        const int rr = gt / nstep, ll = gt - rr * nstep;
        const int act = min(G, nitem - rr * G);
        return pbase + (unsigned long long)G * nstep * rr + (unsigned long long)act * (ll + 1);
    };
    const double _dx = a._dx, _dy = a._dy, _dz = a._dz, th = a.theta_dtau;
    const double inv3 = jr_inv(3.0);
    const double dtr_inf = jr_inv(th + 1.0);  // compute_dτ_r with 1/(G dt) = 0: fma(η, 0, 1) = 1
    const double r_th = jr_div_rcp(th);       // refined 1/θ_dτ for the exact quotient x / θ_dτ
    const uint64_t pst = jr_l2_policy(a.pol_st);
    const size_t pxy = (size_t)a.PX * a.PY;

    // ---- register queue: state of plane k carried along z ----
    double vx0 = 0, vy0 = 0, vz0 = 0;                  // V(k) at this thread's position
    double dxx0 = 0, dyy0 = 0, exy0 = 0;               // ∂xVx, ∂yVy, ε_xy of plane k
    double eta0 = 1, etaxy0 = 1, sxz0 = 0, syz0 = 0;   // η(k), η̄xy(k), η pair sums of plane k
    double g0 = 1, gxy0 = 1, gsxz0 = 0, gsyz0 = 0;     // same for G (finite dt)
    double tzz_p = 0, P_p = 0, fz_p = 0, ett_p = 1;    // plane k−1: new τzz, new P, ρgz, ητ
    double txz_b = 0, tyz_b = 0, sRz = 0;              // new bottom-edge stresses (kz = k), partial Rz of face k

    // queue entries of plane A from the slot that holds V(A), η(A) (reads at own position and W/E/S/N/SW neighbours)
#define JR_FILL_QUEUE(p)                                                                                  \
    do {                                                                                                  \
        vx0 = (p)[T_Vx * TILE]; vy0 = (p)[T_Vy * TILE]; vz0 = (p)[T_Vz * TILE]; eta0 = (p)[T_eta * TILE]; \
        dxx0 = (-vx0 + (p)[T_Vx * TILE + 1]) * _dx;                                                       \
        dyy0 = (-vy0 + (p)[T_Vy * TILE + 32]) * _dy;                                                      \
        exy0 = 0.5 * (_dy * (vx0 - (p)[T_Vx * TILE - 32]) + _dx * (vy0 - (p)[T_Vy * TILE - 1]));          \
        {                                                                                                 \
            const double eW_ = (p)[T_eta * TILE - 1], eS_ = (p)[T_eta * TILE - 32], eSW_ = (p)[T_eta * TILE - 33]; \
            etaxy0 = 0.25 * (eSW_ + eS_ + eW_ + eta0);                                                    \
            sxz0 = eW_ + eta0;                                                                            \
            syz0 = eS_ + eta0;                                                                            \
        }                                                                                                 \
        if (FINITE_DT) {                                                                                  \
            g0 = (p)[M::G * TILE];                                                                        \
            const double gW_ = (p)[M::G * TILE - 1], gS_ = (p)[M::G * TILE - 32], gSW_ = (p)[M::G * TILE - 33]; \
            gxy0 = 0.25 * (gSW_ + gS_ + gW_ + g0);                                                        \
            gsxz0 = gW_ + g0;                                                                             \
            gsyz0 = gS_ + g0;                                                                             \
        }                                                                                                 \
    } while (0)

    // boundary conditions of the in-kernel flow_bcs! (MULTI): ghost-image directions of this thread (−1 low side,
    // +1 high side, 0 none) for sources on the first / last interior row; per item (x, y) and per step (z)
#define JR_GHOSTS(q, v, z, ma, sa, ea, mb, sb, eb)                                                            \
    do {                                                                                                      \
        if ((ma) | (mb)) {                                                                                    \
            const double ga_ = (z) ? 1.0 : ((ma) > 0 ? a.bc_sg[2 * (ea) + 1] : a.bc_sg[2 * (ea)]);                \
            const double gb_ = (z) ? 1.0 : ((mb) > 0 ? a.bc_sg[2 * (eb) + 1] : a.bc_sg[2 * (eb)]);                \
            if (ma) jr_st_hint((q) + (ma) * (ptrdiff_t)(sa), ga_ * (v), pst);                                 \
            if (mb) jr_st_hint((q) + (mb) * (ptrdiff_t)(sb), gb_ * (v), pst);                                 \
            if ((ma) && (mb)) jr_st_hint((q) + (ma) * (ptrdiff_t)(sa) + (mb) * (ptrdiff_t)(sb), (ga_ * gb_) * (v), pst); \
        }                                                                                                     \
    } while (0)

    int slot = 0;
    uint32_t parity = 0;
    const int niter = MULTI ? a.niter : 1;
    for (int it = 0; it < niter; ++it) {
    double *const outset = (MULTI && (it & 1)) ? a.out2 : a.out;
    if (MULTI) {
        mS = (it & 1) ? &a.mS5b : &a.mS5;
        pbase = a.progress_base + (unsigned long long)it * (unsigned long long)nitem * (unsigned long long)nstep;
        p_g = 0; p_r = 0; p_l = 0;
    }
    p_decode();
    // prologue: the first DEPTH steps
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
        if (p_g < my_steps) {
            if (ty == 0) {
                if (jr_elect_one()) {
                    if (d == 0 && it == 0) {
                        jr_tma_prefetch_desc(&a.mS5);
                        jr_tma_prefetch_desc(&a.mC1);
                        jr_tma_prefetch_desc(&a.mC4);
                        if (MULTI) jr_tma_prefetch_desc(&a.mS5b);
                    }
                    p_halo_wait();
                    p_issue(0);
                }
            }
            p_advance();
        }
    }
    int g = 0;
    for (int r = 0; r < my_rounds; ++r) {
        const int item = r * G + cta;
        const int chunk = item / ntile, t = item - chunk * ntile;
        const int by = t / a.ntx, bx = t - by * a.ntx;
        const int X = bx * TXW + tx, Y = by * TY + ty;  // box coordinates of this thread; cell (gi, gj) = (X-1, Y-1)
        const int gi = X - 1, gj = Y - 1;
        const int kb = chunk * a.kchunk, ke = min(kb + a.kchunk, nz);
        const bool own = tx >= 1 && tx <= TXW && ty >= 1 && ty <= TY;
        const bool cell = own && gi < nx && gj < ny;
        const bool vxy = own && gi <= nx && gj <= ny;  // xy edge exists
        const bool vxz = own && gi <= nx && gj < ny;   // xz edge exists
        const bool vyz = own && gi < nx && gj <= ny;   // yz edge exists
        const bool stVx = cell && gi >= 1, stVy = cell && gj >= 1;
        // MULTI: this thread holds sources of boundary / ghost values in x or y (first / last interior row, boundary faces)
        const bool bxy = MULTI && own && (gi == 0 || gi >= nx - 1 || gj == 0 || gj >= ny - 1);
        // out-set pointer of (X, Y) in plane group Z = k+1: ((Z·10 + a)·PY + Y)·PX + X; starts at Z = kb (step k = kb−1)
        double *po = outset + ((size_t)kb * S_N) * pxy + (size_t)Y * a.PX + X;

        for (int l = 0; l < nstep; ++l) {
            const int k = kb - 2 + l;
            jr_mbar_wait(&full_bar[slot], parity);
            double *const p = sm + (size_t)slot * SLOT + tid;

            double txx_n, tyy_n, tzz_n, txy_n, txz_n, tyz_n, P_n, divV, RP, exx, eyy, ezz, exz_t, eyz_t;
            if (l > 0) {
                // ---- R1: top-edge strain rates / viscosities (plane k+1 arrives), centre of plane k, six new stresses ----
                const double vz1 = p[T_Vz * TILE], eta1 = p[T_eta * TILE];
                exz_t = 0.5 * (_dz * (p[T_Vx * TILE] - vx0) + _dx * (vz1 - p[T_Vz * TILE - 1]));
                eyz_t = 0.5 * (_dz * (p[T_Vy * TILE] - vy0) + _dy * (vz1 - p[T_Vz * TILE - 32]));
                const double etaxz_t = 0.25 * (sxz0 + p[T_eta * TILE - 1] + eta1);
                const double etayz_t = 0.25 * (syz0 + p[T_eta * TILE - 32] + eta1);
                const double c_txx = p[T_txx * TILE], c_tyy = p[T_tyy * TILE], c_tzz = p[T_tzz * TILE];
                const double c_txy = p[T_txy * TILE], c_txz = p[T_txz * TILE], c_tyz = p[T_tyz * TILE];
                const double c_P = p[T_P * TILE];
                const double dzz = (-vz0 + vz1) * _dz;
                divV = dxx0 + dyy0 + dzz;
                const double d3 = divV * inv3;
                exx = dxx0 - d3; eyy = dyy0 - d3; ezz = dzz - d3;
                if (FINITE_DT) {
                    const double dt = a.dt;
                    const double g1 = p[M::G * TILE];
                    const double gxz_t = 0.25 * (gsxz0 + p[M::G * TILE - 1] + g1);
                    const double gyz_t = 0.25 * (gsyz0 + p[M::G * TILE - 32] + g1);
                    P_n = c_P;
                    jr_compute_P_point(RP, P_n, p[M::P0 * TILE], divV, p[M::Q * TILE], eta0, p[M::K * TILE], g0, dt, a.r, th);
                    {
                        const double _Gdt = jr_inv(g0 * dt), dtr = jr_dtau_r(th, eta0, _Gdt);
                        txx_n = c_txx + jr_stress_increment(c_txx, p[M::oxx * TILE], eta0, exx, _Gdt, dtr);
                        tyy_n = c_tyy + jr_stress_increment(c_tyy, p[M::oyy * TILE], eta0, eyy, _Gdt, dtr);
                        tzz_n = c_tzz + jr_stress_increment(c_tzz, p[M::ozz * TILE], eta0, ezz, _Gdt, dtr);
                    }
                    {
                        const double _Gdt = jr_inv(gxy0 * dt), dtr = jr_dtau_r(th, etaxy0, _Gdt);
                        txy_n = c_txy + jr_stress_increment(c_txy, p[M::oxy * TILE], etaxy0, exy0, _Gdt, dtr);
                    }
                    {
                        const double _Gdt = jr_inv(gxz_t * dt), dtr = jr_dtau_r(th, etaxz_t, _Gdt);
                        txz_n = c_txz + jr_stress_increment(c_txz, p[M::oxz * TILE], etaxz_t, exz_t, _Gdt, dtr);
                    }
                    {
                        const double _Gdt = jr_inv(gyz_t * dt), dtr = jr_dtau_r(th, etayz_t, _Gdt);
                        tyz_n = c_tyz + jr_stress_increment(c_tyz, p[M::oyz * TILE], etayz_t, eyz_t, _Gdt, dtr);
                    }
                } else {
                    // _Kdt = _Gdt = _dt = 0 exactly (PressureKernels.jl:186-195 with dt = Inf)
                    RP = -divV;
                    const double psi = jr_div_by(jr_inv_nr(jr_inv_nr(eta0)) * a.r, th, r_th);
                    P_n = (-divV) * psi + c_P;
                    txx_n = c_txx + dtr_inf * fma(2.0 * eta0, exx, -c_txx);
                    tyy_n = c_tyy + dtr_inf * fma(2.0 * eta0, eyy, -c_tyy);
                    tzz_n = c_tzz + dtr_inf * fma(2.0 * eta0, ezz, -c_tzz);
                    txy_n = c_txy + dtr_inf * fma(2.0 * etaxy0, exy0, -c_txy);
                    txz_n = c_txz + dtr_inf * fma(2.0 * etaxz_t, exz_t, -c_txz);
                    tyz_n = c_tyz + dtr_inf * fma(2.0 * etayz_t, eyz_t, -c_tyz);
                }
                // ---- publish the new stresses / pressure in place (only their owner read the old values) ----
                p[T_txx * TILE] = txx_n; p[T_tyy * TILE] = tyy_n; p[T_P * TILE] = P_n;
                p[T_txy * TILE] = txy_n; p[T_txz * TILE] = txz_n; p[T_tyz * TILE] = tyz_n;
            }
            // lane 0 of warp 0 samples the grid-wide progress before the barrier (latency hidden behind it)
            unsigned long long seen = 0;
            if (tid == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.progress) : "memory");
            __syncthreads();
            // every thread is past its reads of the previous step's slot: refill it with the loads of step g+DEPTH
            if (ty == 0) {
                if (tid == 0) {
                    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(a.progress), "l"(1ull) : "memory");
                    if (p_g < my_steps) {
                        const int gt = g - a.slack;  // soft lock-step: everybody has reached step g − slack
                        if (gt >= 0) {
                            const unsigned long long target = p_target(gt);
                            // (the lock-step is a performance device — shared halo rows are fetched together — not a correctness one:
                            //  a launch that is not cooperative gives up after spin_cap polls instead of relying on co-residency)
                            for (int spins = 0; seen < target && (a.spin_cap == 0 || spins < a.spin_cap); ++spins)
                                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.progress) : "memory");
                        }
                        p_halo_wait();
                        jr_fence_proxy_async();
                        p_issue(a.stagger_ns >= 0 ? 1 : 0);
                    }
                }
                __syncwarp();
            } else if (TRAIL && ty == BY - 1) {
                // northern halo warp (no store work): every thread of the CTA has issued the stores of the steps before this barrier —
                // publish their count for the trailing CTAs (release: fence, then the flag)
                if (tx == 0 && l >= 1 && (l % a.tr.sig_every) == 0) {
                    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.tr.done + cta), "l"(a.tr.done_base + (unsigned long long)l) : "memory");
                }
            } else if (ty == BY - 1 && a.stagger_ns >= 0) {
                // second half of the step's loads, from the northern halo warp (it has no store work either), a little later:
                // the grid runs in lock-step, so issuing everything at the barrier makes the whole GPU's requests arrive in bursts
                if (p_g < my_steps && p_l > 0) {
                    if (jr_elect_one()) {
                        if (a.stagger_ns > 0) __nanosleep(a.stagger_ns);
                        jr_fence_proxy_async();
                        p_issue(2);
                    }
                }
                __syncwarp();
            }
            if (p_g < my_steps) p_advance();
            ++g;

            if (l > 0) {
                // ---- R2: momentum residuals and velocity update of plane k, partial Rz of face k+1 ----
                const double c_fz = RHOG ? p[M::fz * TILE] : a.fzc, c_ett = p[M::ett * TILE];
                const double sRz_next = _dx * (p[T_txz * TILE + 1] - txz_n) + _dy * (p[T_tyz * TILE + 32] - tyz_n);
                const bool kin = k >= kb && k < ke;  // this chunk owns plane k (k = kb−1 is the warm-up plane)
                // MULTI: flow_bcs! inside the kernel.  mz: CTA-uniform z ghost-image direction of this plane; bz: this thread
                // holds the source of some boundary / ghost value at this step (rare) — the only test on the common path
                const int mz = !MULTI ? 0 : (k == 0 ? -1 : (k == nz - 1 ? 1 : 0));
                const bool bz = MULTI && (bxy || mz != 0);
#define JR_MX (gi == 0 ? -1 : (gi == nx - 1 ? 1 : 0))
#define JR_MY (gj == 0 ? -1 : (gj == ny - 1 ? 1 : 0))
                // TRAIL: flow_bcs! on the two z faces for the elements whose source this thread holds (the x / y interior of the planes;
                // their rings belong to the trailing CTAs).  Tangential ghost planes Z = 0 / nz+1: ± the new interior value (free slip /
                // no slip) or, without a boundary condition on that side, the in-value; normal faces Vz(K = 0 / nz): 0 or the in-value.
                const bool zlo_bc = TRAIL && (a.tr.fs_lo[2] | a.tr.ns_lo[2]), zhi_bc = TRAIL && (a.tr.fs_hi[2] | a.tr.ns_hi[2]);
                if (TRAIL && k == -1 && !zlo_bc) {
                    if (stVx) jr_st_hint(&po[S_Vx * pxy], vx0, pst);
                    if (stVy) jr_st_hint(&po[S_Vy * pxy], vy0, pst);
                }
                if (kin) {
                    if (cell) {
                        jr_st_hint(&po[S_P * pxy], P_n, pst); jr_st_hint(&po[S_txx * pxy], txx_n, pst);
                        jr_st_hint(&po[S_tyy * pxy], tyy_n, pst); jr_st_hint(&po[S_tzz * pxy], tzz_n, pst);
                        if (DIAG) {
                            const size_t c = ((size_t)k * ny + gj) * nx + gi;
                            a.divV[c] = divV; a.RP[c] = RP; a.exx[c] = exx; a.eyy[c] = eyy; a.ezz[c] = ezz;
                            a.dP[c] = P_n; a.dtxx[c] = txx_n; a.dtyy[c] = tyy_n; a.dtzz[c] = tzz_n;
                        }
                    }
                    if (vxy) {
                        jr_st_hint(&po[S_txy * pxy], txy_n, pst);
                        if (DIAG) {
                            const size_t c = ((size_t)k * (ny + 1) + gj) * (nx + 1) + gi;
                            a.exy[c] = exy0; a.dtxy[c] = txy_n;
                        }
                    }
                    // x-momentum: face gi between cells gi−1 (W) and gi
                    if (stVx) {
                        const double R = (-p[T_txx * TILE - 1] + txx_n) * _dx + _dy * (p[T_txy * TILE + 32] - txy_n) +
                                         _dz * (txz_n - txz_b) - (-p[T_P * TILE - 1] + P_n) * _dx -
                                         0.5 * (RHOG ? (p[M::fx * TILE - 1] + p[M::fx * TILE]) : (a.fxc + a.fxc));
                        const double vn = vx0 + jr_div_nr(R * a.eta_dtau, 0.5 * (p[M::ett * TILE - 1] + c_ett));
                        jr_st_hint(&po[S_Vx * pxy], vn, pst);
                        if (TRAIL) {
                            if (k == 0 && zlo_bc) jr_st_hint(&po[S_Vx * pxy] - S_N * pxy, (a.tr.fs_lo[2] ? 1.0 : -1.0) * vn, pst);
                            if (k == nz - 1) jr_st_hint(&po[S_Vx * pxy] + S_N * pxy, zhi_bc ? (a.tr.fs_hi[2] ? 1.0 : -1.0) * vn : p[T_Vx * TILE], pst);
                        }
                        if (PUSH) jr_push_v(a.push, 0, gi, gj, k, nx, ny, nz, (size_t)(&po[S_Vx * pxy] - outset), vn);
                        if (DIAG) {
                            a.Rx[((size_t)k * ny + gj) * (nx - 1) + (gi - 1)] = R;
                            const size_t c = ((size_t)(k + 1) * (ny + 2) + gj + 1) * (nx + 1) + gi;
                            a.Ux[c] = vn * a.dt; a.dVx[c] = vn;
                        }
                        if (MULTI && bz) JR_GHOSTS(&po[S_Vx * pxy], vn, false, JR_MY, a.PX, 1, mz, S_N * pxy, 2);
                    } else if (MULTI && bz && vxz && (gi == 0 || gi == nx)) {
                        // boundary-normal face: kept (free slip) or zeroed (no slip), and its tangential ghosts
                        const bool z = (gi == 0 ? a.bc_nsn[0] : a.bc_nsn[1]) != 0;
                        const double v = z ? 0.0 : vx0;
                        jr_st_hint(&po[S_Vx * pxy], v, pst);
                        JR_GHOSTS(&po[S_Vx * pxy], v, z, JR_MY, a.PX, 1, mz, S_N * pxy, 2);
                    }
                    // y-momentum: face gj between cells gj−1 (S) and gj
                    if (stVy) {
                        const double R = _dx * (p[T_txy * TILE + 1] - txy_n) + _dy * (tyy_n - p[T_tyy * TILE - 32]) +
                                         _dz * (tyz_n - tyz_b) - (-p[T_P * TILE - 32] + P_n) * _dy -
                                         0.5 * (RHOG ? (p[M::fy * TILE - 32] + p[M::fy * TILE]) : (a.fyc + a.fyc));
                        const double vn = vy0 + jr_div_nr(R * a.eta_dtau, 0.5 * (p[M::ett * TILE - 32] + c_ett));
                        jr_st_hint(&po[S_Vy * pxy], vn, pst);
                        if (TRAIL) {
                            if (k == 0 && zlo_bc) jr_st_hint(&po[S_Vy * pxy] - S_N * pxy, (a.tr.fs_lo[2] ? 1.0 : -1.0) * vn, pst);
                            if (k == nz - 1) jr_st_hint(&po[S_Vy * pxy] + S_N * pxy, zhi_bc ? (a.tr.fs_hi[2] ? 1.0 : -1.0) * vn : p[T_Vy * TILE], pst);
                        }
                        if (PUSH) jr_push_v(a.push, 1, gi, gj, k, nx, ny, nz, (size_t)(&po[S_Vy * pxy] - outset), vn);
                        if (DIAG) {
                            a.Ry[((size_t)k * (ny - 1) + (gj - 1)) * nx + gi] = R;
                            const size_t c = ((size_t)(k + 1) * (ny + 1) + gj) * (nx + 2) + gi + 1;
                            a.Uy[c] = vn * a.dt; a.dVy[c] = vn;
                        }
                        if (MULTI && bz) JR_GHOSTS(&po[S_Vy * pxy], vn, false, JR_MX, 1, 0, mz, S_N * pxy, 2);
                    } else if (MULTI && bz && vyz && (gj == 0 || gj == ny)) {
                        const bool z = (gj == 0 ? a.bc_nsn[2] : a.bc_nsn[3]) != 0;
                        const double v = z ? 0.0 : vy0;
                        jr_st_hint(&po[S_Vy * pxy], v, pst);
                        JR_GHOSTS(&po[S_Vy * pxy], v, z, JR_MX, 1, 0, mz, S_N * pxy, 2);
                    }
                    // z-momentum: face k between planes k−1 and k
                    if (cell && k >= 1) {
                        const double R = sRz + (-tzz_p + tzz_n) * _dz - (-P_p + P_n) * _dz - 0.5 * (fz_p + c_fz);
                        const double vn = vz0 + jr_div_nr(R * a.eta_dtau, 0.5 * (ett_p + c_ett));
                        jr_st_hint(&po[S_Vz * pxy], vn, pst);
                        if (PUSH) jr_push_v(a.push, 2, gi, gj, k, nx, ny, nz, (size_t)(&po[S_Vz * pxy] - outset), vn);
                        if (DIAG) {
                            a.Rz[((size_t)(k - 1) * ny + gj) * nx + gi] = R;
                            const size_t c = ((size_t)k * (ny + 2) + gj + 1) * (nx + 2) + gi + 1;
                            a.Uz[c] = vn * a.dt; a.dVz[c] = vn;
                        }
                        if (MULTI && bz) JR_GHOSTS(&po[S_Vz * pxy], vn, false, JR_MX, 1, 0, JR_MY, a.PX, 1);
                    }
                    if (TRAIL && cell) {
                        // z boundary-normal faces: Vz(K = 0) = box plane 1 from the queue (k = 0), Vz(K = nz) = box plane nz+1 from the
                        // arrival plane (k = nz − 1)
                        if (k == 0) jr_st_hint(&po[S_Vz * pxy], a.tr.ns_lo[2] ? 0.0 : vz0, pst);
                        if (k == nz - 1) jr_st_hint(&po[(S_N + S_Vz) * pxy], a.tr.ns_hi[2] ? 0.0 : p[T_Vz * TILE], pst);
                    }
                    if (MULTI && mz != 0 && cell) {
                        // z-normal boundary faces: Vz(K = 0) from the queue (k = 0), Vz(K = nz) from the arrival plane
                        // (k = nz − 1); nz ≥ 3 on this path, so never both
                        const bool z = (mz < 0 ? a.bc_nsn[4] : a.bc_nsn[5]) != 0;
                        const double v = z ? 0.0 : (mz < 0 ? vz0 : p[T_Vz * TILE]);
                        double *const q = (mz < 0 ? po : po + S_N * pxy) + S_Vz * pxy;
                        jr_st_hint(q, v, pst);
                        JR_GHOSTS(q, v, z, JR_MX, 1, 0, JR_MY, a.PX, 1);
                    }
                }
                // top edges (kz = k+1) belong to the chunk that owns plane k; the kz = 0 edges to chunk 0's warm-up
                if (kin || k == -1) {
                    double *const pe = po + S_N * pxy;
                    if (vxz) {
                        jr_st_hint(&pe[S_txz * pxy], txz_n, pst);
                        if (DIAG) {
                            const size_t c = ((size_t)(k + 1) * ny + gj) * (nx + 1) + gi;
                            a.exz[c] = exz_t; a.dtxz[c] = txz_n;
                        }
                    }
                    if (vyz) {
                        jr_st_hint(&pe[S_tyz * pxy], tyz_n, pst);
                        if (DIAG) {
                            const size_t c = ((size_t)(k + 1) * (ny + 1) + gj) * nx + gi;
                            a.eyz[c] = eyz_t; a.dtyz[c] = tyz_n;
                        }
                    }
                }
                tzz_p = tzz_n; P_p = P_n; fz_p = c_fz; ett_p = c_ett;
                txz_b = txz_n; tyz_b = tyz_n; sRz = sRz_next;
                po += S_N * pxy;
            }
            // ---- the queue for the next step: plane k+1 becomes plane k ----
            JR_FILL_QUEUE(p);
            if (++slot == NST) { slot = 0; parity ^= 1u; }
        }
    }
    if (TRAIL) {
        __syncthreads();
        if (tid == 0) {
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.tr.done + cta), "l"(a.tr.done_base + (unsigned long long)nstep) : "memory");
        }
    }
    if (MULTI && it + 1 < niter) {
        // grid-wide barrier between iterations: every store of this iteration (generic proxy) must be visible to
        // the TMA loads (async proxy) of every other CTA, and nobody may overwrite a set others still read
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(a.gbar), "l"(1ull) : "memory");
            const unsigned long long target = a.gbar_base + (unsigned long long)(it + 1) * (unsigned long long)G;
            unsigned long long seen = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.gbar) : "memory");
            } while (seen < target);
            __threadfence();
            asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        __syncthreads();
    }
    }  // iterations
#undef JR_FILL_QUEUE
#undef JR_GHOSTS
#undef JR_MX
#undef JR_MY
}

// ------------------------------------------------------------------------------------------------------
// Box-layout helpers
struct BoxArr {       // one array of a box set
    double *p;        // set base + a·pxy
    long sy, sz;      // strides in elements: PX, NA·pxy
};
__device__ __forceinline__ size_t box_idx(const BoxArr &b, int X, int Y, int Z) { return (size_t)Z * b.sz + (size_t)Y * b.sy + X; }

// dense (n0,n1,n2) ↔ box at offset (o0,o1,o2).  mode 0: dense → box, 1: box → dense,
// 2: dense → box over the whole ghosted extent (n+2)^3 with clamped source indices (η, G: the reference's clamped
//    neighbour reads, MiniKernels.jl:133-147, become plain reads of the ghost copies)
struct PackJob {
    double *dense;
    BoxArr box;
    int n[3], o[3], mode;
};
#define PACK_MAX 25
struct PackArgs {
    PackJob j[PACK_MAX];
    int njobs;
};
__global__ void k_box_pack(const __grid_constant__ PackArgs pa)
{
    const PackJob &J = pa.j[blockIdx.y];
    const int g = J.mode == 2 ? 2 : 0;
    const int e0 = J.n[0] + g, e1 = J.n[1] + g, e2 = J.n[2] + g;
    for (long r = blockIdx.x; r < (long)e1 * e2; r += gridDim.x) {
        const int j = (int)(r % e1), k = (int)(r / e1);
        for (int i = threadIdx.x; i < e0; i += blockDim.x) {
            if (J.mode == 2) {
                const int si = jr_clamp(i - 1, 0, J.n[0] - 1), sj = jr_clamp(j - 1, 0, J.n[1] - 1), sk = jr_clamp(k - 1, 0, J.n[2] - 1);
                J.box.p[box_idx(J.box, i, j, k)] = J.dense[((size_t)sk * J.n[1] + sj) * J.n[0] + si];
            } else {
                const size_t d = ((size_t)k * J.n[1] + j) * J.n[0] + i;
                const size_t b = box_idx(J.box, i + J.o[0], j + J.o[1], k + J.o[2]);
                if (J.mode == 0) J.box.p[b] = J.dense[d];
                else J.dense[d] = J.box.p[b];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Ping-pong boundary kernel on the box layout: fills every element of V_out the fused kernel does not compute,
// i.e. the tangential ghost layers and the boundary-normal faces, from (a) the freshly computed interior of
// V_out and (b) V_in for layers no boundary condition touches (prescribed values, or halo planes that the
// exchange overwrites afterwards).  Semantics = no_slip! → free_slip! of the reference applied as complete
// sweeps (no_slip.jl:21-54, free_slip.jl:15-70 incl. quirk Q2); every ghost value is gathered from its
// fully-clamped source with the product of the per-dimension signs, so ghost edges/corners are deterministic.
// With diag set it also writes U = V·dt (dense user array) for those elements from V_in — the reference takes U
// before flow_bcs! (Stokes3D.jl:118-119).
struct BcArrB {
    BoxArr in, out;
    double *U;   // dense
    double *Vd;  // dense V of the user (written together with U on observable iterations)
    int n[3];    // dense extents of this velocity component
    int o[3];    // dense → box offset
    int normal;  // normal dimension of this component
    long set_off;  // out.p − base of the out-set (push exchange: offsets are relative to the set, which every rank lays out alike)
};
struct BcArgsB {
    BcArrB A[3];
    int lo_fs[3], hi_fs[3], lo_ns[3], hi_ns[3];  // per dimension and side: free-slip / no-slip active
    int diag;
    double dt;
    int do_push, ncell[3];                       // multi-GPU push exchange: elements on a send plane also go to the neighbours' ghosts,
    PushArgs push;                               // elements the neighbours push to us (halo planes) are left alone
    int pack;                                    // multi-GPU: blocks 18 … 35 of the launch pack the send planes (post-flow_bcs! values: the
    double *stage;                               // same gather evaluated on the planes ol − 1 / n − ol) into this rank's staging buffer,
    long stage_off[3];                           // laid out as k_halo_pack does (comm.cuh) — one launch less per exchange
    int skip_lo[3], skip_hi[3];                  // multi-GPU: this face has a neighbour — the exchange that follows overwrites the whole
                                                 // plane, nobody reads it in between (the sources of every gather are interior values)
};

// signal + wait of the push protocol outside the iteration kernel (solve exit: everybody's last pushes have landed)
__global__ void k_push_sync(const __grid_constant__ PushArgs pu)
{
    const int tid = threadIdx.x;
    if (tid < 27) {
        jr_comm_sig *pe = pu.sig_peer[tid];
        if (pe) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&pe->push_flags[pu.rank]), "l"(pu.epoch) : "memory");
            const unsigned long long *mine = &pu.sig_mine->push_flags[pu.nbr_rank[tid]];
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            } while (v < pu.epoch);
        }
    }
}

#define BC_ROWS 4   // rows per thread: the four gathers are independent and issued back to back (the kernel is pure latency)
__global__ void __launch_bounds__(256) k_bc_box3(const __grid_constant__ BcArgsB b)
{
    const bool packjob = blockIdx.z >= 18;
    const int bz = packjob ? blockIdx.z - 18 : blockIdx.z;
    const int which = bz / 6, plane = bz % 6;  // component, (dim, lo/hi)
    const BcArrB &A = b.A[which];
    const int d = plane >> 1, hi = plane & 1;
    const bool nbr = hi ? b.skip_hi[d] : b.skip_lo[d];
    if (packjob ? !nbr : (!b.diag && nbr)) return;
    const int ol_d = 2 + A.n[d] - b.ncell[d];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    // fastest-varying free coordinate on threadIdx.x: for d = 0 planes use (dim1, dim2), else dim0 first
    const int u = (d == 0) ? 1 : 0, v = (d == 2) ? 1 : 2;
    if (p >= A.n[u]) return;
    size_t ic[BC_ROWS], dc[BC_ROWS];
    double val[BC_ROWS], vin[BC_ROWS];
    bool ok[BC_ROWS], rcv[BC_ROWS];
    int pf[BC_ROWS][3];
#pragma unroll
    for (int r = 0; r < BC_ROWS; r++) {
        const int q = (blockIdx.y * BC_ROWS + r) * blockDim.y + threadIdx.y;
        ok[r] = q < A.n[v];
        val[r] = 0.0; vin[r] = 0.0; ic[r] = 0; dc[r] = 0; rcv[r] = false; pf[r][0] = pf[r][1] = pf[r][2] = 0;
        if (!ok[r]) continue;
        int c[3];
        c[d] = packjob ? (hi ? A.n[d] - ol_d : ol_d - 1) : (hi ? A.n[d] - 1 : 0);
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
        if (b.do_push) {
            // halo planes towards a neighbour are written by that neighbour's pushes: leave them alone (box set), but keep the
            // dense diagnostics below
            bool recv = false;
#pragma unroll
            for (int e = 0; e < 3; e++) recv = recv || (c[e] == 0 && b.push.has_lo[e]) || (c[e] == A.n[e] - 1 && b.push.has_hi[e]);
            rcv[r] = recv;
#pragma unroll
            for (int e = 0; e < 3; e++) {
                const int ol = 2 + A.n[e] - b.ncell[e];
                pf[r][e] = (c[e] == ol - 1 && b.push.has_lo[e]) ? -1 : ((c[e] == A.n[e] - ol && b.push.has_hi[e]) ? 1 : 0);
            }
        }
        ic[r] = box_idx(A.out, c[0] + A.o[0], c[1] + A.o[1], c[2] + A.o[2]);
        const size_t is = box_idx(A.out, s[0] + A.o[0], s[1] + A.o[1], s[2] + A.o[2]);
        bool computed = true;
#pragma unroll
        for (int e = 0; e < 3; e++) computed = computed && s[e] >= 1 && s[e] <= A.n[e] - 2;
        val[r] = zero ? 0.0 : sign * (computed ? A.out.p[is] : A.in.p[is]);
        if (packjob) ic[r] = (size_t)(b.stage_off[which] + jr_stage_plane_off(A.n, d, hi)) + (size_t)q * A.n[u] + p;   // jr_stage_elem
        else if (b.diag) {
            dc[r] = ((size_t)c[2] * A.n[1] + c[1]) * A.n[0] + c[0];
            vin[r] = A.in.p[ic[r]];
        }
    }
    if (packjob) {
#pragma unroll
        for (int r = 0; r < BC_ROWS; r++)
            if (ok[r]) b.stage[ic[r]] = val[r];
        return;
    }
#pragma unroll
    for (int r = 0; r < BC_ROWS; r++) {
        if (!ok[r]) continue;
        if (!rcv[r]) {
            A.out.p[ic[r]] = val[r];
            if (b.do_push && (pf[r][0] | pf[r][1] | pf[r][2])) {
                const long off = A.set_off + (long)ic[r];
#pragma unroll
                for (int m = 1; m < 8; m++) {
                    if (((m & 1) && !pf[r][0]) || ((m & 2) && !pf[r][1]) || ((m & 4) && !pf[r][2])) continue;
                    const int d0 = (m & 1) ? pf[r][0] : 0, d1 = (m & 2) ? pf[r][1] : 0, d2 = (m & 4) ? pf[r][2] : 0;
                    double *base = b.push.peer_out[(d2 + 1) * 9 + (d1 + 1) * 3 + d0 + 1];
                    if (base) base[off - d0 * b.push.delta[0] - d1 * b.push.delta[1] - d2 * b.push.delta[2]] = val[r];
                }
            }
        }
        if (b.diag) {
            A.U[dc[r]] = vin[r] * b.dt;
            A.Vd[dc[r]] = val[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// min / max of up to three equally sized arrays (constant-field detection at solve entry)
#define MINMAX_BLOCKS 512
__global__ void k_minmax3(const double *a0, const double *a1, const double *a2, size_t n, double *part)
{
    __shared__ double sm[6][8];
    const double *A[3] = {a0, a1, a2};
    double mn[3], mx[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        mn[q] = INFINITY; mx[q] = -INFINITY;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const double v = A[q][i];
            mn[q] = fmin(mn[q], v); mx[q] = fmax(mx[q], v);
            if (v != v) { mn[q] = v; mx[q] = -v; }  // NaN poisons the range (never "constant")
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[q] = fmin(mn[q], __shfl_down_sync(0xffffffffu, mn[q], o));
            mx[q] = fmax(mx[q], __shfl_down_sync(0xffffffffu, mx[q], o));
        }
        if ((threadIdx.x & 31) == 0) { sm[2 * q][threadIdx.x >> 5] = mn[q]; sm[2 * q + 1][threadIdx.x >> 5] = mx[q]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[threadIdx.x][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = (threadIdx.x & 1) ? fmax(v, sm[threadIdx.x][w]) : fmin(v, sm[threadIdx.x][w]);
        part[blockIdx.x * 6 + threadIdx.x] = v;
    }
}
// one warp per quantity (6 warps): lane-strided scan of the block partials, then a shuffle reduction
__global__ void k_minmax3_final(const double *part, int nparts, double *out)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= 6) return;
    const bool is_max = q & 1;
    double v = is_max ? -INFINITY : INFINITY;
    bool nan = false;
    for (int b = lane; b < nparts; b += 32) {
        const double x = part[b * 6 + q];
        if (x != x) nan = true;
        v = is_max ? fmax(v, x) : fmin(v, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_down_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, w) : fmin(v, w);
    }
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) out[q] = nan ? NAN : v;
}

// ------------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int jr_encode_tensor_map_f64(CUtensorMap *out, void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                             const uint32_t *box, int l2promo)
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        JR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        JR_REQUIRE(p && qres == cudaDriverEntryPointSuccess, JR_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bd[5], es[5];
    for (int i = 0; i < rank; i++) {
        gd[i] = dims[i];
        bd[i] = box[i];
        es[i] = 1;
        if (i < rank - 1) gs[i] = strides_bytes[i];
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, base, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE,
                    l2promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                    : l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                    : l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                   : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    JR_REQUIRE(r == CUDA_SUCCESS, JR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu)", (int)r,
               rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2]);
    return JR_OK;
}

#define F(name) (s->f[JR_F_##name])

struct VaPlan {
    int nx = 0, ny = 0, nz = 0, PX = 0, PY = 0, PZ = 0;
    bool finite_dt = false;
    size_t pxy = 0;
    double *S[2] = {nullptr, nullptr}, *C = nullptr, *D = nullptr;
    CUtensorMap mS5[2], mC1, mC4, mD1, mD2, mD7;
    int BY = 0, nchunk = 1;
    int pol_ld = 2, pol_st = 1, l2promo = 3, slack = 1, stagger_ns = -1;
    unsigned long long *progress = nullptr, progress_base = 0, gbar_base = 0;
    bool rhog_const = false;  // ρg arrays are spatially constant: not streamed
    double fc[3] = {0, 0, 0};
    // box sets are zeroed when they are (re)allocated or re-shaped only: no kernel ever writes outside an array's own extent,
    // so the zero padding the stencils rely on survives from one solve to the next
    void *zeroed[4] = {nullptr, nullptr, nullptr, nullptr};
    int zdims[4] = {0, 0, 0, 0};  // nx, ny, nz, finite_dt of the zeroed layout
    bool last_diag = false;       // the last iteration was an observable one: the user's dense arrays are current
    // overlapped direct halo exchange (multi-GPU, see VaArgs::halo_prog)
    bool ovl = false, ovl_pending = false;
    int ovl_head = 8, ovl_chunk = 16, ovl_ctas = 6, halo_faces = 0;
    cudaStream_t ovl_stream = nullptr;
    cudaEvent_t ev_head = nullptr, ev_rest = nullptr;
    unsigned long long *hprog = nullptr, hbase = 0, ovl_count = 0;
    // TRAIL (flow_bcs! by trailing CTAs inside the iteration launch, see TrailArgs)
    bool trail = false;
    int trail_max = 8, trail_nb = 16, trail_nb_tail = 4, trail_tail = 16, trail_sig = 1;
    unsigned long long *done = nullptr, done_base = 0;
    // multi-GPU push exchange (see PushArgs): the state sets of every rank are CUDA-IPC mapped here
    bool push = false;
    void *shared_ptr[2] = {nullptr, nullptr};   // the S pointers the mappings below belong to
    std::vector<void *> peerS[2];
};
static std::map<jr_context *, VaPlan> g_plans;

int jr_stokes3d_VA_fused_supported(const jr_fields *s, const jr_stokes_opts *o)
{
    for (int q = 0; q < 6; q++)
        if (o->periodic[q]) return JR_ERR_UNSUPPORTED;  // periodic wrap: reference-structured path
    if (s->n[0] < 3 || s->n[1] < 3 || s->n[2] < 3) return JR_ERR_UNSUPPORTED;
    return JR_OK;
}

static int set_tma_map(CUtensorMap *m, double *base, const VaPlan &P, int NA, int nbox)
{
    const uint64_t dims[4] = {(uint64_t)P.PX, (uint64_t)P.PY, (uint64_t)NA, (uint64_t)P.PZ};
    const uint64_t str[3] = {(uint64_t)P.PX * 8, (uint64_t)P.pxy * 8, (uint64_t)NA * P.pxy * 8};
    const uint32_t box[4] = {32, (uint32_t)P.BY, (uint32_t)nbox, 1};
    return jr_encode_tensor_map_f64(m, base, 4, dims, str, box, P.l2promo);
}

static int run_pack(jr_context *ctx, const PackArgs &pa, int maxrows)
{
    if (pa.njobs == 0) return JR_OK;
    dim3 grid(maxrows < 4096 ? maxrows : 4096, pa.njobs, 1), block(128, 1, 1);
    k_box_pack<<<grid, block, 0, ctx->stream>>>(pa);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

static BoxArr box_arr(double *set, int a, int NA, const VaPlan &P) { return BoxArr{set + (size_t)a * P.pxy, (long)P.PX, (long)NA * (long)P.pxy}; }

// choose the tile height and z-chunking: maximise (round efficiency) × (1 − warm-up share) × (owned-row share).
// With the persistent grid G = min(resident slots, items) a CTA runs ceil(items / G) items.
static void choose_tiling(VaPlan &P, int sm_count)
{
    const int cand[3] = {10, 8, 16};
    double best = -1;
    for (int c = 0; c < 3; c++) {
        if (P.finite_dt && cand[c] != 8) continue;  // 25 tiles per slot: only the 8-row tile fits two CTAs per SM
        const int BY = cand[c], TY = BY - 2;
        const long tiles = (long)((P.nx + 1 + TXW - 1) / TXW) * ((P.ny + 1 + TY - 1) / TY);
        const long slots = (long)sm_count * (BY <= 10 ? 2 : 1);
        const double row_eff = (double)(P.ny + 1) / ((double)((P.ny + 1 + TY - 1) / TY) * BY);
        for (int nch = 1; nch <= 32; nch++) {
            const int kch = (P.nz + nch - 1) / nch;
            if (nch > 1 && kch < 12) break;
            const int real = (P.nz + kch - 1) / kch;
            const long items = tiles * real;
            const long G = items < slots ? items : slots;
            const long rounds = (items + G - 1) / G;
            // useful plane-steps / (slots × steps every slot is held)
            // measured on B200: the one-CTA-per-SM tile (16 rows) hides latency worse than two CTAs of 8 / 10 rows
            const double occ = BY <= 10 ? 1.0 : 0.92;
            const double eff = (double)tiles * P.nz / ((double)slots * rounds * (kch + 2)) * row_eff * occ;
            if (eff > best * 1.0001) { best = eff; P.BY = BY; P.nchunk = real; }
        }
    }
}

int jr_stokes3d_VA_fused_begin(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o)
{
    VaPlan &P = g_plans[ctx];
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const bool fin = std::isfinite(o->dt);
    P.nx = nx; P.ny = ny; P.nz = nz;
    P.PX = (nx + 2 + 3) & ~3; P.PY = ny + 2; P.PZ = nz + 2;
    P.pxy = (size_t)P.PX * P.PY;
    P.finite_dt = fin;
    const size_t plane = P.pxy * sizeof(double);
    void *p = nullptr;
    int st;
    const char *names[4] = {"box_S0", "box_S1", "box_C", "box_D"};
    const int NA[4] = {S_N, S_N, C_N, D_N};
    double **dst[4] = {&P.S[0], &P.S[1], &P.C, &P.D};
    for (int q = 0; q < 4; q++) {
        if (q == 3 && !fin) { P.D = nullptr; continue; }
        const size_t bytes = plane * NA[q] * P.PZ;
        if ((st = jr_ctx_scratch(ctx, names[q], bytes, &p))) return st;
        *dst[q] = (double *)p;
        const bool same = P.zeroed[q] == p && P.zdims[0] == nx && P.zdims[1] == ny && P.zdims[2] == nz;
        if (!same) JR_CUDA(cudaMemsetAsync(p, 0, bytes, ctx->stream));
        P.zeroed[q] = p;
    }
    P.zdims[0] = nx; P.zdims[1] = ny; P.zdims[2] = nz; P.zdims[3] = fin;
    P.last_diag = false;
    choose_tiling(P, ctx->sm_count);
    // test / tuning overrides of the tiling heuristic
    if (const char *fb = getenv("JRB200_VA_BY")) {
        const int b = atoi(fb);
        if (!fin && (b == 8 || b == 10 || b == 16)) P.BY = b;
    }
    if (const char *fc = getenv("JRB200_VA_NCHUNK")) {
        const int c = atoi(fc);
        if (c >= 1 && c <= nz) {
            const int kch = (nz + c - 1) / c;
            P.nchunk = (nz + kch - 1) / kch;
        }
    }
    if ((st = jr_ctx_scratch(ctx, "va_progress", 256, &p))) return st;
    P.progress = (unsigned long long *)p;
    P.progress_base = 0;
    P.gbar_base = 0;
    JR_CUDA(cudaMemsetAsync(p, 0, 256, ctx->stream));
    if ((st = jr_ctx_scratch(ctx, "va_done", 1024 * sizeof(unsigned long long), &p))) return st;
    P.done = (unsigned long long *)p;
    P.done_base = 0;
    JR_CUDA(cudaMemsetAsync(p, 0, 1024 * sizeof(unsigned long long), ctx->stream));
    // opt-in (JRB200_VA_TRAIL=1): bit-exact, but measured SLOWER (0.85–0.97 ms against 0.535 ms per iteration at 255^3,
    // profiles/r02_trail_experiment.md): publishing "the stores of step l are complete" needs a gpu-scope release in every main CTA,
    // and that fence waits for the SM's outstanding stores (≈ 1.7 µs under the z-march's load) in a warp the per-step barrier waits for
    P.trail = false; P.trail_max = 8; P.trail_nb = 16; P.trail_nb_tail = 4; P.trail_tail = 16; P.trail_sig = 1;
    if (const char *e = getenv("JRB200_VA_TRAIL")) P.trail = atoi(e) != 0;
    if (const char *e = getenv("JRB200_VA_TRAIL_CTAS")) P.trail_max = atoi(e) < 1 ? 1 : atoi(e);
    if (const char *e = getenv("JRB200_VA_TRAIL_NB")) P.trail_nb = atoi(e) < 1 ? 1 : atoi(e);
    if (const char *e = getenv("JRB200_VA_TRAIL_SIG")) P.trail_sig = atoi(e) < 1 ? 1 : atoi(e);
    if (const char *e = getenv("JRB200_VA_TRAIL_NBT")) P.trail_nb_tail = atoi(e) < 1 ? 1 : atoi(e);
    if (const char *e = getenv("JRB200_VA_TRAIL_TAIL")) P.trail_tail = atoi(e) < 0 ? 0 : atoi(e);
    // constant body force?  (one pass over ρg per solve; ρg ≡ 0 in SolVi / Taylor-Green / Burstedde-type benchmarks)
    {
        void *part = nullptr, *mm = nullptr;
        if ((st = jr_ctx_scratch(ctx, "minmax_part", (MINMAX_BLOCKS * 6 + 8) * sizeof(double), &part))) return st;
        mm = (double *)part + MINMAX_BLOCKS * 6;
        k_minmax3<<<MINMAX_BLOCKS, 256, 0, ctx->stream>>>(F(rhogx), F(rhogy), F(rhogz), (size_t)nx * ny * nz, (double *)part);
        k_minmax3_final<<<1, 192, 0, ctx->stream>>>((const double *)part, MINMAX_BLOCKS, (double *)mm);
        ctx->launches += 2;
        JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, mm, 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        JR_CUDA(cudaStreamSynchronize(ctx->stream));
        const double *h = ctx->h_pinned;
        P.rhog_const = h[0] == h[1] && h[2] == h[3] && h[4] == h[5] && !getenv("JRB200_VA_STREAM_RHOG");
        P.fc[0] = h[0]; P.fc[1] = h[2]; P.fc[2] = h[4];
    }
    P.slack = 1; P.pol_ld = 2; P.pol_st = 1; P.l2promo = 3;   // defaults (the knobs below are read at every solve entry)
    if (const char *e = getenv("JRB200_VA_SLACK")) P.slack = atoi(e);
    P.stagger_ns = -1;
    if (const char *e = getenv("JRB200_VA_STAGGER_NS")) P.stagger_ns = atoi(e);
    if (const char *e = getenv("JRB200_VA_POL_LD")) P.pol_ld = atoi(e);
    if (const char *e = getenv("JRB200_VA_POL_ST")) P.pol_st = atoi(e);
    if (const char *e = getenv("JRB200_VA_L2PROMO")) P.l2promo = atoi(e);
    // multi-GPU: in-iteration push exchange (opt-in, JRB200_VA_PUSH=1; every dimension must be large enough for the send planes —
    // index 2 / n − 3 — to be interior).  Bit-exact (tests/mgpu_worker.py) but measured SLOWER than pack + pull on 2 B200s
    // (0.688 vs 0.582 ms per iteration at 255^3, profiles/r02_push_exchange.md): the x-face planes are columns of the box layout, i.e.
    // ≈ 2·10^5 scattered 8-byte NVLink stores per face and iteration issued from the CTAs that pace the lock-stepped grid.
    // multi-GPU, opt-in (JRB200_VA_OVL=1): the pull of the exchange split along z, its bulk overlapped with the next iteration's z-march
    // (the @hide_communication analogue).  Bit-exact on 2 GPUs, measured no faster than pack + pull after every iteration (0.563 vs
    // 0.550 ms per iteration at 255^3 per GPU; the whole exchange costs 15 µs there, profiles/r02_overlap_exchange.md): default off.
    P.ovl = false; P.ovl_pending = false;
    {
        const char *e = getenv("JRB200_VA_OVL");
        const char *pu = getenv("JRB200_VA_PUSH");
        const bool want = (e && atoi(e) != 0) && !(pu && atoi(pu) == 1);
        if (ctx->comm && ctx->comm->nranks > 1 && !(ctx->comm->periods[0] | ctx->comm->periods[1] | ctx->comm->periods[2]) && want) {
            if (!P.ovl_stream) {
                int lo = 0, hi = 0;
                JR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                JR_CUDA(cudaStreamCreateWithPriority(&P.ovl_stream, cudaStreamNonBlocking, hi));
                JR_CUDA(cudaEventCreateWithFlags(&P.ev_head, cudaEventDisableTiming));
                JR_CUDA(cudaEventCreateWithFlags(&P.ev_rest, cudaEventDisableTiming));
            }
            if ((st = jr_ctx_scratch(ctx, "va_halo_prog", 64 * sizeof(unsigned long long), &p))) return st;
            if (P.hprog != (unsigned long long *)p) {
                P.hprog = (unsigned long long *)p;
                JR_CUDA(cudaMemsetAsync(p, 0, 64 * sizeof(unsigned long long), ctx->stream));
                P.hbase = 0; P.ovl_count = 0;
            }
            // rest CTAs: 256 threads × 74 registers each — they fit beside ONE main CTA on an SM, and the 288 column tiles of 255^3 leave
            // 8 such slots (2 × 148 − 288); whichever kernel becomes resident first, 296 − ovl_ctas ≥ 288 slots remain for the tiles
            P.ovl_head = 8; P.ovl_chunk = 16; P.ovl_ctas = 6;
            if (const char *h = getenv("JRB200_VA_OVL_HEAD")) P.ovl_head = atoi(h) < 1 ? 1 : atoi(h);
            if (const char *h = getenv("JRB200_VA_OVL_CHUNK")) P.ovl_chunk = atoi(h) < 1 ? 1 : atoi(h);
            if (const char *h = getenv("JRB200_VA_OVL_CTAS")) P.ovl_ctas = atoi(h) < 1 ? 1 : (atoi(h) > 32 ? 32 : atoi(h));
            P.halo_faces = 0;
            for (int d = 0; d < 3; d++) {
                if (ctx->comm->has_lo[d]) P.halo_faces |= 1 << (2 * d);
                if (ctx->comm->has_hi[d]) P.halo_faces |= 2 << (2 * d);
            }
            P.ovl = true;
        }
    }
    P.push = false;
    if (ctx->comm && ctx->comm->nranks > 1 && !(ctx->comm->periods[0] | ctx->comm->periods[1] | ctx->comm->periods[2]) && nx >= 8 && ny >= 8 && nz >= 8 && getenv("JRB200_VA_PUSH") && atoi(getenv("JRB200_VA_PUSH")) == 1) {
        for (int q = 0; q < 2; q++) {
            if (P.shared_ptr[q] != (void *)P.S[q] || (int)P.peerS[q].size() != ctx->comm->nranks) {
                P.peerS[q].assign(ctx->comm->nranks, nullptr);
                if ((st = jr_comm_share(ctx, P.S[q], P.peerS[q].data()))) return st;
                P.shared_ptr[q] = (void *)P.S[q];
            }
        }
        P.push = true;
    }
    if ((st = set_tma_map(&P.mS5[0], P.S[0], P, S_N, 5))) return st;
    if ((st = set_tma_map(&P.mS5[1], P.S[1], P, S_N, 5))) return st;
    if ((st = set_tma_map(&P.mC1, P.C, P, C_N, 1))) return st;
    if ((st = set_tma_map(&P.mC4, P.C, P, C_N, 4))) return st;
    if (fin) {
        if ((st = set_tma_map(&P.mD1, P.D, P, D_N, 1))) return st;
        if ((st = set_tma_map(&P.mD2, P.D, P, D_N, 2))) return st;
        if ((st = set_tma_map(&P.mD7, P.D, P, D_N, 7))) return st;
    }

    // pack the user's dense arrays into the box sets
    PackArgs pa;
    pa.njobs = 0;
    auto add = [&](double *dense, double *set, int a, int NAq, int n0, int n1, int n2, int o0, int o1, int o2, int mode) {
        PackJob &J = pa.j[pa.njobs++];
        J.dense = dense; J.box = box_arr(set, a, NAq, P);
        J.n[0] = n0; J.n[1] = n1; J.n[2] = n2; J.o[0] = o0; J.o[1] = o1; J.o[2] = o2; J.mode = mode;
    };
    add(F(Vx), P.S[0], S_Vx, S_N, nx + 1, ny + 2, nz + 2, 1, 0, 0, 0);
    add(F(Vy), P.S[0], S_Vy, S_N, nx + 2, ny + 1, nz + 2, 0, 1, 0, 0);
    add(F(Vz), P.S[0], S_Vz, S_N, nx + 2, ny + 2, nz + 1, 0, 0, 1, 0);
    add(F(P), P.S[0], S_P, S_N, nx, ny, nz, 1, 1, 1, 0);
    add(F(txx), P.S[0], S_txx, S_N, nx, ny, nz, 1, 1, 1, 0);
    add(F(tyy), P.S[0], S_tyy, S_N, nx, ny, nz, 1, 1, 1, 0);
    add(F(tzz), P.S[0], S_tzz, S_N, nx, ny, nz, 1, 1, 1, 0);
    add(F(tyz), P.S[0], S_tyz, S_N, nx, ny + 1, nz + 1, 1, 1, 1, 0);
    add(F(txz), P.S[0], S_txz, S_N, nx + 1, ny, nz + 1, 1, 1, 1, 0);
    add(F(txy), P.S[0], S_txy, S_N, nx + 1, ny + 1, nz, 1, 1, 1, 0);
    add(F(eta), P.C, C_eta, C_N, nx, ny, nz, 0, 0, 0, 2);
    add(F(etatau), P.C, C_ett, C_N, nx, ny, nz, 1, 1, 1, 0);
    add(F(rhogx), P.C, C_fx, C_N, nx, ny, nz, 1, 1, 1, 0);
    add(F(rhogy), P.C, C_fy, C_N, nx, ny, nz, 1, 1, 1, 0);
    add(F(rhogz), P.C, C_fz, C_N, nx, ny, nz, 1, 1, 1, 0);
    if (fin) {
        add(F(G), P.D, D_G, D_N, nx, ny, nz, 0, 0, 0, 2);
        add(F(K), P.D, D_K, D_N, nx, ny, nz, 1, 1, 1, 0);
        add(F(P0), P.D, D_P0, D_N, nx, ny, nz, 1, 1, 1, 0);
        add(F(Q), P.D, D_Q, D_N, nx, ny, nz, 1, 1, 1, 0);
        add(F(txx_o), P.D, D_oxx, D_N, nx, ny, nz, 1, 1, 1, 0);
        add(F(tyy_o), P.D, D_oyy, D_N, nx, ny, nz, 1, 1, 1, 0);
        add(F(tzz_o), P.D, D_ozz, D_N, nx, ny, nz, 1, 1, 1, 0);
        add(F(tyz_o), P.D, D_oyz, D_N, nx, ny + 1, nz + 1, 1, 1, 1, 0);
        add(F(txz_o), P.D, D_oxz, D_N, nx + 1, ny, nz + 1, 1, 1, 1, 0);
        add(F(txy_o), P.D, D_oxy, D_N, nx + 1, ny + 1, nz, 1, 1, 1, 0);
    }
    if ((st = run_pack(ctx, pa, (ny + 2) * (nz + 2)))) return st;
    return JR_OK;
}

#define JR_TRAIL_NA 7777  // internal: this plan leaves no CTA slot for trailing CTAs (or runs z-chunks): use the two-launch iteration
template <int BY, bool FIN, bool DG, int NSTv, bool RHOG, bool MULTI = false, bool PUSH = false, bool TRAIL = false>
static int launch_one(jr_context *ctx, VaPlan &P, VaArgs &a)
{
    constexpr int TY = BY - 2;
    constexpr int smem = NSTv * SlotMap<FIN, RHOG>::NARR * 32 * BY * 8;
    static int cta_per_sm = 0;
    if (!cta_per_sm) {
        JR_CUDA(cudaFuncSetAttribute(k_va_tma<BY, FIN, DG, NSTv, RHOG, MULTI, PUSH, TRAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        JR_CUDA(cudaFuncSetAttribute(k_va_tma<BY, FIN, DG, NSTv, RHOG, MULTI, PUSH, TRAIL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        int nb = 0;
        JR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_va_tma<BY, FIN, DG, NSTv, RHOG, MULTI, PUSH, TRAIL>, 32 * BY, smem));
        JR_REQUIRE(nb >= 1, JR_ERR_CUDA, "k_va_tma<%d> does not fit on an SM (%d B shared memory)", BY, smem);
        cta_per_sm = nb;
        if (getenv("JRB200_VERBOSE"))
            fprintf(stderr, "[jrb200] k_va_tma<BY=%d,finite_dt=%d,diag=%d,stages=%d>: %d B smem, %d CTA/SM, nchunk=%d slack=%d\n",
                    BY, (int)FIN, (int)DG, NSTv, smem, nb, P.nchunk, P.slack);
    }
    a.ntx = (P.nx + 1 + TXW - 1) / TXW;
    a.nty = (P.ny + 1 + TY - 1) / TY;
    a.nchunk = (P.nz + a.kchunk - 1) / a.kchunk;
    const long items = (long)a.ntx * a.nty * a.nchunk;
    const long slots = (long)ctx->sm_count * cta_per_sm;
    // multi-GPU overlap: the direct exchange of this iteration's velocities runs UNDER the next launch on a second stream.  A cooperative
    // launch never shares the GPU with another kernel (the driver serialises it: measured, the exchange then simply runs first), so
    // these launches are ordinary ones; ovl_ctas CTA slots stay free for the exchange, and the soft lock-step no longer assumes
    // co-residency (spin_cap)
    const bool coop = !P.ovl;
    const long reserve = coop ? 0 : P.ovl_ctas;
    const long avail = slots - reserve > 1 ? slots - reserve : 1;
    int G = (int)(items < avail ? items : avail);
    a.spin_cap = coop ? 0 : 4096;
    if (TRAIL) {
        // every column tile must be resident at once (one round, one z-chunk) with at least one CTA slot left for the trailing CTAs
        if (!coop || a.nchunk != 1 || items + 1 > slots) return JR_TRAIL_NA;
        const long nt = slots - items < P.trail_max ? slots - items : P.trail_max;
        a.tr.ntrail = (int)nt;
        a.tr.done = P.done;
        a.tr.done_base = P.done_base;
        P.done_base += (unsigned long long)(a.kchunk + 2);
        G = (int)(items + nt);
    }
    a.progress = P.progress;
    a.progress_base = P.progress_base;
    a.slack = P.slack;
    a.stagger_ns = P.stagger_ns;
    const int nit = MULTI ? a.niter : 1;
    P.progress_base += (unsigned long long)nit * items * (a.kchunk + 2);  // every item posts kchunk+2 steps per iteration
    a.gbar = P.progress + 16;  // its own 128-B line
    a.gbar_base = P.gbar_base;
    P.gbar_base += (unsigned long long)(nit - 1) * G;
    void *args[1] = {(void *)&a};
    // cooperative launch: the soft lock-step spins on other CTAs, so all G CTAs must be resident
    if (coop)
        JR_CUDA(cudaLaunchCooperativeKernel((const void *)k_va_tma<BY, FIN, DG, NSTv, RHOG, MULTI, PUSH, TRAIL>, dim3(G, 1, 1), dim3(32, BY, 1), args,
                                            smem, ctx->stream));
    else
        JR_CUDA(cudaLaunchKernel((const void *)k_va_tma<BY, FIN, DG, NSTv, RHOG, MULTI, PUSH, TRAIL>, dim3(G, 1, 1), dim3(32, BY, 1), args, smem,
                                 ctx->stream));
    return JR_OK;
}

template <bool RHOG, bool PUSH>
static int launch_va_t(jr_context *ctx, VaPlan &P, VaArgs &a, int diag)
{
    if (P.finite_dt)
        return diag ? launch_one<8, true, true, 2, RHOG, false, PUSH>(ctx, P, a) : launch_one<8, true, false, 2, RHOG, false, PUSH>(ctx, P, a);
    switch (P.BY) {
    case 8: return diag ? launch_one<8, false, true, 3, RHOG, false, PUSH>(ctx, P, a) : launch_one<8, false, false, 3, RHOG, false, PUSH>(ctx, P, a);
    case 10: return diag ? launch_one<10, false, true, 3, RHOG, false, PUSH>(ctx, P, a) : launch_one<10, false, false, 3, RHOG, false, PUSH>(ctx, P, a);
    default: return diag ? launch_one<16, false, true, 3, RHOG, false, PUSH>(ctx, P, a) : launch_one<16, false, false, 3, RHOG, false, PUSH>(ctx, P, a);
    }
}
static int launch_va(jr_context *ctx, VaPlan &P, VaArgs &a, int diag)
{
    if (P.push) return P.rhog_const ? launch_va_t<false, true>(ctx, P, a, diag) : launch_va_t<true, true>(ctx, P, a, diag);
    return P.rhog_const ? launch_va_t<false, false>(ctx, P, a, diag) : launch_va_t<true, false>(ctx, P, a, diag);
}
// flow_bcs! by trailing CTAs of the same launch (never with diagnostics); JR_TRAIL_NA: not applicable to this plan
template <bool RHOG>
static int launch_va_trail_t(jr_context *ctx, VaPlan &P, VaArgs &a)
{
    if (P.finite_dt) return launch_one<8, true, false, 2, RHOG, false, false, true>(ctx, P, a);
    switch (P.BY) {
    case 8: return launch_one<8, false, false, 3, RHOG, false, false, true>(ctx, P, a);
    case 10: return launch_one<10, false, false, 3, RHOG, false, false, true>(ctx, P, a);
    default: return launch_one<16, false, false, 3, RHOG, false, false, true>(ctx, P, a);
    }
}
// several iterations per launch, boundary conditions inside the kernel (never with diagnostics)
template <bool RHOG>
static int launch_va_multi_t(jr_context *ctx, VaPlan &P, VaArgs &a)
{
    if (P.finite_dt) return launch_one<8, true, false, 2, RHOG, true>(ctx, P, a);
    switch (P.BY) {
    case 8: return launch_one<8, false, false, 3, RHOG, true>(ctx, P, a);
    case 10: return launch_one<10, false, false, 3, RHOG, true>(ctx, P, a);
    default: return launch_one<16, false, false, 3, RHOG, true>(ctx, P, a);
    }
}

static void fill_args(VaArgs &a, const VaPlan &P, const jr_fields *s, const jr_stokes_opts *o, int parity)
{
    const int nx = P.nx, ny = P.ny, nz = P.nz;
    a.mS5 = P.mS5[parity ? 1 : 0]; a.mC1 = P.mC1; a.mC4 = P.mC4;
    if (P.finite_dt) { a.mD1 = P.mD1; a.mD2 = P.mD2; a.mD7 = P.mD7; }
    else { a.mD1 = P.mC1; a.mD2 = P.mC1; a.mD7 = P.mC1; }
    a.out = P.S[parity ? 0 : 1];
    a.mS5b = P.mS5[parity ? 0 : 1];
    a.out2 = P.S[parity ? 1 : 0];
    a.niter = 1;
    a.gbar = nullptr; a.gbar_base = 0;
    a.divV = F(divV); a.RP = F(RP); a.exx = F(exx); a.eyy = F(eyy); a.ezz = F(ezz); a.eyz = F(eyz); a.exz = F(exz); a.exy = F(exy);
    a.Rx = F(Rx); a.Ry = F(Ry); a.Rz = F(Rz); a.Ux = F(Ux); a.Uy = F(Uy); a.Uz = F(Uz);
    a.dVx = F(Vx); a.dVy = F(Vy); a.dVz = F(Vz); a.dP = F(P); a.dtxx = F(txx); a.dtyy = F(tyy); a.dtzz = F(tzz);
    a.dtyz = F(tyz); a.dtxz = F(txz); a.dtxy = F(txy);
    a.nx = nx; a.ny = ny; a.nz = nz; a.PX = P.PX; a.PY = P.PY;
    a.kchunk = (nz + P.nchunk - 1) / P.nchunk;
    a.pol_ld = P.pol_ld; a.pol_st = P.pol_st;
    a.fxc = P.fc[0]; a.fyc = P.fc[1]; a.fzc = P.fc[2];
    a._dx = o->_di[0]; a._dy = o->_di[1]; a._dz = o->_di[2]; a.dt = o->dt; a.r = o->r; a.theta_dtau = o->theta_dtau;
    a.eta_dtau = o->eta_dtau;
    // flags: left,right,front,back,top,bot.  no_slip: bot → z lo, top → z hi; free_slip (Q2): top → z lo, bot → z hi
    const int32_t *fs = o->free_slip, *ns = o->no_slip;
    const int fsl[6] = {fs[0], fs[1], fs[2], fs[3], fs[4], fs[5]};
    const int nsl[6] = {ns[0], ns[1], ns[2], ns[3], ns[5], ns[4]};
    for (int q = 0; q < 6; q++) {
        a.bc_nsn[q] = nsl[q] ? 1 : 0;
        a.bc_sg[q] = fsl[q] ? 1.0 : -1.0;
    }
    memset(&a.push, 0, sizeof(a.push));
    a.halo_prog = P.ovl_pending ? P.hprog : nullptr;
    a.halo_base = P.hbase; a.halo_nslots = P.ovl_ctas; a.halo_head = P.ovl_head; a.halo_pz = P.PZ;
    a.halo_faces = P.halo_faces;
    memset(&a.tr, 0, sizeof(a.tr));
    a.tr.in = P.S[parity ? 1 : 0];
    a.tr.nb = P.trail_nb; a.tr.nb_tail = P.trail_nb_tail; a.tr.tail_planes = P.trail_tail; a.tr.sig_every = P.trail_sig;
    a.tr.fs_lo[0] = fs[0]; a.tr.fs_hi[0] = fs[1]; a.tr.ns_lo[0] = ns[0]; a.tr.ns_hi[0] = ns[1];
    a.tr.fs_lo[1] = fs[2]; a.tr.fs_hi[1] = fs[3]; a.tr.ns_lo[1] = ns[2]; a.tr.ns_hi[1] = ns[3];
    a.tr.fs_lo[2] = fs[4]; a.tr.fs_hi[2] = fs[5]; a.tr.ns_lo[2] = ns[5]; a.tr.ns_hi[2] = ns[4];
}

// neighbour tables of the push exchange for the iteration that writes set `outq`
static void fill_push(PushArgs &pu, const jr_context *ctx, const VaPlan &P, int outq)
{
    const jr_comm *cm = ctx->comm;
    memset(&pu, 0, sizeof(pu));
    pu.rank = cm->rank;
    pu.sig_mine = cm->dev.sig[cm->rank];
    for (int t = 0; t < 27; t++) {
        const int nb = cm->dev.nbr[t];
        pu.nbr_rank[t] = nb;
        if (nb >= 0 && nb != cm->rank) {
            pu.peer_out[t] = (double *)P.peerS[outq][nb];
            pu.sig_peer[t] = cm->dev.sig[nb];
        }
    }
    for (int d = 0; d < 3; d++) { pu.has_lo[d] = cm->has_lo[d]; pu.has_hi[d] = cm->has_hi[d]; }
    pu.delta[0] = (long)(P.nx - 2);
    pu.delta[1] = (long)(P.ny - 2) * P.PX;
    pu.delta[2] = (long)(P.nz - 2) * S_N * (long)P.pxy;
}

// can the kernel apply flow_bcs! itself?  every side needs a free-slip or no-slip flag (tangential ghosts are then
// images of interior values); single rank only (the halo exchange sits between iterations)
static bool multi_supported(const jr_context *ctx, const jr_stokes_opts *o)
{
    if (ctx->comm) return false;
    // opt-in (JRB200_VA_MULTI=1).  Measured on the pool's B200s (profiles/r01_va_multi_experiment.md): the kernel runs at the
    // 1000 W board power cap, so the launch gaps this removes are not on the critical path (the SM clock recovers in
    // them), and the in-kernel boundary code costs the lock-stepped grid more than the separate 16 µs BC kernel.
    const char *e = getenv("JRB200_VA_MULTI");
    if (!e || atoi(e) == 0) return false;
    const int32_t *fs = o->free_slip, *ns = o->no_slip;
    const int fsl[6] = {fs[0], fs[1], fs[2], fs[3], fs[4], fs[5]};
    const int nsl[6] = {ns[0], ns[1], ns[2], ns[3], ns[5], ns[4]};
    for (int q = 0; q < 6; q++)
        if (!fsl[q] && !nsl[q]) return false;
    return true;
}

// how many iterations one launch may run (0: not supported here, use jr_stokes3d_VA_fused_iter)
int jr_stokes3d_VA_fused_multi_max(jr_context *ctx, const jr_stokes_opts *o)
{
    if (!multi_supported(ctx, o)) return 0;
    int cap = 256;
    if (const char *e = getenv("JRB200_VA_MULTI_MAX")) cap = atoi(e);
    return cap < 1 ? 0 : cap;
}

// `niter` iterations without diagnostics in ONE launch, starting from set `parity` (0: S0 holds the state)
int jr_stokes3d_VA_fused_multi(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int niter, int parity)
{
    auto it = g_plans.find(ctx);
    JR_REQUIRE(it != g_plans.end() && it->second.S[0], JR_ERR_ARG, "fused iteration without jr_stokes3d_VA_fused_begin");
    JR_REQUIRE(niter >= 1 && multi_supported(ctx, o), JR_ERR_ARG, "multi-iteration launch not supported for this setup");
    VaPlan &P = it->second;
    VaArgs a;
    fill_args(a, P, s, o, parity);
    a.niter = niter;
    int st = P.rhog_const ? launch_va_multi_t<false>(ctx, P, a) : launch_va_multi_t<true>(ctx, P, a);
    if (st) return st;
    P.last_diag = false;
    ctx->launches += 1;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

// `parity` 0: set S0 → S1, 1: S1 → S0.
int jr_stokes3d_VA_fused_iter(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int diag, int parity)
{
    auto it = g_plans.find(ctx);
    JR_REQUIRE(it != g_plans.end() && it->second.S[0], JR_ERR_ARG, "fused iteration without jr_stokes3d_VA_fused_begin");
    VaPlan &P = it->second;
    const int nx = P.nx, ny = P.ny, nz = P.nz;
    double *in = P.S[parity ? 1 : 0], *out = P.S[parity ? 0 : 1];

    VaArgs a;
    fill_args(a, P, s, o, parity);
    if (P.push) {
        fill_push(a.push, ctx, P, parity ? 0 : 1);
        a.push.epoch = ++ctx->comm->push_epoch;
    }
    int st = JR_TRAIL_NA;
    if (!diag && P.trail && !P.push) {
        // one launch: the boundary conditions trail the z-march on the CTA slots the column tiles leave free
        st = P.rhog_const ? launch_va_trail_t<false>(ctx, P, a) : launch_va_trail_t<true>(ctx, P, a);
        if (st && st != JR_TRAIL_NA) return st;
    }
    const bool trailed = st == JR_OK;
    if (!trailed && (st = launch_va(ctx, P, a, diag))) return st;
    P.last_diag = diag != 0;

    BcArgsB b;
    b.A[0] = BcArrB{box_arr(in, S_Vx, S_N, P), box_arr(out, S_Vx, S_N, P), F(Ux), F(Vx), {nx + 1, ny + 2, nz + 2}, {1, 0, 0}, 0, (long)S_Vx * (long)P.pxy};
    b.A[1] = BcArrB{box_arr(in, S_Vy, S_N, P), box_arr(out, S_Vy, S_N, P), F(Uy), F(Vy), {nx + 2, ny + 1, nz + 2}, {0, 1, 0}, 1, (long)S_Vy * (long)P.pxy};
    b.A[2] = BcArrB{box_arr(in, S_Vz, S_N, P), box_arr(out, S_Vz, S_N, P), F(Uz), F(Vz), {nx + 2, ny + 2, nz + 1}, {0, 0, 1}, 2, (long)S_Vz * (long)P.pxy};
    // flags: left,right,front,back,top,bot.  no_slip: bot → z lo, top → z hi; free_slip (Q2): top → z lo, bot → z hi
    const int32_t *fs = o->free_slip, *ns = o->no_slip;
    b.lo_fs[0] = fs[0]; b.hi_fs[0] = fs[1]; b.lo_ns[0] = ns[0]; b.hi_ns[0] = ns[1];
    b.lo_fs[1] = fs[2]; b.hi_fs[1] = fs[3]; b.lo_ns[1] = ns[2]; b.hi_ns[1] = ns[3];
    b.lo_fs[2] = fs[4]; b.hi_fs[2] = fs[5]; b.lo_ns[2] = ns[5]; b.hi_ns[2] = ns[4];
    b.diag = diag; b.dt = o->dt;
    b.do_push = P.push ? 1 : 0;
    if (P.push) b.push = a.push;
    else memset(&b.push, 0, sizeof(b.push));
    b.ncell[0] = nx; b.ncell[1] = ny; b.ncell[2] = nz;
    for (int d = 0; d < 3; d++) {
        const bool mg = ctx->comm && ctx->comm->active && !P.push;
        b.skip_lo[d] = mg && ctx->comm->has_lo[d];
        b.skip_hi[d] = mg && ctx->comm->has_hi[d];
    }
    // update_halo!(Vx, Vy, Vz)  Stokes3D.jl:120, first half: the BC launch also packs the send planes (default exchange)
    const bool mg = ctx->comm && ctx->comm->active && !P.push;
    jr_harr H[3];
    HaloArgs hh;
    b.pack = 0; b.stage = nullptr; b.stage_off[0] = b.stage_off[1] = b.stage_off[2] = 0;
    if (mg) {
        for (int q = 0; q < 3; q++) {
            const BcArrB &A = b.A[q];
            H[q].p = A.out.p; H[q].sy = A.out.sy; H[q].sz = A.out.sz;
            for (int d = 0; d < 3; d++) { H[q].n[d] = A.n[d]; H[q].o[d] = A.o[d]; }
            H[q].ol[0] = 2 + A.n[0] - nx; H[q].ol[1] = 2 + A.n[1] - ny; H[q].ol[2] = 2 + A.n[2] - nz;
        }
        // the exchange of the previous iteration (second stream, opt-in overlap) has fed the kernel above; formally joined here,
        // before the flag barrier below tells the peers that this rank no longer reads their previous set
        if (P.ovl_pending) {
            JR_CUDA(cudaStreamWaitEvent(ctx->stream, P.ev_rest, 0));
            P.ovl_pending = false;
        }
        if (!trailed && !(P.ovl && !diag)) {
            if ((st = jr_comm_halo_begin(ctx, H, 3, &hh, &b.stage))) return st;
            b.pack = 1;
            for (int q = 0; q < 3; q++) b.stage_off[q] = hh.stage_off[q];
        }
    }
    int m = nx > ny ? nx : ny;
    m = (m > nz ? m : nz) + 2;
    dim3 bgrid((m + 31) / 32, (m + 8 * BC_ROWS - 1) / (8 * BC_ROWS), b.pack ? 36 : 18), bblock(32, 8, 1);
    if (!trailed) k_bc_box3<<<bgrid, bblock, 0, ctx->stream>>>(b);
    ctx->launches += trailed ? 1 : 2;
    JR_CHECK_LAUNCH();
    // update_halo!(Vx, Vy, Vz)  Stokes3D.jl:120, second half — pushed by the two kernels above (P.push), else flag barrier + pull
    if (mg) {
        if (b.pack) return jr_comm_halo_pull(ctx, &hh);
        if (!P.ovl || diag) return jr_comm_halo(ctx, H, 3);
        // head: flag barrier + the first planes (incl. the low z ghost plane) on the compute stream, with a full grid;
        // rest: the other planes in z order by a few CTAs on the second stream, under the next iteration's z-march
        P.hbase = (++P.ovl_count) * 65536ull;
        if ((st = jr_comm_halo_z(ctx, H, 3, P.PZ, P.ovl_head, P.ovl_chunk, P.ovl_ctas, P.ovl_stream, P.ev_head, P.ev_rest, P.hprog, P.hbase))) return st;
        P.ovl_pending = true;
    }
    return JR_OK;
}

// make the user's dense arrays current; `niter` = iterations done since begin.  An observable (DIAG) iteration has
// already written the whole state to the dense arrays (kernel + BC kernel), so after a solve — whose loop can only end
// on such an iteration — there is nothing to unpack; with a communicator attached the halo planes of V were exchanged
// on the box set afterwards, so V is still unpacked.
int jr_stokes3d_VA_fused_finish(jr_context *ctx, const jr_fields *s, int64_t niter)
{
    auto it = g_plans.find(ctx);
    JR_REQUIRE(it != g_plans.end() && it->second.S[0], JR_ERR_ARG, "fused finish without begin");
    VaPlan &P = it->second;
    const int nx = P.nx, ny = P.ny, nz = P.nz;
    double *cur = P.S[niter & 1];
    if (P.ovl_pending) {
        JR_CUDA(cudaStreamWaitEvent(ctx->stream, P.ev_rest, 0));
        P.ovl_pending = false;
    }
    if (P.last_diag && !ctx->comm && niter > 0) return JR_OK;
    if (P.push && niter > 0) {
        // the neighbours' pushes of the last iteration must have landed before the halo planes are read back
        PushArgs pu;
        fill_push(pu, ctx, P, 0);
        pu.epoch = ++ctx->comm->push_epoch;
        k_push_sync<<<1, 32, 0, ctx->stream>>>(pu);
        ctx->launches++;
        JR_CHECK_LAUNCH();
    }
    const bool v_only = P.last_diag && niter > 0;
    PackArgs pa;
    pa.njobs = 0;
    auto add = [&](double *dense, int a, int n0, int n1, int n2, int o0, int o1, int o2) {
        if (v_only && a != S_Vx && a != S_Vy && a != S_Vz) return;
        PackJob &J = pa.j[pa.njobs++];
        J.dense = dense; J.box = box_arr(cur, a, S_N, P);
        J.n[0] = n0; J.n[1] = n1; J.n[2] = n2; J.o[0] = o0; J.o[1] = o1; J.o[2] = o2; J.mode = 1;
    };
    add(F(Vx), S_Vx, nx + 1, ny + 2, nz + 2, 1, 0, 0);
    add(F(Vy), S_Vy, nx + 2, ny + 1, nz + 2, 0, 1, 0);
    add(F(Vz), S_Vz, nx + 2, ny + 2, nz + 1, 0, 0, 1);
    add(F(P), S_P, nx, ny, nz, 1, 1, 1);
    add(F(txx), S_txx, nx, ny, nz, 1, 1, 1);
    add(F(tyy), S_tyy, nx, ny, nz, 1, 1, 1);
    add(F(tzz), S_tzz, nx, ny, nz, 1, 1, 1);
    add(F(tyz), S_tyz, nx, ny + 1, nz + 1, 1, 1, 1);
    add(F(txz), S_txz, nx + 1, ny, nz + 1, 1, 1, 1);
    add(F(txy), S_txy, nx + 1, ny + 1, nz, 1, 1, 1);
    return run_pack(ctx, pa, (ny + 2) * (nz + 2));
}

// tiling / layout facts of the last fused plan of this context (benchmarks and tests report them)
extern "C" int jr_stokes3d_VA_plan_info(jr_context *ctx, int32_t info[8])
{
    JR_REQUIRE(ctx && info, JR_ERR_ARG, "null argument");
    auto it = g_plans.find(ctx);
    JR_REQUIRE(it != g_plans.end(), JR_ERR_ARG, "no fused plan on this context yet");
    const VaPlan &P = it->second;
    info[0] = P.BY; info[1] = P.nchunk; info[2] = P.rhog_const ? 1 : 0; info[3] = P.finite_dt ? 1 : 0;
    info[4] = P.PX; info[5] = P.PY; info[6] = P.PZ; info[7] = P.slack;
    return JR_OK;
}
