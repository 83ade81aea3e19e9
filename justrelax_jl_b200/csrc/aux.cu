// aux.cu — the per-time-step kernels the reference calls between the Stokes and the thermal PT loops, so that a coupled time
// step stays on the GPU:
//   velocity2vertex! / velocity2center!   src/Interpolations.jl:212-289
//   compute_lithostatic_pressure!         src/Utils.jl:541-617 (column integration; across ranks: peer-memory gather of the column
//                                          weights of the ranks stacked above, replacing MPI.Allgather on the vertical sub-communicator)
//   compute_shear_heating!                src/thermal_diffusion/ShearHeating.jl:14-72
#include "rheo.cuh"
#include "comm.cuh"

// ---------------------------------------------------------------------------------------------------------------------------------
// velocity2vertex!: out arrays of extents e[3] (the reference launches over size(Vx_v)); V arrays with their staggered extents
__global__ void k_vel2vertex3(int e0, int e1, int e2, int nx, int ny, int nz, double *__restrict__ Xv, double *__restrict__ Yv, double *__restrict__ Zv,
                              const double *__restrict__ Vx, const double *__restrict__ Vy, const double *__restrict__ Vz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > e0 || j > e1 || k > e2) return;
    const size_t o = IX3(e0, e1, i, j, k);
#define VX(a, b, c) Vx[IX3(nx + 1, ny + 2, a, b, c)]
#define VY(a, b, c) Vy[IX3(nx + 2, ny + 1, a, b, c)]
#define VZ(a, b, c) Vz[IX3(nx + 2, ny + 2, a, b, c)]
    Xv[o] = 0.25 * (VX(i, j, k) + VX(i, j + 1, k) + VX(i, j, k + 1) + VX(i, j + 1, k + 1));
    Yv[o] = 0.25 * (VY(i, j, k) + VY(i + 1, j, k) + VY(i, j, k + 1) + VY(i + 1, j, k + 1));
    Zv[o] = 0.25 * (VZ(i, j, k) + VZ(i, j + 1, k) + VZ(i + 1, j, k) + VZ(i + 1, j + 1, k));
}
__global__ void k_vel2center3(int e0, int e1, int e2, int nx, int ny, int nz, double *__restrict__ Xc, double *__restrict__ Yc, double *__restrict__ Zc,
                              const double *__restrict__ Vx, const double *__restrict__ Vy, const double *__restrict__ Vz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > e0 || j > e1 || k > e2) return;
    const size_t o = IX3(e0, e1, i, j, k);
    Xc[o] = (VX(i, j + 1, k + 1) + VX(i + 1, j + 1, k + 1)) / 2;
    Yc[o] = (VY(i + 1, j, k + 1) + VY(i + 1, j + 1, k + 1)) / 2;
    Zc[o] = (VZ(i + 1, j + 1, k) + VZ(i + 1, j + 1, k + 1)) / 2;
#undef VX
#undef VY
#undef VZ
}
__global__ void k_vel2vertex2(int e0, int e1, int nx, int ny, double *__restrict__ Xv, double *__restrict__ Yv, const double *__restrict__ Vx,
                              const double *__restrict__ Vy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > e0 || j > e1) return;
    Xv[IX2(e0, i, j)] = (Vx[IX2(nx + 1, i, j)] + Vx[IX2(nx + 1, i, j + 1)]) / 2;
    Yv[IX2(e0, i, j)] = (Vy[IX2(nx + 2, i, j)] + Vy[IX2(nx + 2, i + 1, j)]) / 2;
}
__global__ void k_vel2center2(int e0, int e1, int nx, int ny, double *__restrict__ Xc, double *__restrict__ Yc, const double *__restrict__ Vx,
                              const double *__restrict__ Vy)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > e0 || j > e1) return;
    Xc[IX2(e0, i, j)] = (Vx[IX2(nx + 1, i, j + 1)] + Vx[IX2(nx + 1, i + 1, j + 1)]) / 2;
    Yc[IX2(e0, i, j)] = (Vy[IX2(nx + 2, i + 1, j)] + Vy[IX2(nx + 2, i + 1, j + 1)]) / 2;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// compute_lithostatic_pressure!: P = reverse(cumsum(reverse(w))) − w/2 along the last dimension, w = ρg·dz.  One thread per column
// (ncol = product of the leading extents; consecutive threads → consecutive columns: coalesced), sequential from the top like the
// reference's cumsum.  `above` (may be NULL): weight of the cells held by the ranks stacked above, one value per column.
__global__ void k_litho_column(size_t ncol, int nz, double *__restrict__ P, const double *__restrict__ rhog, double dz, const double *__restrict__ dzv,
                               const double *__restrict__ above, int nshared, double *__restrict__ contrib)
{
    const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    double acc = 0.0, own = 0.0;
    bool first = true;
    const double off = above ? above[c] : 0.0;
    for (int k = nz - 1; k >= 0; k--) {
        const double w = rhog[(size_t)k * ncol + c] * (dzv ? dzv[k] : dz);
        acc = first ? w : acc + w;
        first = false;
        double p = acc - w / 2;
        if (above) p += off;
        P[(size_t)k * ncol + c] = p;
        if (contrib && k == nshared) own = acc;   // Σ_{k ≥ nshared} w: the cells the rank below does not hold
    }
    if (contrib) contrib[c] = own;
}
// offsets from the ranks above: each rank has published `contrib` in its staging buffer; sum the columns of the ranks with a larger
// vertical coordinate in rank order (bit-identical on every rank of a column)
__global__ void k_litho_above(const __grid_constant__ jr_comm_dev cd, unsigned long long epoch, int buf, size_t ncol, int nabove, const int *__restrict__ ranks_above,
                              double *__restrict__ above)
{
    jr_comm_barrier_dev(cd, epoch);
    const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    double acc = 0.0;
    for (int q = 0; q < nabove; q++) {
        const double v = cd.stage[ranks_above[q]][buf][c];
        acc = q == 0 ? v : acc + v;
    }
    above[c] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// compute_shear_heating!: H_s = max(0, Σ_phase ratio · Χ_phase · τij (εij − εij_el)), εij_el = ½ (τij − τij_o)/(G dt), G phase-mixed;
// τ, τ_o at the centres (shear: the `_c` copies), ε shear averaged from its staggered location (cache_tensors, StressUpdate.jl:190-301).
// The contraction runs over the full symmetric tensor (shear components twice).
struct ShArgs {
    int nd, nx, ny, nz;
    double dt;
    const double *txx, *tyy, *tzz, *tyz, *txz, *txy;         // centre arrays
    const double *oxx, *oyy, *ozz, *oyz, *oxz, *oxy;
    const double *exx, *eyy, *ezz, *eyz, *exz, *exy;         // normals at centres, shear staggered
    const double *ph_c;
    double chi[JR_MAX_PHASES];
    double *out;
};
__global__ void k_shear_heating(const __grid_constant__ ShArgs a, const __grid_constant__ jr_phase_tab pt)
{
    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i > nx || j > ny || k > nz) return;
    const size_t nc = (size_t)nx * ny * nz, c = IX3(nx, ny, i, j, k);
    double G = 0.0, chi = 0.0;
    if (a.ph_c) {
        G = jr_ratio_G(pt, a.ph_c, nc, c);
        for (int p = 0; p < pt.n; p++) {
            const double r = a.ph_c[(size_t)p * nc + c];
            chi += (r == 0.0) ? 0.0 : a.chi[p] * r;
        }
    } else {
        G = pt.G[0];
        chi = a.chi[0];
    }
    const double _Gdt = jr_inv(G * a.dt);
    double t[6], to[6], e[6];
    int n;
    if (a.nd == 2) {
        n = 3;
        t[0] = a.txx[c]; t[1] = a.tyy[c]; t[2] = a.txy[c];
        to[0] = a.oxx[c]; to[1] = a.oyy[c]; to[2] = a.oxy[c];
        e[0] = a.exx[c]; e[1] = a.eyy[c];
        e[2] = (((a.exy[IX2(nx + 1, i, j)] + a.exy[IX2(nx + 1, i + 1, j)]) + a.exy[IX2(nx + 1, i, j + 1)]) + a.exy[IX2(nx + 1, i + 1, j + 1)]) / 4;
    } else {
        n = 6;
        t[0] = a.txx[c]; t[1] = a.tyy[c]; t[2] = a.tzz[c]; t[3] = a.tyz[c]; t[4] = a.txz[c]; t[5] = a.txy[c];
        to[0] = a.oxx[c]; to[1] = a.oyy[c]; to[2] = a.ozz[c]; to[3] = a.oyz[c]; to[4] = a.oxz[c]; to[5] = a.oxy[c];
        e[0] = a.exx[c]; e[1] = a.eyy[c]; e[2] = a.ezz[c];
        // _av_yz / _av_xz / _av_xy  MiniKernels.jl:149-171: the four edges around the cell centre
        e[3] = 0.25 * (a.eyz[IX3(nx, ny + 1, i, j, k)] + a.eyz[IX3(nx, ny + 1, i, j + 1, k)] + a.eyz[IX3(nx, ny + 1, i, j, k + 1)] + a.eyz[IX3(nx, ny + 1, i, j + 1, k + 1)]);
        e[4] = 0.25 * (a.exz[IX3(nx + 1, ny, i, j, k)] + a.exz[IX3(nx + 1, ny, i + 1, j, k)] + a.exz[IX3(nx + 1, ny, i, j, k + 1)] + a.exz[IX3(nx + 1, ny, i + 1, j, k + 1)]);
        e[5] = 0.25 * (a.exy[IX3(nx + 1, ny + 1, i, j, k)] + a.exy[IX3(nx + 1, ny + 1, i + 1, j, k)] + a.exy[IX3(nx + 1, ny + 1, i, j + 1, k)] + a.exy[IX3(nx + 1, ny + 1, i + 1, j + 1, k)]);
    }
    const int nn = a.nd;   // number of normal components
    double H = 0.0;
    for (int q = 0; q < n; q++) {
        const double eel = 0.5 * ((t[q] - to[q]) * _Gdt);
        const double term = t[q] * (e[q] - eel);
        H += q < nn ? term : 2.0 * term;
    }
    a.out[c] = fmax(0.0, chi * H);
}

#define F(name) (s->f[JR_F_##name])
extern "C" {

int jr_velocity2vertex(jr_context *ctx, int32_t ndim, const int32_t n[3], const int32_t out_ext[3], double *Vx_v, double *Vy_v, double *Vz_v, const double *Vx,
                       const double *Vy, const double *Vz)
{
    JR_REQUIRE(ctx && n && out_ext && Vx_v && Vy_v && Vx && Vy, JR_ERR_ARG, "jr_velocity2vertex: null argument");
    JR_REQUIRE(ndim == 2 || (ndim == 3 && Vz_v && Vz), JR_ERR_ARG, "jr_velocity2vertex: ndim = %d", ndim);
    // reads Vx[i, j+1(, k+1)], Vy[i+1, j(, k+1)], Vz[i+1, j+1, k]: the output extents must stay inside the staggered arrays
    JR_REQUIRE(out_ext[0] <= n[0] + 1 && out_ext[1] <= n[1] + 1 && (ndim == 2 || out_ext[2] <= n[2] + 1), JR_ERR_SHAPE,
               "jr_velocity2vertex: output extents exceed the vertex grid");
    JR_CUDA(cudaSetDevice(ctx->device));
    dim3 blk(32, 8), grd((out_ext[0] + 31) / 32, (out_ext[1] + 7) / 8, ndim == 3 ? out_ext[2] : 1);
    if (ndim == 3) k_vel2vertex3<<<grd, blk, 0, ctx->stream>>>(out_ext[0], out_ext[1], out_ext[2], n[0], n[1], n[2], Vx_v, Vy_v, Vz_v, Vx, Vy, Vz);
    else k_vel2vertex2<<<grd, blk, 0, ctx->stream>>>(out_ext[0], out_ext[1], n[0], n[1], Vx_v, Vy_v, Vx, Vy);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

int jr_velocity2center(jr_context *ctx, int32_t ndim, const int32_t n[3], const int32_t out_ext[3], double *Vx_c, double *Vy_c, double *Vz_c, const double *Vx,
                       const double *Vy, const double *Vz)
{
    JR_REQUIRE(ctx && n && out_ext && Vx_c && Vy_c && Vx && Vy, JR_ERR_ARG, "jr_velocity2center: null argument");
    JR_REQUIRE(ndim == 2 || (ndim == 3 && Vz_c && Vz), JR_ERR_ARG, "jr_velocity2center: ndim = %d", ndim);
    JR_REQUIRE(out_ext[0] <= n[0] && out_ext[1] <= n[1] && (ndim == 2 || out_ext[2] <= n[2]), JR_ERR_SHAPE, "jr_velocity2center: output extents exceed the cell grid");
    JR_CUDA(cudaSetDevice(ctx->device));
    dim3 blk(32, 8), grd((out_ext[0] + 31) / 32, (out_ext[1] + 7) / 8, ndim == 3 ? out_ext[2] : 1);
    if (ndim == 3) k_vel2center3<<<grd, blk, 0, ctx->stream>>>(out_ext[0], out_ext[1], out_ext[2], n[0], n[1], n[2], Vx_c, Vy_c, Vz_c, Vx, Vy, Vz);
    else k_vel2center2<<<grd, blk, 0, ctx->stream>>>(out_ext[0], out_ext[1], n[0], n[1], Vx_c, Vy_c, Vx, Vy);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

int jr_lithostatic_pressure(jr_context *ctx, int32_t ndim, const int32_t n[3], double *P, const double *rhog, double dz, const double *dz_cells, int across_ranks,
                            int32_t ncell_vertical)
{
    JR_REQUIRE(ctx && n && P && rhog, JR_ERR_ARG, "jr_lithostatic_pressure: null argument");
    JR_REQUIRE(ndim == 2 || ndim == 3, JR_ERR_ARG, "jr_lithostatic_pressure: ndim = %d", ndim);
    JR_CUDA(cudaSetDevice(ctx->device));
    const int nz = n[ndim - 1];
    const size_t ncol = ndim == 2 ? (size_t)n[0] : (size_t)n[0] * n[1];
    const int vdim = ndim - 1;   // vertical dimension of the rank grid: IGG's dims[N]
    jr_comm *cm = ctx->comm;
    const int nvert = cm ? cm->dims[vdim] : 1;
    // the three-argument method throws when the vertical direction is split across ranks  Utils.jl:541-550
    JR_REQUIRE(across_ranks || nvert == 1, JR_ERR_ARG,
               "the vertical direction is split across MPI ranks; pass the `IGG` topology as fourth argument to integrate the columns across ranks");
    const unsigned blocks = (unsigned)((ncol + 255) / 256);
    if (!across_ranks || nvert == 1) {
        k_litho_column<<<blocks, 256, 0, ctx->stream>>>(ncol, nz, P, rhog, dz, dz_cells, nullptr, 0, nullptr);
        ctx->launches++;
        JR_CHECK_LAUNCH();
        return JR_OK;
    }
    // Utils.jl:580-583
    JR_REQUIRE(!cm->periods[vdim], JR_ERR_UNSUPPORTED, "the lithostatic pressure of a column that is periodic along the vertical direction is undefined");
    // cells at the bottom of the local column that the rank below also holds  Utils.jl:585-590
    const int nshared = nz - ncell_vertical + 2;
    JR_REQUIRE(nshared >= 0 && nshared < nz, JR_ERR_SHAPE,
               "a local column of %d cells cannot share %d cells with the rank below; `P` must be a field of the global grid", nz, nshared);
    // 1. my contribution (the cells the rank below does not hold) → my staging buffer; 2. barrier, sum the ranks above; 3. integrate with the offset
    int st = jr_comm_reserve_stage(ctx, ncol);
    if (st) return st;
    const unsigned long long epoch = ++cm->epoch;
    const int buf = (int)(epoch & 1);
    void *above_v = nullptr, *ranks_v = nullptr;
    if ((st = jr_ctx_scratch(ctx, "litho_above", ncol * sizeof(double), &above_v))) return st;
    if ((st = jr_ctx_scratch(ctx, "litho_ranks", JR_COMM_MAX_RANKS * sizeof(int), &ranks_v))) return st;
    int ranks_above[JR_COMM_MAX_RANKS], nabove = 0;
    for (int cz = cm->coords[vdim] + 1; cz < nvert; cz++) {
        int c3[3] = {cm->coords[0], cm->coords[1], cm->coords[2]};
        c3[vdim] = cz;
        ranks_above[nabove++] = (c3[0] * cm->dims[1] + c3[1]) * cm->dims[2] + c3[2];   // MPI_Cart_rank, row-major
    }
    JR_CUDA(cudaMemcpyAsync(ranks_v, ranks_above, sizeof(int) * (nabove ? nabove : 1), cudaMemcpyHostToDevice, ctx->stream));
    k_litho_column<<<blocks, 256, 0, ctx->stream>>>(ncol, nz, P, rhog, dz, dz_cells, nullptr, nshared, (double *)cm->stage_mine[buf]);
    k_litho_above<<<blocks, 256, 0, ctx->stream>>>(cm->dev, epoch, buf, ncol, nabove, (const int *)ranks_v, (double *)above_v);
    k_litho_column<<<blocks, 256, 0, ctx->stream>>>(ncol, nz, P, rhog, dz, dz_cells, (const double *)above_v, 0, nullptr);
    ctx->launches += 3;
    JR_CHECK_LAUNCH();
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_compute_shear_heating(jr_context *ctx, const jr_fields *s, const jr_vc_inputs *vc, const double *chi_host, double dt, double *shear_heating)
{
    JR_REQUIRE(ctx && s && vc && chi_host && shear_heating, JR_ERR_ARG, "jr_compute_shear_heating: null argument");
    JR_REQUIRE(s->ndim == 2 || s->ndim == 3, JR_ERR_SHAPE, "ndim = %d", s->ndim);
    jr_phase_tab pt;
    int st = jr_make_phase_tab(vc, &pt);
    if (st) return st;
    JR_REQUIRE(vc->ph_center || vc->nphase == 1, JR_ERR_ARG, "a multi-phase rheology needs the centre phase ratios");
    ShArgs a;
    memset(&a, 0, sizeof(a));
    a.nd = s->ndim; a.nx = s->n[0]; a.ny = s->n[1]; a.nz = s->ndim == 3 ? s->n[2] : 1; a.dt = dt;
    a.txx = F(txx); a.tyy = F(tyy); a.oxx = F(txx_o); a.oyy = F(tyy_o); a.exx = F(exx); a.eyy = F(eyy); a.exy = F(exy);
    a.txy = F(txy_c); a.oxy = F(txy_o_c);
    JR_REQUIRE(a.txx && a.tyy && a.txy && a.oxx && a.oyy && a.oxy && a.exx && a.eyy && a.exy, JR_ERR_SHAPE, "compute_shear_heating!: τ, τ_o (centre) and ε are required");
    if (s->ndim == 3) {
        a.tzz = F(tzz); a.tyz = F(tyz_c); a.txz = F(txz_c); a.ozz = F(tzz_o); a.oyz = F(tyz_o_c); a.oxz = F(txz_o_c); a.ezz = F(ezz); a.eyz = F(eyz); a.exz = F(exz);
        JR_REQUIRE(a.tzz && a.tyz && a.txz && a.ozz && a.oyz && a.oxz && a.ezz && a.eyz && a.exz, JR_ERR_SHAPE, "compute_shear_heating! 3D: missing tensor component");
    }
    a.ph_c = vc->ph_center;
    for (int p = 0; p < vc->nphase; p++) a.chi[p] = chi_host[p];
    a.out = shear_heating;
    JR_CUDA(cudaSetDevice(ctx->device));
    dim3 blk(32, 8), grd((a.nx + 31) / 32, (a.ny + 7) / 8, a.nz);
    k_shear_heating<<<grd, blk, 0, ctx->stream>>>(a, pt);
    ctx->launches++;
    JR_CHECK_LAUNCH();
    return JR_OK;
}

} // extern "C"
