// stokes3d.cu — host driver of the 3D Stokes PT loops of libjrb200.
//
// Replaces the host while-loop of `_solve!` (src/stokes/Stokes3D.jl:25-186, variant 3D-VA): same loop
// condition, same `nout` sampling of the residual norms, same normalisers (quirk Q4), same history
// vectors; the per-iteration body is one fused sm_100a kernel (stokes3d_fused.cu) or, with
// JR_FLAG_UNFUSED, the reference-structured kernel sequence (stokes3d_unfused.cu).
#include "common.cuh"
#include "comm.cuh"
#include <algorithm>

#define F(name) (s->f[JR_F_##name])

static int check_fields_VA(const jr_fields *s, const jr_stokes_opts *o)
{
    JR_REQUIRE(s && o, JR_ERR_ARG, "null fields/opts");
    JR_REQUIRE(s->ndim == 3, JR_ERR_SHAPE, "3D solver called with ndim=%d", s->ndim);
    JR_REQUIRE(s->n[0] >= 3 && s->n[1] >= 3 && s->n[2] >= 3, JR_ERR_SHAPE, "grid must be at least 3 cells per dimension");
    const int req[] = {JR_F_P, JR_F_P0, JR_F_divV, JR_F_Q, JR_F_Vx, JR_F_Vy, JR_F_Vz, JR_F_Ux, JR_F_Uy, JR_F_Uz,
                       JR_F_txx, JR_F_tyy, JR_F_tzz, JR_F_tyz, JR_F_txz, JR_F_txy,
                       JR_F_txx_o, JR_F_tyy_o, JR_F_tzz_o, JR_F_tyz_o, JR_F_txz_o, JR_F_txy_o,
                       JR_F_exx, JR_F_eyy, JR_F_ezz, JR_F_eyz, JR_F_exz, JR_F_exy,
                       JR_F_eta, JR_F_etatau, JR_F_Rx, JR_F_Ry, JR_F_Rz, JR_F_RP,
                       JR_F_rhogx, JR_F_rhogy, JR_F_rhogz, JR_F_K, JR_F_G};
    for (int q : req) JR_REQUIRE(s->f[q] != nullptr, JR_ERR_SHAPE, "required field '%s' is NULL", jr_field_name(q));
    JR_REQUIRE(o->nout >= 1, JR_ERR_ARG, "nout must be >= 1");
    return JR_OK;
}

// ητ = deepcopy(η); compute_maxloc!(ητ, η); update_halo!(ητ)   Stokes3D.jl:55-57
static int pre_VA(jr_context *ctx, const jr_fields *s)
{
    const int32_t w[3] = {1, 1, 1};
    int st = jr_launch_maxloc3d(ctx, F(etatau), F(eta), s->n, w);
    if (st) return st;
    const jr_harr H = jr_harr_dense(F(etatau), s->n, s->n);
    return jr_comm_halo(ctx, &H, 1);
}

static int one_iter_VA(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, bool fused, int write_diag, int64_t it)
{
    if (fused) return jr_stokes3d_VA_fused_iter(ctx, s, o, write_diag, (int)(it & 1));
    return jr_stokes3d_VA_unfused_iter(ctx, s, o);
}

// τ_o ← τ  (multi_copy!, Stokes3D.jl:172-173)
static int post_VA(jr_context *ctx, const jr_fields *s)
{
    const size_t nx = s->n[0], ny = s->n[1], nz = s->n[2];
    cudaStream_t st = ctx->stream;
    const size_t nc = nx * ny * nz * 8;
    JR_CUDA(cudaMemcpyAsync(F(txx_o), F(txx), nc, cudaMemcpyDeviceToDevice, st));
    JR_CUDA(cudaMemcpyAsync(F(tyy_o), F(tyy), nc, cudaMemcpyDeviceToDevice, st));
    JR_CUDA(cudaMemcpyAsync(F(tzz_o), F(tzz), nc, cudaMemcpyDeviceToDevice, st));
    JR_CUDA(cudaMemcpyAsync(F(tyz_o), F(tyz), nx * (ny + 1) * (nz + 1) * 8, cudaMemcpyDeviceToDevice, st));
    JR_CUDA(cudaMemcpyAsync(F(txz_o), F(txz), (nx + 1) * ny * (nz + 1) * 8, cudaMemcpyDeviceToDevice, st));
    JR_CUDA(cudaMemcpyAsync(F(txy_o), F(txy), (nx + 1) * (ny + 1) * nz * 8, cudaMemcpyDeviceToDevice, st));
    if (F(tyz_c) && F(tyz_o_c)) JR_CUDA(cudaMemcpyAsync(F(tyz_o_c), F(tyz_c), nc, cudaMemcpyDeviceToDevice, st));
    if (F(txz_c) && F(txz_o_c)) JR_CUDA(cudaMemcpyAsync(F(txz_o_c), F(txz_c), nc, cudaMemcpyDeviceToDevice, st));
    if (F(txy_c) && F(txy_o_c)) JR_CUDA(cudaMemcpyAsync(F(txy_o_c), F(txy_c), nc, cudaMemcpyDeviceToDevice, st));
    return JR_OK;
}

extern "C" {

// ---- iteration session: the state lives in the library's layout between calls ------------------------------------
struct VaSession {
    bool active = false, fused = false;
    jr_fields f;
    jr_stokes_opts o;
    int64_t iter = 0;
};
static std::map<jr_context *, VaSession> g_sessions;

// `niter` iterations of the open session; the last one observable (diagnostics + dense state written) when observe_last
static int session_run(jr_context *ctx, VaSession &S, int64_t niter, int observe_last)
{
    const jr_fields *s = &S.f;
    const jr_stokes_opts *o = &S.o;
    int st;
    // iterations whose diagnostics nobody reads may run several per launch (the kernel applies flow_bcs! itself)
    const int64_t multi = (S.fused && !(ctx->flags & JR_FLAG_DIAG_EVERY_ITER)) ? jr_stokes3d_VA_fused_multi_max(ctx, o) : 0;
    for (int64_t it = 0; it < niter;) {
        const int64_t nb = std::min<int64_t>(niter - (observe_last ? 1 : 0) - it, multi);
        if (nb >= 1 && multi >= 1) {
            if ((st = jr_stokes3d_VA_fused_multi(ctx, s, o, (int)nb, (int)(S.iter & 1)))) return st;
            it += nb; S.iter += nb;
            continue;
        }
        const int diag = (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) ? 1 : (observe_last && it == niter - 1);
        if ((st = one_iter_VA(ctx, s, o, S.fused, diag, S.iter))) return st;
        it++; S.iter++;
    }
    return JR_OK;
}

int jr_stokes3d_VA_begin(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    int st = check_fields_VA(s, o);
    if (st) return st;
    JR_CUDA(cudaSetDevice(ctx->device));
    VaSession &S = g_sessions[ctx];
    JR_REQUIRE(!S.active, JR_ERR_ARG, "jr_stokes3d_VA_begin: a session is already open on this context (call jr_stokes3d_VA_end)");
    // both paths are CUDA paths of this library; the fused kernel covers uniform grids with
    // free-slip/no-slip faces, everything else runs the reference-structured kernel sequence.
    S.fused = !(ctx->flags & JR_FLAG_UNFUSED) && jr_stokes3d_VA_fused_supported(s, o) == JR_OK;
    S.f = *s; S.o = *o; S.iter = 0;
    ctx->launches = 0;
    if ((st = pre_VA(ctx, s))) return st;
    if (S.fused && (st = jr_stokes3d_VA_fused_begin(ctx, s, o))) return st;
    S.active = true;
    return JR_OK;
}

int jr_stokes3d_VA_step(jr_context *ctx, int64_t niter, int observe_last, jr_stokes_result *res)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    auto it = g_sessions.find(ctx);
    JR_REQUIRE(it != g_sessions.end() && it->second.active, JR_ERR_ARG, "jr_stokes3d_VA_step without jr_stokes3d_VA_begin");
    JR_REQUIRE(niter >= 0, JR_ERR_ARG, "niter must be >= 0");
    JR_CUDA(cudaSetDevice(ctx->device));
    const int64_t l0 = ctx->launches;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    int st = session_run(ctx, it->second, niter, observe_last);
    if (st) return st;
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (res) {
        float ms = 0.f;
        JR_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        res->iter = it->second.iter;
        res->nhist = 0;
        res->err = NAN;
        res->time_s = ms * 1e-3;
        res->kernel_launches = ctx->launches - l0;
    }
    return JR_OK;
}

int jr_stokes3d_VA_end(jr_context *ctx)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    auto it = g_sessions.find(ctx);
    JR_REQUIRE(it != g_sessions.end() && it->second.active, JR_ERR_ARG, "jr_stokes3d_VA_end without jr_stokes3d_VA_begin");
    VaSession &S = it->second;
    S.active = false;
    int st;
    if (S.fused && (st = jr_stokes3d_VA_fused_finish(ctx, &S.f, S.iter))) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    return JR_OK;
}

int jr_stokes3d_iterate_VA(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, int64_t niter, jr_stokes_result *res)
{
    JR_REQUIRE(ctx, JR_ERR_ARG, "null context");
    JR_CUDA(cudaSetDevice(ctx->device));
    // the timed region of this call includes entering and leaving the TMA box layout (jr_stokes3d_VA_step times iterations only)
    cudaEvent_t e0, e1;
    JR_CUDA(cudaEventCreate(&e0));
    JR_CUDA(cudaEventCreate(&e1));
    JR_CUDA(cudaEventRecord(e0, ctx->stream));
    int st = jr_stokes3d_VA_begin(ctx, s, o);
    if (!st) {
        st = session_run(ctx, g_sessions[ctx], niter, 1);
        const int st2 = jr_stokes3d_VA_end(ctx);
        if (!st) st = st2;
    }
    if (!st) {
        cudaEventRecord(e1, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        if (res) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            res->iter = niter;
            res->nhist = 0;
            res->err = NAN;
            res->time_s = ms * 1e-3;
            res->kernel_launches = ctx->launches;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return st;
}

int jr_stokes3d_solve_VA(jr_context *ctx, const jr_fields *s, const jr_stokes_opts *o, jr_stokes_result *res)
{
    JR_REQUIRE(ctx && res, JR_ERR_ARG, "null context/result");
    int st = check_fields_VA(s, o);
    if (st) return st;
    JR_REQUIRE(res->err_evo1 && res->err_evo2 && res->norm_Rx && res->norm_Ry && res->norm_Rz && res->norm_divV, JR_ERR_ARG,
               "result history arrays must be provided");
    JR_CUDA(cudaSetDevice(ctx->device));
    // both paths are CUDA paths of this library; the fused kernel covers uniform grids with
    // free-slip/no-slip faces, everything else runs the reference-structured kernel sequence.
    const bool fused = !(ctx->flags & JR_FLAG_UNFUSED) && jr_stokes3d_VA_fused_supported(s, o) == JR_OK;
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    void *slots_v = nullptr;
    if ((st = jr_ctx_scratch(ctx, "norm_slots", 16 * sizeof(double), &slots_v))) return st;
    double *slots = (double *)slots_v;

    ctx->launches = 0;
    double err_it1 = 1.0, err = 1.0;
    int64_t iter = 0, cont = 0;
    if ((st = pre_VA(ctx, s))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (fused && (st = jr_stokes3d_VA_fused_begin(ctx, s, o))) return st;
    const int64_t multi = (fused && !(ctx->flags & JR_FLAG_DIAG_EVERY_ITER)) ? jr_stokes3d_VA_fused_multi_max(ctx, o) : 0;
    while (iter < 2 || (((err / err_it1) > o->eps_rel && err > o->eps_abs) && iter <= o->iterMax)) {
        // the iterations up to the next sample (or iterMax) whose diagnostics nobody reads: one launch.  The loop
        // condition cannot change in between (err is only updated at samples; iter stays ≤ iterMax).
        if (multi >= 1 && iter >= 1) {
            const int64_t nb = std::min<int64_t>(std::min<int64_t>(o->nout - 1 - (iter % o->nout), o->iterMax - iter), multi);
            if (nb >= 1) {
                if ((st = jr_stokes3d_VA_fused_multi(ctx, s, o, (int)nb, (int)(iter & 1)))) return st;
                iter += nb;
                continue;
            }
        }
        // diagnostics (∇V, ε, R, U) are only read at `nout` samples and after the loop; the loop can only
        // end right after a sample or once iter > iterMax, so writing them on those iterations reproduces
        // the reference's final state exactly.
        const int64_t next = iter + 1;
        const int diag = (ctx->flags & JR_FLAG_DIAG_EVERY_ITER) ? 1 : ((next % o->nout == 0) || next > o->iterMax || next < 2);
        if ((st = one_iter_VA(ctx, s, o, fused, diag, iter))) return st;
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            const int32_t nRx[3] = {nx - 1, ny, nz}, nRy[3] = {nx, ny - 1, nz}, nRz[3] = {nx, ny, nz - 1}, nP[3] = {nx, ny, nz};
            if ((st = jr_launch_sumsq(ctx, F(Rx), nRx, 1, slots + 0))) return st;
            if ((st = jr_launch_sumsq(ctx, F(Ry), nRy, 1, slots + 1))) return st;
            if ((st = jr_launch_sumsq(ctx, F(Rz), nRz, 1, slots + 2))) return st;
            if ((st = jr_launch_sumsq(ctx, F(RP), nP, 0, slots + 3))) return st;
            // norm_mpi: MPI.Allreduce(sum) of the local sums of squares (overlap cells counted twice, as in the reference)
            if ((st = jr_comm_allreduce_dev(ctx, slots, 4, 0))) return st;
            JR_CUDA(cudaMemcpyAsync(ctx->h_pinned, slots, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            JR_CUDA(cudaStreamSynchronize(ctx->stream));
            // normalisers: Stokes3D.jl:129-142 (‖R‖₂ / N, quirk Q4)
            const double gx = o->n_g[0], gy = o->n_g[1], gz = o->n_g[2];
            const double nrx = sqrt(ctx->h_pinned[0]) / ((gx - 2) * (gy - 1) * (gz - 1));
            const double nry = sqrt(ctx->h_pinned[1]) / ((gx - 1) * (gy - 2) * (gz - 1));
            const double nrz = sqrt(ctx->h_pinned[2]) / ((gx - 1) * (gy - 1) * (gz - 2));
            const double nrp = sqrt(ctx->h_pinned[3]) / (gx * gy * gz);
            res->norm_Rx[cont] = nrx; res->norm_Ry[cont] = nry; res->norm_Rz[cont] = nrz; res->norm_divV[cont] = nrp;
            err = fmax(fmax(nrx, nry), fmax(nrz, nrp));
            if (std::isnan(nrx) || std::isnan(nry) || std::isnan(nrz) || std::isnan(nrp)) err = NAN;
            res->err_evo1[cont] = err;
            res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), fmax(res->norm_Rz[0], res->norm_divV[0]));
            if (std::isnan(err)) {
                res->iter = iter; res->nhist = cont; res->err = err;
                if (fused) jr_stokes3d_VA_fused_finish(ctx, s, iter);
                cudaStreamSynchronize(ctx->stream);
                jr_set_error("NaN(s)");  // reference: isnan(err) && error("NaN(s)")  Stokes3D.jl:162
                return JR_ERR_NAN;
            }
        }
    }
    if (fused && (st = jr_stokes3d_VA_fused_finish(ctx, s, iter))) return st;
    JR_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    if ((st = post_VA(ctx, s))) return st;
    JR_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    JR_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->iter = iter; res->nhist = cont; res->err = err;
    res->time_s = ms * 1e-3;
    res->kernel_launches = ctx->launches;
    return JR_OK;
}

} // extern "C"
