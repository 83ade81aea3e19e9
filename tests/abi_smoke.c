/*
 * abi_smoke.c — libjrb200 driven from a plain C host: no Python, no torch, no CUDA headers.
 *
 * What a Julia `JustRelaxB200Ext` does through ccall (julia/ext/JustRelaxB200Ext.jl), spelled in C: create a context,
 * allocate every StokesArrays slot with jr_malloc, upload the host arrays of the SolVi3D setup with jr_memcpy_h2d
 * (test/test_stokes_solvi3D.jl:25-61 + miniapps/benchmarks/stokes3D/solvi/SolVi3D.jl:45-130 at 16^3), apply flow_bcs!,
 * run solve! (jr_stokes3d_solve_VA), download with jr_memcpy_d2h and check the reference test's own criterion:
 * norm_Rx[end] < 1e-8.  Exit status 0 = pass.
 *
 * Build / run (tests/test_abi_smoke.py does this on the GPU box):
 *   gcc -O1 -std=c11 -Iinclude tests/abi_smoke.c -o /tmp/abi_smoke -Ljustrelax_jl_b200 -ljrb200 -lm -Wl,-rpath,$PWD/justrelax_jl_b200
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jrb200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int st_ = (call);                                                                \
        if (st_ != JR_OK) {                                                              \
            fprintf(stderr, "%s failed: status %d: %s\n", #call, st_, jr_last_error()); \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

static int slot(const char *name)
{
    for (int i = 0; i < jr_field_count(); i++)
        if (strcmp(jr_field_name(i), name) == 0) return i;
    fprintf(stderr, "no field slot named %s\n", name);
    exit(2);
}

int main(void)
{
    enum { n = 16 };
    const int nx = n, ny = n, nz = n;
    const double lx = 10.0, ly = 10.0, lz = 10.0, rc = 1.0, deta = 1.0e-3, ebg = 1.0;
    const double dx = lx / nx, dy = ly / ny, dz = lz / nz;
    if (jr_abi_version() != JRB200_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }

    /* ---- host setup: viscosity with a weak sphere, 10 smoothing passes (SolVi3D.jl:9-45) ---- */
    const size_t nc = (size_t)nx * ny * nz;
    double *eta = malloc(nc * 8), *eta2 = malloc(nc * 8);
#define C3(i, j, k) ((size_t)(k) * ny * nx + (size_t)(j) * nx + (i))
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const double x = i * dx + 0.5 * dx - 0.5 * lx, y = j * dy + 0.5 * dy - 0.5 * ly, z = k * dz + 0.5 * dz - 0.5 * lz;
                eta[C3(i, j, k)] = sqrt(x * x + y * y + z * z) <= rc ? deta : 1.0;
            }
    for (int pass = 0; pass < 10; pass++) {
        memcpy(eta2, eta, nc * 8);
        for (int k = 1; k < nz - 1; k++)
            for (int j = 1; j < ny - 1; j++)
                for (int i = 1; i < nx - 1; i++) {
                    const double c = eta[C3(i, j, k)];
                    const double d2x = (eta[C3(i + 1, j, k)] - c) - (c - eta[C3(i - 1, j, k)]);
                    const double d2y = (eta[C3(i, j + 1, k)] - c) - (c - eta[C3(i, j - 1, k)]);
                    const double d2z = (eta[C3(i, j, k + 1)] - c) - (c - eta[C3(i, j, k - 1)]);
                    eta2[C3(i, j, k)] = c + 1.0 / 6.1 / 1.0 * (d2x + d2y + d2z);
                }
        double *t = eta; eta = eta2; eta2 = t;
    }
    /* pure-shear velocity (pureshear_bc!, src/boundaryconditions/pure_shear.jl:15-32): Vx = εbg x, Vy = εbg x(!), Vz = −εbg z */
    const size_t nVx = (size_t)(nx + 1) * (ny + 2) * (nz + 2), nVy = (size_t)(nx + 2) * (ny + 1) * (nz + 2), nVz = (size_t)(nx + 2) * (ny + 2) * (nz + 1);
    double *Vx = calloc(nVx, 8), *Vy = calloc(nVy, 8), *Vz = calloc(nVz, 8);
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 0; i <= nx; i++) Vx[((size_t)k * (ny + 2) + j) * (nx + 1) + i] = ebg * (i * dx);
    for (int k = 1; k <= nz; k++)
        for (int j = 0; j <= ny; j++)
            for (int i = 1; i <= nx; i++) Vy[((size_t)k * (ny + 1) + j) * (nx + 2) + i] = ebg * (j * dx);
    for (int k = 0; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) Vz[((size_t)k * (ny + 2) + j) * (nx + 2) + i] = -ebg * (k * dz);

    /* ---- device: context, every slot of StokesArrays the 3D-VA solve touches ---- */
    jr_context *ctx = NULL;
    CHECK(jr_context_create(0, NULL, &ctx));
    jr_fields f;
    memset(&f, 0, sizeof f);
    f.ndim = 3; f.n[0] = nx; f.n[1] = ny; f.n[2] = nz;
    const size_t nyz = (size_t)nx * (ny + 1) * (nz + 1), nxz = (size_t)(nx + 1) * ny * (nz + 1), nxy = (size_t)(nx + 1) * (ny + 1) * nz;
    struct { const char *name; size_t count; double fill; } alloc[] = {
        {"P", nc, 0}, {"P0", nc, 0}, {"divV", nc, 0}, {"Q", nc, 0}, {"RP", nc, 0}, {"eta", nc, 1}, {"etatau", nc, 0},
        {"Vx", nVx, 0}, {"Vy", nVy, 0}, {"Vz", nVz, 0}, {"Ux", nVx, 0}, {"Uy", nVy, 0}, {"Uz", nVz, 0},
        {"txx", nc, 0}, {"tyy", nc, 0}, {"tzz", nc, 0}, {"tyz", nyz, 0}, {"txz", nxz, 0}, {"txy", nxy, 0},
        {"txx_o", nc, 0}, {"tyy_o", nc, 0}, {"tzz_o", nc, 0}, {"tyz_o", nyz, 0}, {"txz_o", nxz, 0}, {"txy_o", nxy, 0},
        {"exx", nc, 0}, {"eyy", nc, 0}, {"ezz", nc, 0}, {"eyz", nyz, 0}, {"exz", nxz, 0}, {"exy", nxy, 0},
        {"Rx", (size_t)(nx - 1) * ny * nz, 0}, {"Ry", (size_t)nx * (ny - 1) * nz, 0}, {"Rz", (size_t)nx * ny * (nz - 1), 0},
        {"rhogx", nc, 0}, {"rhogy", nc, 0}, {"rhogz", nc, 0}, {"K", nc, INFINITY}, {"G", nc, 1.0},
    };
    for (size_t q = 0; q < sizeof alloc / sizeof alloc[0]; q++) {
        void *p = NULL;
        CHECK(jr_malloc(ctx, alloc[q].count * 8, &p));
        CHECK(jr_fill_f64(ctx, (double *)p, alloc[q].fill, alloc[q].count));
        f.f[slot(alloc[q].name)] = (double *)p;
    }
    CHECK(jr_memcpy_h2d(ctx, f.f[slot("eta")], eta, nc * 8));
    CHECK(jr_memcpy_h2d(ctx, f.f[slot("Vx")], Vx, nVx * 8));
    CHECK(jr_memcpy_h2d(ctx, f.f[slot("Vy")], Vy, nVy * 8));
    CHECK(jr_memcpy_h2d(ctx, f.f[slot("Vz")], Vz, nVz * 8));

    /* ---- PTStokesCoeffs(li, di; CFL = 1/√3) (src/types/stokes.jl:203-229), free slip on all faces, kwargs of the test ---- */
    jr_stokes_opts o;
    memset(&o, 0, sizeof o);
    const double Re = 3.0 * M_PI, r = 0.7, CFL = 1.0 / sqrt(3.0), ltau = fmin(lx, fmin(ly, lz)), Vpdtau = fmin(dx, fmin(dy, dz)) * CFL;
    o.r = r; o.theta_dtau = ltau * (r + 4.0 / 3.0) / (Re * Vpdtau); o.eta_dtau = Vpdtau * ltau / Re;
    o.eps_rel = 1.0e-6; o.eps_abs = 1.0e-12;
    o._di[0] = 1.0 / dx; o._di[1] = 1.0 / dy; o._di[2] = 1.0 / dz;
    o.dt = INFINITY; o.iterMax = 5000; o.nout = 100;
    o.n_g[0] = nx; o.n_g[1] = ny; o.n_g[2] = nz;
    for (int q = 0; q < 6; q++) o.free_slip[q] = 1;
    o.viscosity_relaxation = 1.0e-2; o.lambda_relaxation = 0.2; o.visc_cutoff_lo = -INFINITY; o.visc_cutoff_hi = INFINITY;
    const int32_t none[6] = {0, 0, 0, 0, 0, 0};
    CHECK(jr_flow_bcs3d(ctx, f.f[slot("Vx")], f.f[slot("Vy")], f.f[slot("Vz")], f.n, o.free_slip, none, none));

    enum { CAP = 5000 / 100 + 3 };
    double err_evo1[CAP], nRx[CAP], nRy[CAP], nRz[CAP], nDiv[CAP];
    int64_t err_evo2[CAP];
    jr_stokes_result res;
    memset(&res, 0, sizeof res);
    res.err_evo1 = err_evo1; res.err_evo2 = err_evo2; res.norm_Rx = nRx; res.norm_Ry = nRy; res.norm_Rz = nRz; res.norm_divV = nDiv;
    CHECK(jr_stokes3d_solve_VA(ctx, &f, &o, &res));
    if (res.nhist < 1) { fprintf(stderr, "no residual sample recorded\n"); return 1; }
    const double last = nRx[res.nhist - 1];

    /* ---- download and sanity-check the solution ---- */
    double *P = malloc(nc * 8), *Vxo = malloc(nVx * 8);
    CHECK(jr_memcpy_d2h(ctx, P, f.f[slot("P")], nc * 8));
    CHECK(jr_memcpy_d2h(ctx, Vxo, f.f[slot("Vx")], nVx * 8));
    double pmax = 0.0, vmax = 0.0;
    int finite = 1;
    for (size_t q = 0; q < nc; q++) { if (!isfinite(P[q])) finite = 0; pmax = fmax(pmax, fabs(P[q])); }
    for (size_t q = 0; q < nVx; q++) { if (!isfinite(Vxo[q])) finite = 0; vmax = fmax(vmax, fabs(Vxo[q])); }

    /* ---- the iteration session: 3 iterations, the last one observable ---- */
    jr_stokes_result r2;
    memset(&r2, 0, sizeof r2);
    CHECK(jr_stokes3d_VA_begin(ctx, &f, &o));
    CHECK(jr_stokes3d_VA_step(ctx, 3, 1, &r2));
    CHECK(jr_stokes3d_VA_end(ctx));

    for (size_t q = 0; q < sizeof alloc / sizeof alloc[0]; q++) CHECK(jr_free(ctx, f.f[slot(alloc[q].name)]));
    CHECK(jr_context_destroy(ctx));
    printf("abi_smoke: SolVi3D %d^3 solve!: iter = %lld, norm_Rx[end] = %.3e, max|P| = %.4f, max|Vx| = %.4f, kernels launched = %lld, "
           "session: 3 iterations in %.3f ms\n",
           n, (long long)res.iter, last, pmax, vmax, (long long)res.kernel_launches, r2.time_s * 1e3);
    if (!(last < 1.0e-8)) { fprintf(stderr, "FAIL: norm_Rx[end] = %g is not < 1e-8\n", last); return 1; }
    if (!finite || !(pmax > 0.0) || !(vmax > 9.9 && vmax < 10.5)) { fprintf(stderr, "FAIL: solution is not sane\n"); return 1; }
    if (r2.iter != 3 || !(r2.time_s > 0.0)) { fprintf(stderr, "FAIL: session result\n"); return 1; }
    return 0;
}
