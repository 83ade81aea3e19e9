/*
 * aux.c — CPU ORACLE (test infrastructure, NOT product code).
 * Restatement of the per-time-step kernels around the PT loops:
 *   velocity2vertex! / velocity2center!   src/Interpolations.jl:212-289
 *   compute_lithostatic_pressure!         src/Utils.jl:541-617 (single rank: `_integrate_column!`; the weight of the ranks above is
 *                                          passed in as `above`, one value per column, the way `_weight_above` returns it)
 *   compute_shear_heating!                src/thermal_diffusion/ShearHeating.jl:14-72 — GeoParams.jl's ConstantShearheating (third party,
 *                                          compat 0.7.19, not vendored) restated from its documented definition
 *                                          H_s = Χ·τij(εij − εij_el) over the full symmetric tensor; the reference's own test only
 *                                          asserts H_s ≥ 0 (test/test_shearheating2D.jl:246): parity unpinned at unit level.
 */
#include "jr_oracle.h"
#include "vc_common.h"
#define F(name) (s->f[ORC_F_##name])

void orc_velocity2vertex(int nd, const int32_t n[3], const int32_t e[3], double *Xv, double *Yv, double *Zv, const double *Vx, const double *Vy, const double *Vz)
{
    const int nx = n[0], ny = n[1];
    if (nd == 2) {
        for (int j = 1; j <= e[1]; j++)
            for (int i = 1; i <= e[0]; i++) {
                Xv[IX2(e[0], i, j)] = (Vx[IX2(nx + 1, i, j)] + Vx[IX2(nx + 1, i, j + 1)]) / 2;
                Yv[IX2(e[0], i, j)] = (Vy[IX2(nx + 2, i, j)] + Vy[IX2(nx + 2, i + 1, j)]) / 2;
            }
        return;
    }
#define VX(a, b, c) Vx[IX3(nx + 1, ny + 2, a, b, c)]
#define VY(a, b, c) Vy[IX3(nx + 2, ny + 1, a, b, c)]
#define VZ(a, b, c) Vz[IX3(nx + 2, ny + 2, a, b, c)]
    for (int k = 1; k <= e[2]; k++)
        for (int j = 1; j <= e[1]; j++)
            for (int i = 1; i <= e[0]; i++) {
                const size_t o = IX3(e[0], e[1], i, j, k);
                Xv[o] = 0.25 * (VX(i, j, k) + VX(i, j + 1, k) + VX(i, j, k + 1) + VX(i, j + 1, k + 1));
                Yv[o] = 0.25 * (VY(i, j, k) + VY(i + 1, j, k) + VY(i, j, k + 1) + VY(i + 1, j, k + 1));
                Zv[o] = 0.25 * (VZ(i, j, k) + VZ(i, j + 1, k) + VZ(i + 1, j, k) + VZ(i + 1, j + 1, k));
            }
}

void orc_velocity2center(int nd, const int32_t n[3], const int32_t e[3], double *Xc, double *Yc, double *Zc, const double *Vx, const double *Vy, const double *Vz)
{
    const int nx = n[0], ny = n[1];
    if (nd == 2) {
        for (int j = 1; j <= e[1]; j++)
            for (int i = 1; i <= e[0]; i++) {
                Xc[IX2(e[0], i, j)] = (Vx[IX2(nx + 1, i, j + 1)] + Vx[IX2(nx + 1, i + 1, j + 1)]) / 2;
                Yc[IX2(e[0], i, j)] = (Vy[IX2(nx + 2, i + 1, j)] + Vy[IX2(nx + 2, i + 1, j + 1)]) / 2;
            }
        return;
    }
    for (int k = 1; k <= e[2]; k++)
        for (int j = 1; j <= e[1]; j++)
            for (int i = 1; i <= e[0]; i++) {
                const size_t o = IX3(e[0], e[1], i, j, k);
                Xc[o] = (VX(i, j + 1, k + 1) + VX(i + 1, j + 1, k + 1)) / 2;
                Yc[o] = (VY(i + 1, j, k + 1) + VY(i + 1, j + 1, k + 1)) / 2;
                Zc[o] = (VZ(i + 1, j + 1, k) + VZ(i + 1, j + 1, k + 1)) / 2;
            }
#undef VX
#undef VY
#undef VZ
}

/* P .= reverse(cumsum(reverse(w, dims = N), dims = N), dims = N) .- w ./ 2 [.+ above]   Utils.jl:552-571 */
void orc_lithostatic_pressure(int nd, const int32_t n[3], double *P, const double *rhog, double dz, const double *dzv, const double *above)
{
    const int nz = n[nd - 1];
    const size_t ncol = nd == 2 ? (size_t)n[0] : (size_t)n[0] * n[1];
    for (size_t c = 0; c < ncol; c++) {
        double acc = 0.0;
        for (int k = nz - 1; k >= 0; k--) {
            const double w = rhog[(size_t)k * ncol + c] * (dzv ? dzv[k] : dz);
            acc = (k == nz - 1) ? w : acc + w;
            double p = acc - w / 2;
            if (above) p += above[c];
            P[(size_t)k * ncol + c] = p;
        }
    }
}

void orc_shear_heating(const orc_fields *s, const orc_vc_inputs *vc, const double *chi, double dt, double *out)
{
    const int nd = s->ndim, nx = s->n[0], ny = s->n[1], nz = nd == 3 ? s->n[2] : 1;
    const size_t nc = (size_t)nx * ny * nz;
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                const size_t c = IX3(nx, ny, i, j, k);
                double G, X = 0.0;
                if (vc->ph_center) {
                    G = ratio_G(vc, vc->ph_center, nc, c);
                    for (int p = 0; p < vc->nphase; p++) { const double r = vc->ph_center[(size_t)p * nc + c]; X += (r == 0.0) ? 0.0 : chi[p] * r; }
                } else { G = vc->phases[0].G; X = chi[0]; }
                const double _Gdt = orc_inv(G * dt);
                double t[6], to[6], e[6];
                int m;
                if (nd == 2) {
                    m = 3;
                    t[0] = F(txx)[c]; t[1] = F(tyy)[c]; t[2] = F(txy_c)[c];
                    to[0] = F(txx_o)[c]; to[1] = F(tyy_o)[c]; to[2] = F(txy_o_c)[c];
                    e[0] = F(exx)[c]; e[1] = F(eyy)[c];
                    e[2] = (((F(exy)[IX2(nx + 1, i, j)] + F(exy)[IX2(nx + 1, i + 1, j)]) + F(exy)[IX2(nx + 1, i, j + 1)]) + F(exy)[IX2(nx + 1, i + 1, j + 1)]) / 4;
                } else {
                    m = 6;
                    t[0] = F(txx)[c]; t[1] = F(tyy)[c]; t[2] = F(tzz)[c]; t[3] = F(tyz_c)[c]; t[4] = F(txz_c)[c]; t[5] = F(txy_c)[c];
                    to[0] = F(txx_o)[c]; to[1] = F(tyy_o)[c]; to[2] = F(tzz_o)[c]; to[3] = F(tyz_o_c)[c]; to[4] = F(txz_o_c)[c]; to[5] = F(txy_o_c)[c];
                    e[0] = F(exx)[c]; e[1] = F(eyy)[c]; e[2] = F(ezz)[c];
                    e[3] = 0.25 * (F(eyz)[IX3(nx, ny + 1, i, j, k)] + F(eyz)[IX3(nx, ny + 1, i, j + 1, k)] + F(eyz)[IX3(nx, ny + 1, i, j, k + 1)] + F(eyz)[IX3(nx, ny + 1, i, j + 1, k + 1)]);
                    e[4] = 0.25 * (F(exz)[IX3(nx + 1, ny, i, j, k)] + F(exz)[IX3(nx + 1, ny, i + 1, j, k)] + F(exz)[IX3(nx + 1, ny, i, j, k + 1)] + F(exz)[IX3(nx + 1, ny, i + 1, j, k + 1)]);
                    e[5] = 0.25 * (F(exy)[IX3(nx + 1, ny + 1, i, j, k)] + F(exy)[IX3(nx + 1, ny + 1, i + 1, j, k)] + F(exy)[IX3(nx + 1, ny + 1, i, j + 1, k)] + F(exy)[IX3(nx + 1, ny + 1, i + 1, j + 1, k)]);
                }
                double H = 0.0;
                for (int q = 0; q < m; q++) {
                    const double eel = 0.5 * ((t[q] - to[q]) * _Gdt);
                    const double term = t[q] * (e[q] - eel);
                    H += q < nd ? term : 2.0 * term;
                }
                out[c] = fmax(0.0, X * H);
            }
}
