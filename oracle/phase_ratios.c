/*
 * phase_ratios.c — CPU ORACLE (test infrastructure, NOT product code).
 * Restatement of the grid-based phase-ratio construction update_phase_ratios_{2,3}D! (src/phases/PhaseRatios.jl:21-78):
 *   phase_ratios_center_from_arrays_kernel!    :90-117    values ./ sum, clamp, threshold 1e-5, renormalise
 *   phase_ratios_vertex_from_arrays_kernel!    :134-212   (bi/tri)linear weights of the ≤ 2^nd cells around a vertex
 *   phase_ratios_face_from_arrays_kernel!      :232-296   the two cells sharing a face, weight 1/2 each
 *   phase_ratios_midpoint_from_arrays_kernel!  :328-396   the four cells around an edge midpoint (xy, yz, xz), weight 1/4 each
 * Output layout [phase][node] (the CellArray flattened).  Known answers: test/test_phase_ratios3D.jl:31-84, test/test_rheology.jl:444-500.
 */
#include "jr_oracle.h"

#define MAXP 16

static void finish(double *w, int N, double total_weight, double *out, size_t stride, size_t idx)
{
    for (int k = 0; k < N; k++) w[k] /= total_weight;
    for (int k = 0; k < N; k++) w[k] = fmin(fmax(w[k], 0.0), 1.0);
    double total = 0.0;
    for (int k = 0; k < N; k++) { w[k] = w[k] < 1.0e-5 ? 0.0 : w[k]; total += w[k]; }
    for (int k = 0; k < N; k++) out[(size_t)k * stride + idx] = w[k] / total;
}

/* xc[d], xv[d]: cell-centre / vertex coordinates (length n[d], n[d]+1).  Any output pointer may be NULL. */
void orc_phase_ratios_from_arrays(int nd, const int32_t n[3], int N, const double *const *ph, const double *const *xc, const double *const *xv,
                                  double *center, double *vertex, double *Vx, double *Vy, double *Vz, double *xy, double *yz, double *xz)
{
    const int nx = n[0], ny = n[1], nz = nd == 3 ? n[2] : 1;
#define PH(k, i, j, l) ph[k][IX3(nx, ny, i, j, l)]
    if (center) {
        const size_t st = (size_t)nx * ny * nz;
        for (int l = 1; l <= nz; l++)
            for (int j = 1; j <= ny; j++)
                for (int i = 1; i <= nx; i++) {
                    double v[MAXP], total = 0.0;
                    for (int k = 0; k < N; k++) { v[k] = PH(k, i, j, l); total = k == 0 ? v[k] : total + v[k]; }
                    for (int k = 0; k < N; k++) { v[k] = v[k] / total; v[k] = v[k] < 0.0 ? 0.0 : (v[k] > 1.0 ? 1.0 : v[k]); v[k] = v[k] < 1.0e-5 ? 0.0 : v[k]; }
                    double ft = 0.0;
                    for (int k = 0; k < N; k++) ft = k == 0 ? v[k] : ft + v[k];
                    for (int k = 0; k < N; k++) center[(size_t)k * st + IX3(nx, ny, i, j, l)] = v[k] / ft;
                }
    }
    if (vertex) {
        const int vx = nx + 1, vy = ny + 1, vz = nd == 3 ? nz + 1 : 1;
        const size_t st = (size_t)vx * vy * vz;
        const double dx = xv[0][1] - xv[0][0], dy = xv[1][1] - xv[1][0], dz = nd == 3 ? xv[2][1] - xv[2][0] : 1.0;   /* JustPIC.compute_dx(xvi) */
        for (int l = 1; l <= vz; l++)
            for (int j = 1; j <= vy; j++)
                for (int i = 1; i <= vx; i++) {
                    double w[MAXP], tw = 0.0;
                    for (int k = 0; k < N; k++) w[k] = 0.0;
                    for (int o1 = -1; o1 <= 0; o1++)
                        for (int o2 = -1; o2 <= 0; o2++)
                            for (int o3 = (nd == 3 ? -1 : 0); o3 <= 0; o3++) {
                                const int ic = i + o1, jc = j + o2, lc = nd == 3 ? l + o3 : 1;
                                if (!(1 <= ic && ic <= nx && 1 <= jc && jc <= ny && 1 <= lc && lc <= nz)) continue;
                                double weight;
                                if (nd == 2) {
                                    const double wx = fma(-fabs(xv[0][i - 1] - xc[0][ic - 1]), 1.0 / dx, 1.0);
                                    const double wy = fma(-fabs(xv[1][j - 1] - xc[1][jc - 1]), 1.0 / dy, 1.0);
                                    weight = wx * wy;
                                } else {
                                    weight = 1.0;
                                    weight *= (1.0 - fabs(xv[0][i - 1] - xc[0][ic - 1]) * (1.0 / dx));
                                    weight *= (1.0 - fabs(xv[1][j - 1] - xc[1][jc - 1]) * (1.0 / dy));
                                    weight *= (1.0 - fabs(xv[2][l - 1] - xc[2][lc - 1]) * (1.0 / dz));
                                }
                                tw += weight;
                                for (int k = 0; k < N; k++) w[k] += weight * PH(k, ic, jc, lc);
                            }
                    finish(w, N, tw, vertex, st, IX3(vx, vy, i, j, l));
                }
    }
    /* faces: staggered in one dimension */
    double *face[3] = {Vx, Vy, nd == 3 ? Vz : NULL};
    for (int dim = 0; dim < nd; dim++) {
        if (!face[dim]) continue;
        const int off[3] = {dim == 0, dim == 1, dim == 2};
        const int e[3] = {nx + off[0], ny + off[1], nz + (nd == 3 ? off[2] : 0)};
        const size_t st = (size_t)e[0] * e[1] * e[2];
        for (int l = 1; l <= e[2]; l++)
            for (int j = 1; j <= e[1]; j++)
                for (int i = 1; i <= e[0]; i++) {
                    double w[MAXP], tw = 0.0;
                    for (int k = 0; k < N; k++) w[k] = 0.0;
                    for (int side = 0; side <= 1; side++) {
                        const int ic = off[0] ? i - 1 + side : i, jc = off[1] ? j - 1 + side : j, lc = (nd == 3 && off[2]) ? l - 1 + side : l;
                        if (!(1 <= ic && ic <= nx && 1 <= jc && jc <= ny && 1 <= lc && lc <= nz)) continue;
                        tw += 0.5;
                        for (int k = 0; k < N; k++) w[k] += 0.5 * PH(k, ic, jc, lc);
                    }
                    finish(w, N, tw, face[dim], st, IX3(e[0], e[1], i, j, l));
                }
    }
    /* midpoints (3D): staggered in two dimensions */
    if (nd == 3) {
        double *mid[3] = {xy, yz, xz};
        const int offs[3][3] = {{1, 1, 0}, {0, 1, 1}, {1, 0, 1}};
        for (int m = 0; m < 3; m++) {
            if (!mid[m]) continue;
            const int *off = offs[m];
            const int e[3] = {nx + off[0], ny + off[1], nz + off[2]};
            const size_t st = (size_t)e[0] * e[1] * e[2];
            for (int l = 1; l <= e[2]; l++)
                for (int j = 1; j <= e[1]; j++)
                    for (int i = 1; i <= e[0]; i++) {
                        double w[MAXP], tw = 0.0;
                        for (int k = 0; k < N; k++) w[k] = 0.0;
                        for (int corner = 1; corner <= 4; corner++) {
                            const int first = corner <= 2 ? -1 : 0, second = (corner & 1) ? -1 : 0;
                            const int ic = i + off[0] * first, jo = off[0] == 1 ? second : first, jc = j + off[1] * jo, lc = l + off[2] * second;
                            if (!(1 <= ic && ic <= nx && 1 <= jc && jc <= ny && 1 <= lc && lc <= nz)) continue;
                            tw += 0.25;
                            for (int k = 0; k < N; k++) w[k] += 0.25 * PH(k, ic, jc, lc);
                        }
                        finish(w, N, tw, mid[m], st, IX3(e[0], e[1], i, j, l));
                    }
        }
    }
#undef PH
}
