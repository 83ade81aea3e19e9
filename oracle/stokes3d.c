/*
 * stokes3d.c — CPU ORACLE (test infrastructure, NOT product code).
 * Restatement of the 3D visco-elastic Stokes PT loop, variant 3D-VA:
 *   src/stokes/Stokes3D.jl:25-186 (driver)
 *   src/stokes/VelocityKernels.jl:3-6,59-104,182-242
 *   src/stokes/PressureKernels.jl:10-15,186-195
 *   src/stokes/StressKernels.jl:2-5,149-230 ; src/rheology/StressUpdate.jl:70
 *   src/MiniKernels.jl ; src/Utils.jl:62-72,409-461,698-701
 *   src/boundaryconditions/{free_slip,no_slip,periodic}.jl ; src/types/displacement.jl
 * All indices 1-based through the IX3 macro, loops k→j→i.
 */
#include "jr_oracle.h"
#include <stdlib.h>
#include <string.h>

#define F(name) (s->f[ORC_F_##name])

static const char *k_names[] = {
#define X(n) #n,
    JR_STOKES_FIELDS(X)
#undef X
};
const char *orc_field_name(int i) { return (i >= 0 && i < ORC_F_COUNT) ? k_names[i] : NULL; }
int orc_field_count(void) { return ORC_F_COUNT; }

/* compute_∇V!  VelocityKernels.jl:3-6 ; div MiniKernels.jl:104-105 ; _d_xi/_d_yi/_d_zi :53-55 */
void orc_compute_divV3(const orc_fields *s, const double _di[3])
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const double _dx = _di[0], _dy = _di[1], _dz = _di[2];
    const double *Vx = F(Vx), *Vy = F(Vy), *Vz = F(Vz);
    double *divV = F(divV);
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                double dx = (-Vx[IX3(nx + 1, ny + 2, i, j + 1, k + 1)] + Vx[IX3(nx + 1, ny + 2, i + 1, j + 1, k + 1)]) * _dx;
                double dy = (-Vy[IX3(nx + 2, ny + 1, i + 1, j, k + 1)] + Vy[IX3(nx + 2, ny + 1, i + 1, j + 1, k + 1)]) * _dy;
                double dz = (-Vz[IX3(nx + 2, ny + 2, i + 1, j + 1, k)] + Vz[IX3(nx + 2, ny + 2, i + 1, j + 1, k + 1)]) * _dz;
                divV[IX3(nx, ny, i, j, k)] = dx + dy + dz;
            }
}

/* _compute_P!  PressureKernels.jl:186-195 (muladd -> fma) */
static inline void compute_P_point(double *RP, double *P, double P0, double divV, double Q, double eta,
                                   double K, double G, double dt, double r, double theta_dtau)
{
    double _Kdt = orc_inv(K * dt);
    double _Gdt = orc_inv(G * dt);
    double _dt = orc_inv(dt);
    double Pc = *P;
    *RP = fma(-(Pc - P0), _Kdt, (-divV + (Q * _dt)));
    double psi = orc_inv(orc_inv(eta) + _Gdt) * r / theta_dtau;
    *P = ((fma(P0, _Kdt, (-divV + (Q * _dt)))) * psi + Pc) / (1 + _Kdt * psi);
}

/* compute_P! compressible form, PressureKernels.jl:10-15.  `eta` is η (3D-VA, Stokes3D.jl:85)
 * or ητ (2D-V2, Stokes2D.jl:231-233) — Q5. */
void orc_compute_P_VA(const orc_fields *s, const double *eta, double dt, double r, double theta_dtau)
{
    const size_t N = (size_t)s->n[0] * s->n[1] * (s->ndim == 3 ? s->n[2] : 1);
    double *P = F(P), *RP = F(RP);
    const double *P0 = F(P0), *divV = F(divV), *Q = F(Q), *K = F(K), *G = F(G);
#pragma omp parallel for schedule(static)
    for (size_t I = 0; I < N; I++)
        compute_P_point(&RP[I], &P[I], P0[I], divV[I], Q[I], eta[I], K[I], G[I], dt, r, theta_dtau);
}

/* compute_strain_rate! 3D  VelocityKernels.jl:59-104.  Launch range ni.+1 (VA, Stokes3D.jl:92)
 * or ni (VC, Stokes3D.jl:533, quirk Q20). */
void orc_compute_strain_rate3(const orc_fields *s, const double _di[3], int range_plus1)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const double _dx = _di[0], _dy = _di[1], _dz = _di[2];
    const double *Vx = F(Vx), *Vy = F(Vy), *Vz = F(Vz), *divV = F(divV);
    double *exx = F(exx), *eyy = F(eyy), *ezz = F(ezz), *eyz = F(eyz), *exz = F(exz), *exy = F(exy);
    const int p = range_plus1 ? 1 : 0;
#define VX(i, j, k) Vx[IX3(nx + 1, ny + 2, i, j, k)]
#define VY(i, j, k) Vy[IX3(nx + 2, ny + 1, i, j, k)]
#define VZ(i, j, k) Vz[IX3(nx + 2, ny + 2, i, j, k)]
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz + p; k++)
        for (int j = 1; j <= ny + p; j++)
            for (int i = 1; i <= nx + p; i++) {
                if (i <= nx && j <= ny && k <= nz) {
                    double d3 = divV[IX3(nx, ny, i, j, k)] * orc_inv(3.0);
                    exx[IX3(nx, ny, i, j, k)] = (-VX(i, j + 1, k + 1) + VX(i + 1, j + 1, k + 1)) * _dx - d3;
                    eyy[IX3(nx, ny, i, j, k)] = (-VY(i + 1, j, k + 1) + VY(i + 1, j + 1, k + 1)) * _dy - d3;
                    ezz[IX3(nx, ny, i, j, k)] = (-VZ(i + 1, j + 1, k) + VZ(i + 1, j + 1, k + 1)) * _dz - d3;
                }
                if (i <= nx && j <= ny + 1 && k <= nz + 1)
                    eyz[IX3(nx, ny + 1, i, j, k)] =
                        0.5 * (_dz * (VY(i + 1, j, k + 1) - VY(i + 1, j, k)) + _dy * (VZ(i + 1, j + 1, k) - VZ(i + 1, j, k)));
                if (i <= nx + 1 && j <= ny && k <= nz + 1)
                    exz[IX3(nx + 1, ny, i, j, k)] =
                        0.5 * (_dz * (VX(i, j + 1, k + 1) - VX(i, j + 1, k)) + _dx * (VZ(i + 1, j + 1, k) - VZ(i, j + 1, k)));
                if (i <= nx + 1 && j <= ny + 1 && k <= nz)
                    exy[IX3(nx + 1, ny + 1, i, j, k)] =
                        0.5 * (_dy * (VX(i, j + 1, k + 1) - VX(i, j, k + 1)) + _dx * (VY(i + 1, j, k + 1) - VY(i, j, k + 1)));
            }
#undef VX
#undef VY
#undef VZ
}

/* compute_dτ_r  StressUpdate.jl:70 ; compute_stress_increment StressKernels.jl:2-5 */
static inline double dtau_r(double theta_dtau, double eta, double _Gdt) { return orc_inv(theta_dtau + fma(eta, _Gdt, 1.0)); }
static inline double stress_increment(double t, double t_o, double eta, double e, double _Gdt, double dtr)
{
    return dtr * fma(2.0 * eta, e, fma(-(t - t_o) * eta, _Gdt, -t));
}

/* compute_τ! 3D visco-elastic  StressKernels.jl:149-230; clamped 4-cell averages MiniKernels.jl:133-147 */
void orc_compute_tau3_VE(const orc_fields *s, double dt, double theta_dtau)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const double *eta = F(eta), *G = F(G);
    double *txx = F(txx), *tyy = F(tyy), *tzz = F(tzz), *tyz = F(tyz), *txz = F(txz), *txy = F(txy);
    const double *txx_o = F(txx_o), *tyy_o = F(tyy_o), *tzz_o = F(tzz_o), *tyz_o = F(tyz_o), *txz_o = F(txz_o), *txy_o = F(txy_o);
    const double *exx = F(exx), *eyy = F(eyy), *ezz = F(ezz), *eyz = F(eyz), *exz = F(exz), *exy = F(exy);
#define C(A, i, j, k) A[IX3(nx, ny, i, j, k)]
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz + 1; k++)
        for (int j = 1; j <= ny + 1; j++)
            for (int i = 1; i <= nx + 1; i++) {
                const int i0 = orc_clamp(i - 1, 1, nx), i1 = orc_clamp(i, 1, nx);
                const int j0 = orc_clamp(j - 1, 1, ny), j1 = orc_clamp(j, 1, ny);
                const int k0 = orc_clamp(k - 1, 1, nz), k1 = orc_clamp(k, 1, nz);
                if (i <= nx && j <= ny && k <= nz) {
                    size_t I = IX3(nx, ny, i, j, k);
                    double _Gdt = orc_inv(G[I] * dt);
                    double e = eta[I];
                    double dtr = dtau_r(theta_dtau, e, _Gdt);
                    txx[I] += stress_increment(txx[I], txx_o[I], e, exx[I], _Gdt, dtr);
                    tyy[I] += stress_increment(tyy[I], tyy_o[I], e, eyy[I], _Gdt, dtr);
                    tzz[I] += stress_increment(tzz[I], tzz_o[I], e, ezz[I], _Gdt, dtr);
                }
                if (i <= nx + 1 && j <= ny + 1 && k <= nz) {
                    size_t I = IX3(nx + 1, ny + 1, i, j, k);
                    double e = 0.25 * (C(eta, i0, j0, k) + C(eta, i1, j0, k) + C(eta, i0, j1, k) + C(eta, i1, j1, k));
                    double g = 0.25 * (C(G, i0, j0, k) + C(G, i1, j0, k) + C(G, i0, j1, k) + C(G, i1, j1, k));
                    double _Gdt = orc_inv(g * dt);
                    double dtr = dtau_r(theta_dtau, e, _Gdt);
                    txy[I] += stress_increment(txy[I], txy_o[I], e, exy[I], _Gdt, dtr);
                }
                if (i <= nx + 1 && j <= ny && k <= nz + 1) {
                    size_t I = IX3(nx + 1, ny, i, j, k);
                    double e = 0.25 * (C(eta, i0, j, k0) + C(eta, i1, j, k0) + C(eta, i0, j, k1) + C(eta, i1, j, k1));
                    double g = 0.25 * (C(G, i0, j, k0) + C(G, i1, j, k0) + C(G, i0, j, k1) + C(G, i1, j, k1));
                    double _Gdt = orc_inv(g * dt);
                    double dtr = dtau_r(theta_dtau, e, _Gdt);
                    txz[I] += stress_increment(txz[I], txz_o[I], e, exz[I], _Gdt, dtr);
                }
                if (i <= nx && j <= ny + 1 && k <= nz + 1) {
                    size_t I = IX3(nx, ny + 1, i, j, k);
                    double e = 0.25 * (C(eta, i, j0, k0) + C(eta, i, j1, k0) + C(eta, i, j0, k1) + C(eta, i, j1, k1));
                    double g = 0.25 * (C(G, i, j0, k0) + C(G, i, j1, k0) + C(G, i, j0, k1) + C(G, i, j1, k1));
                    double _Gdt = orc_inv(g * dt);
                    double dtr = dtau_r(theta_dtau, e, _Gdt);
                    tyz[I] += stress_increment(tyz[I], tyz_o[I], e, eyz[I], _Gdt, dtr);
                }
            }
#undef C
}

/* compute_V! 3D  VelocityKernels.jl:182-242 */
void orc_compute_V3(const orc_fields *s, double eta_dtau, const double _di[3])
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const double _dx = _di[0], _dy = _di[1], _dz = _di[2];
    double *Vx = F(Vx), *Vy = F(Vy), *Vz = F(Vz), *Rx = F(Rx), *Ry = F(Ry), *Rz = F(Rz);
    const double *P = F(P), *fx = F(rhogx), *fy = F(rhogy), *fz = F(rhogz), *etatau = F(etatau);
    const double *txx = F(txx), *tyy = F(tyy), *tzz = F(tzz), *tyz = F(tyz), *txz = F(txz), *txy = F(txy);
#define C(A, i, j, k) A[IX3(nx, ny, i, j, k)]
#define XY(i, j, k) txy[IX3(nx + 1, ny + 1, i, j, k)]
#define XZ(i, j, k) txz[IX3(nx + 1, ny, i, j, k)]
#define YZ(i, j, k) tyz[IX3(nx, ny + 1, i, j, k)]
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                if (i <= nx - 1) {
                    double R = (-C(txx, i, j, k) + C(txx, i + 1, j, k)) * _dx + _dy * (XY(i + 1, j + 1, k) - XY(i + 1, j, k)) +
                               _dz * (XZ(i + 1, j, k + 1) - XZ(i + 1, j, k)) - (-C(P, i, j, k) + C(P, i + 1, j, k)) * _dx -
                               0.5 * (C(fx, i, j, k) + C(fx, i + 1, j, k));
                    Rx[IX3(nx - 1, ny, i, j, k)] = R;
                    Vx[IX3(nx + 1, ny + 2, i + 1, j + 1, k + 1)] += R * eta_dtau / (0.5 * (C(etatau, i, j, k) + C(etatau, i + 1, j, k)));
                }
                if (j <= ny - 1) {
                    double R = _dx * (XY(i + 1, j + 1, k) - XY(i, j + 1, k)) + _dy * (C(tyy, i, j + 1, k) - C(tyy, i, j, k)) +
                               _dz * (YZ(i, j + 1, k + 1) - YZ(i, j + 1, k)) - (-C(P, i, j, k) + C(P, i, j + 1, k)) * _dy -
                               0.5 * (C(fy, i, j, k) + C(fy, i, j + 1, k));
                    Ry[IX3(nx, ny - 1, i, j, k)] = R;
                    Vy[IX3(nx + 2, ny + 1, i + 1, j + 1, k + 1)] += R * eta_dtau / (0.5 * (C(etatau, i, j, k) + C(etatau, i, j + 1, k)));
                }
                if (k <= nz - 1) {
                    double R = _dx * (XZ(i + 1, j, k + 1) - XZ(i, j, k + 1)) + _dy * (YZ(i, j + 1, k + 1) - YZ(i, j, k + 1)) +
                               (-C(tzz, i, j, k) + C(tzz, i, j, k + 1)) * _dz - (-C(P, i, j, k) + C(P, i, j, k + 1)) * _dz -
                               0.5 * (C(fz, i, j, k) + C(fz, i, j, k + 1));
                    Rz[IX3(nx, ny, i, j, k)] = R;
                    Vz[IX3(nx + 2, ny + 2, i + 1, j + 1, k + 1)] += R * eta_dtau / (0.5 * (C(etatau, i, j, k) + C(etatau, i, j, k + 1)));
                }
            }
#undef C
#undef XY
#undef XZ
#undef YZ
}

/* velocity2displacement!  types/displacement.jl:7-29 */
void orc_velocity2displacement(const orc_fields *s, double dt)
{
    const int nx = s->n[0], ny = s->n[1], nz = (s->ndim == 3 ? s->n[2] : 0);
    size_t nVx, nVy, nVz = 0;
    if (s->ndim == 3) {
        nVx = (size_t)(nx + 1) * (ny + 2) * (nz + 2);
        nVy = (size_t)(nx + 2) * (ny + 1) * (nz + 2);
        nVz = (size_t)(nx + 2) * (ny + 2) * (nz + 1);
    } else {
        nVx = (size_t)(nx + 1) * (ny + 2);
        nVy = (size_t)(nx + 2) * (ny + 1);
    }
    double *Ux = F(Ux), *Uy = F(Uy), *Uz = F(Uz);
    const double *Vx = F(Vx), *Vy = F(Vy), *Vz = F(Vz);
#pragma omp parallel for schedule(static)
    for (size_t I = 0; I < nVx; I++) Ux[I] = Vx[I] * dt;
#pragma omp parallel for schedule(static)
    for (size_t I = 0; I < nVy; I++) Uy[I] = Vy[I] * dt;
#pragma omp parallel for schedule(static)
    for (size_t I = 0; I < nVz; I++) Uz[I] = Vz[I] * dt;
}

/* free_slip! 3D velocities  free_slip.jl:15-70.
 * bc order: left,right,front,back,top,bot.  NOTE quirk Q2: bc.top fills k=1, bc.bot fills k=end.
 * Schedule: the reference runs all six in one launch (order-dependent ghost edges, Q3);
 * the oracle applies front/back, then top/bot, then left/right as three complete sweeps,
 * which is the fixed point the reference's own tests reach by applying the kernel twice. */
void orc_free_slip3(double *Ax, double *Ay, double *Az, const int32_t n[3], const int32_t bc[6])
{
    const int nx = n[0], ny = n[1], nz = n[2];
    const int ax1 = nx + 1, ax2 = ny + 2, ax3 = nz + 2;
    const int ay1 = nx + 2, ay2 = ny + 1, ay3 = nz + 2;
    const int az1 = nx + 2, az2 = ny + 2, az3 = nz + 1;
#define AX(i, j, k) Ax[IX3(ax1, ax2, i, j, k)]
#define AY(i, j, k) Ay[IX3(ay1, ay2, i, j, k)]
#define AZ(i, j, k) Az[IX3(az1, az2, i, j, k)]
    if (bc[2]) { /* front: y = 1 */
        for (int k = 1; k <= ax3; k++) for (int i = 1; i <= ax1; i++) AX(i, 1, k) = AX(i, 2, k);
        for (int k = 1; k <= az3; k++) for (int i = 1; i <= az1; i++) AZ(i, 1, k) = AZ(i, 2, k);
    }
    if (bc[3]) { /* back: y = end */
        for (int k = 1; k <= ax3; k++) for (int i = 1; i <= ax1; i++) AX(i, ax2, k) = AX(i, ax2 - 1, k);
        for (int k = 1; k <= az3; k++) for (int i = 1; i <= az1; i++) AZ(i, az2, k) = AZ(i, az2 - 1, k);
    }
    if (bc[4]) { /* "top": k = 1 (Q2) */
        for (int j = 1; j <= ax2; j++) for (int i = 1; i <= ax1; i++) AX(i, j, 1) = AX(i, j, 2);
        for (int j = 1; j <= ay2; j++) for (int i = 1; i <= ay1; i++) AY(i, j, 1) = AY(i, j, 2);
    }
    if (bc[5]) { /* "bot": k = end (Q2) */
        for (int j = 1; j <= ax2; j++) for (int i = 1; i <= ax1; i++) AX(i, j, ax3) = AX(i, j, ax3 - 1);
        for (int j = 1; j <= ay2; j++) for (int i = 1; i <= ay1; i++) AY(i, j, ay3) = AY(i, j, ay3 - 1);
    }
    if (bc[0]) { /* left: x = 1 */
        for (int k = 1; k <= ay3; k++) for (int j = 1; j <= ay2; j++) AY(1, j, k) = AY(2, j, k);
        for (int k = 1; k <= az3; k++) for (int j = 1; j <= az2; j++) AZ(1, j, k) = AZ(2, j, k);
    }
    if (bc[1]) { /* right: x = end */
        for (int k = 1; k <= ay3; k++) for (int j = 1; j <= ay2; j++) AY(ay1, j, k) = AY(ay1 - 1, j, k);
        for (int k = 1; k <= az3; k++) for (int j = 1; j <= az2; j++) AZ(az1, j, k) = AZ(az1 - 1, j, k);
    }
}

/* no_slip! 3D  no_slip.jl:21-54 (sequential broadcasts: left,right,front,back,bot,top;
 * here bot is k=1 and top is k=end). */
void orc_no_slip3(double *Ax, double *Ay, double *Az, const int32_t n[3], const int32_t bc[6])
{
    const int nx = n[0], ny = n[1], nz = n[2];
    const int ax1 = nx + 1, ax2 = ny + 2, ax3 = nz + 2;
    const int ay1 = nx + 2, ay2 = ny + 1, ay3 = nz + 2;
    const int az1 = nx + 2, az2 = ny + 2, az3 = nz + 1;
    if (bc[0]) {
        for (int k = 1; k <= ax3; k++) for (int j = 1; j <= ax2; j++) AX(1, j, k) = 0;
        for (int k = 1; k <= ay3; k++) for (int j = 1; j <= ay2; j++) AY(1, j, k) = -AY(2, j, k);
        for (int k = 1; k <= az3; k++) for (int j = 1; j <= az2; j++) AZ(1, j, k) = -AZ(2, j, k);
    }
    if (bc[1]) {
        for (int k = 1; k <= ax3; k++) for (int j = 1; j <= ax2; j++) AX(ax1, j, k) = 0;
        for (int k = 1; k <= ay3; k++) for (int j = 1; j <= ay2; j++) AY(ay1, j, k) = -AY(ay1 - 1, j, k);
        for (int k = 1; k <= az3; k++) for (int j = 1; j <= az2; j++) AZ(az1, j, k) = -AZ(az1 - 1, j, k);
    }
    if (bc[2]) {
        for (int k = 1; k <= ax3; k++) for (int i = 1; i <= ax1; i++) AX(i, 1, k) = -AX(i, 2, k);
        for (int k = 1; k <= ay3; k++) for (int i = 1; i <= ay1; i++) AY(i, 1, k) = 0;
        for (int k = 1; k <= az3; k++) for (int i = 1; i <= az1; i++) AZ(i, 1, k) = -AZ(i, 2, k);
    }
    if (bc[3]) {
        for (int k = 1; k <= ax3; k++) for (int i = 1; i <= ax1; i++) AX(i, ax2, k) = -AX(i, ax2 - 1, k);
        for (int k = 1; k <= ay3; k++) for (int i = 1; i <= ay1; i++) AY(i, ay2, k) = 0;
        for (int k = 1; k <= az3; k++) for (int i = 1; i <= az1; i++) AZ(i, az2, k) = -AZ(i, az2 - 1, k);
    }
    if (bc[5]) { /* bot: k = 1 */
        for (int j = 1; j <= ax2; j++) for (int i = 1; i <= ax1; i++) AX(i, j, 1) = -AX(i, j, 2);
        for (int j = 1; j <= ay2; j++) for (int i = 1; i <= ay1; i++) AY(i, j, 1) = -AY(i, j, 2);
        for (int j = 1; j <= az2; j++) for (int i = 1; i <= az1; i++) AZ(i, j, 1) = 0;
    }
    if (bc[4]) { /* top: k = end */
        for (int j = 1; j <= ax2; j++) for (int i = 1; i <= ax1; i++) AX(i, j, ax3) = -AX(i, j, ax3 - 1);
        for (int j = 1; j <= ay2; j++) for (int i = 1; i <= ay1; i++) AY(i, j, ay3) = -AY(i, j, ay3 - 1);
        for (int j = 1; j <= az2; j++) for (int i = 1; i <= az1; i++) AZ(i, j, az3) = 0;
    }
}

/* periodic_boundary! 3D velocities  periodic.jl:56-98 (left/right, front/back, bot/top sweeps) */
void orc_periodic3(double *Ax, double *Ay, double *Az, const int32_t n[3], const int32_t bc[6])
{
    const int nx = n[0], ny = n[1], nz = n[2];
    const int ax1 = nx + 1, ax2 = ny + 2, ax3 = nz + 2;
    const int ay1 = nx + 2, ay2 = ny + 1, ay3 = nz + 2;
    const int az1 = nx + 2, az2 = ny + 2, az3 = nz + 1;
    if (bc[0]) for (int k = 1; k <= ax3; k++) for (int j = 1; j <= ax2; j++) AX(1, j, k) = AX(ax1, j, k);
    for (int k = 1; k <= ay3; k++) for (int j = 1; j <= ay2; j++) {
        if (bc[0]) AY(1, j, k) = AY(ay1 - 1, j, k);
        if (bc[1]) AY(ay1, j, k) = AY(2, j, k);
    }
    for (int k = 1; k <= az3; k++) for (int j = 1; j <= az2; j++) {
        if (bc[0]) AZ(1, j, k) = AZ(az1 - 1, j, k);
        if (bc[1]) AZ(az1, j, k) = AZ(2, j, k);
    }
    for (int k = 1; k <= ax3; k++) for (int i = 1; i <= ax1; i++) {
        if (bc[2]) AX(i, 1, k) = AX(i, ax2 - 1, k);
        if (bc[3]) AX(i, ax2, k) = AX(i, 2, k);
    }
    if (bc[2]) for (int k = 1; k <= ay3; k++) for (int i = 1; i <= ay1; i++) AY(i, 1, k) = AY(i, ay2, k);
    for (int k = 1; k <= az3; k++) for (int i = 1; i <= az1; i++) {
        if (bc[2]) AZ(i, 1, k) = AZ(i, az2 - 1, k);
        if (bc[3]) AZ(i, az2, k) = AZ(i, 2, k);
    }
    for (int j = 1; j <= ax2; j++) for (int i = 1; i <= ax1; i++) {
        if (bc[5]) AX(i, j, 1) = AX(i, j, ax3 - 1);
        if (bc[4]) AX(i, j, ax3) = AX(i, j, 2);
    }
    for (int j = 1; j <= ay2; j++) for (int i = 1; i <= ay1; i++) {
        if (bc[5]) AY(i, j, 1) = AY(i, j, ay3 - 1);
        if (bc[4]) AY(i, j, ay3) = AY(i, j, 2);
    }
    if (bc[5]) for (int j = 1; j <= az2; j++) for (int i = 1; i <= az1; i++) AZ(i, j, 1) = AZ(i, j, az3);
}
#undef AX
#undef AY
#undef AZ

static int any6(const int32_t b[6]) { return b[0] | b[1] | b[2] | b[3] | b[4] | b[5]; }

/* _flow_bcs!  BoundaryConditions.jl:86-99: no_slip -> free_slip -> periodic */
void orc_flow_bcs3(const orc_fields *s, const orc_stokes_opts *o, int displacement)
{
    double *Ax = displacement ? F(Ux) : F(Vx), *Ay = displacement ? F(Uy) : F(Vy), *Az = displacement ? F(Uz) : F(Vz);
    if (any6(o->no_slip)) orc_no_slip3(Ax, Ay, Az, s->n, o->no_slip);
    if (any6(o->free_slip)) orc_free_slip3(Ax, Ay, Az, s->n, o->free_slip);
    if (any6(o->periodic)) orc_periodic3(Ax, Ay, Az, s->n, o->periodic);
}

/* compute_maxloc! / _maxloc_window_clamped  Utils.jl:409-461 */
void orc_maxloc3(double *B, const double *A, int nx, int ny, int nz, int wx, int wy, int wz)
{
#pragma omp parallel for schedule(static)
    for (int K = 1; K <= nz; K++)
        for (int J = 1; J <= ny; J++)
            for (int I = 1; I <= nx; I++) {
                double x = -INFINITY;
                for (int k = K - wz; k <= K + wz; k++) {
                    int kk = orc_clamp(k, 1, nz);
                    for (int j = J - wy; j <= J + wy; j++) {
                        int jj = orc_clamp(j, 1, ny);
                        for (int i = I - wx; i <= I + wx; i++) {
                            int ii = orc_clamp(i, 1, nx);
                            double a = A[IX3(nx, ny, ii, jj, kk)];
                            if (a > x) x = a;
                        }
                    }
                }
                B[IX3(nx, ny, I, J, K)] = x;
            }
}

/* sum(A[2:end-1, 2:end-1, 2:end-1] .^ 2) (interior != 0) or sum(A .^ 2).  n3 = 1 with a 2-D array
 * slices only the first two dims.  Plain left-to-right accumulation per k-plane, planes added in order
 * (Julia uses pairwise summation; the difference is O(1e-16) relative and is covered by the
 * ±1 % iteration-count tolerance, never by field parity). */
double orc_sumsq_interior(const double *A, int n1, int n2, int n3, int interior)
{
    const int o = interior ? 1 : 0;
    const int k_lo = (n3 > 1) ? 1 + o : 1;
    const int k_hi = (n3 > 1) ? n3 - o : 1;
    double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int k = k_lo; k <= k_hi; k++) {
        double sk = 0.0;
        for (int j = 1 + o; j <= n2 - o; j++)
            for (int i = 1 + o; i <= n1 - o; i++) {
                double a = A[IX3(n1, n2, i, j, k)];
                sk += a * a;
            }
        total += sk;
    }
    return total;
}

/* multi_copy! τ→τ_o  Stokes3D.jl:172-173 (vertex set over ni.+1, centre set over ni) */
void orc_multi_copy_tau3(const orc_fields *s)
{
    const size_t nx = s->n[0], ny = s->n[1], nz = s->n[2];
    const size_t nc = nx * ny * nz;
    memcpy(F(txx_o), F(txx), nc * 8); memcpy(F(tyy_o), F(tyy), nc * 8); memcpy(F(tzz_o), F(tzz), nc * 8);
    memcpy(F(tyz_o), F(tyz), nx * (ny + 1) * (nz + 1) * 8);
    memcpy(F(txz_o), F(txz), (nx + 1) * ny * (nz + 1) * 8);
    memcpy(F(txy_o), F(txy), (nx + 1) * (ny + 1) * nz * 8);
    if (F(tyz_c) && F(tyz_o_c)) memcpy(F(tyz_o_c), F(tyz_c), nc * 8);
    if (F(txz_c) && F(txz_o_c)) memcpy(F(txz_o_c), F(txz_c), nc * 8);
    if (F(txy_c) && F(txy_o_c)) memcpy(F(txy_o_c), F(txy_c), nc * 8);
}

/* one PT iteration of 3D-VA, Stokes3D.jl:78-121 */
static void iterate3d_VA_once(const orc_fields *s, const orc_stokes_opts *o)
{
    orc_compute_divV3(s, o->_di);
    orc_compute_P_VA(s, F(eta), o->dt, o->r, o->theta_dtau);
    orc_compute_strain_rate3(s, o->_di, 1);
    orc_compute_tau3_VE(s, o->dt, o->theta_dtau);
    orc_compute_V3(s, o->eta_dtau, o->_di);
    orc_velocity2displacement(s, o->dt);
    orc_flow_bcs3(s, o, 0);
    /* update_halo!(Vx,Vy,Vz): single rank, no neighbours */
}

static void pre_VA(const orc_fields *s)
{
    /* ητ = deepcopy(η); compute_maxloc!(ητ, η)  Stokes3D.jl:55-57 */
    orc_maxloc3(F(etatau), F(eta), s->n[0], s->n[1], s->n[2], 1, 1, 1);
}

/* pieces of the loop for the multi-rank emulation of the tests (tests/mrank.py does update_halo! between them) */
void orc_pre3d_VA(const orc_fields *s) { pre_VA(s); }
void orc_iterate3d_VA_once(const orc_fields *s, const orc_stokes_opts *o) { iterate3d_VA_once(s, o); }

int orc_iterate3d_VA(const orc_fields *s, const orc_stokes_opts *o, int64_t niter)
{
    pre_VA(s);
    for (int64_t it = 0; it < niter; it++) iterate3d_VA_once(s, o);
    return 0;
}

/* _solve! 3D-VA  Stokes3D.jl:25-186 */
int orc_solve3d_VA(const orc_fields *s, const orc_stokes_opts *o, orc_stokes_result *res)
{
    const int nx = s->n[0], ny = s->n[1], nz = s->n[2];
    double err_it1 = 1.0, err = 1.0;
    int64_t iter = 0, cont = 0;
    pre_VA(s);
    while (iter < 2 || (((err / err_it1) > o->eps_rel && err > o->eps_abs) && iter <= o->iterMax)) {
        iterate3d_VA_once(s, o);
        iter += 1;
        if (iter % o->nout == 0 && iter > 1) {
            double nRx = sqrt(orc_sumsq_interior(F(Rx), nx - 1, ny, nz, 1)) /
                         ((double)(o->n_g[0] - 2) * (o->n_g[1] - 1) * (o->n_g[2] - 1));
            double nRy = sqrt(orc_sumsq_interior(F(Ry), nx, ny - 1, nz, 1)) /
                         ((double)(o->n_g[0] - 1) * (o->n_g[1] - 2) * (o->n_g[2] - 1));
            double nRz = sqrt(orc_sumsq_interior(F(Rz), nx, ny, nz - 1, 1)) /
                         ((double)(o->n_g[0] - 1) * (o->n_g[1] - 1) * (o->n_g[2] - 2));
            double nP = sqrt(orc_sumsq_interior(F(RP), nx, ny, nz, 0)) /
                        ((double)(o->n_g[0]) * (o->n_g[1]) * (o->n_g[2]));
            res->norm_Rx[cont] = nRx; res->norm_Ry[cont] = nRy; res->norm_Rz[cont] = nRz; res->norm_divV[cont] = nP;
            err = fmax(fmax(nRx, nRy), fmax(nRz, nP));
            if (isnan(nRx) || isnan(nRy) || isnan(nRz) || isnan(nP)) err = NAN;
            res->err_evo1[cont] = err; res->err_evo2[cont] = iter;
            cont += 1;
            err_it1 = fmax(fmax(res->norm_Rx[0], res->norm_Ry[0]), fmax(res->norm_Rz[0], res->norm_divV[0]));
            if (isnan(err)) { res->iter = iter; res->nhist = cont; res->err = err; return 1; } /* error("NaN(s)") */
        }
    }
    orc_multi_copy_tau3(s);
    res->iter = iter; res->nhist = cont; res->err = err;
    return 0;
}
