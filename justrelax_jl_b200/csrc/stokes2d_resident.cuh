// stokes2d_resident.cuh — internal interface of the shared-memory-resident 2D-V2 iteration batches (stokes2d_resident.cu)
#pragma once
#include "common.cuh"

struct V2ResArgs {
    int nx, ny;
    double _dx, _dy, dt, r, th, edt;
    int fs_l, fs_r, fs_t, fs_b, ns_l, ns_r, ns_t, ns_b;
    double *Vx[2], *Vy[2], *P[2], *txx[2], *tyy[2], *txy[2];   // the two dense ping-pong sets of the 2D plan
    const double *eta, *etatau, *rhogx, *rhogy, *G, *K, *Q;
};
struct V2ResPlan {
    bool ok = false;   // the state fits on chip and the elastic / compressible terms vanish identically
    int gx = 0, gy = 0, cx = 0, cy = 0, rows = 16;
    size_t smem = 0;
    unsigned long long *flags = nullptr, flag_base = 0;
};
int jr_v2_resident_plan(jr_context *ctx, const V2ResArgs *r, V2ResPlan *p);
int jr_v2_resident_run(jr_context *ctx, const V2ResArgs *r, V2ResPlan *p, int64_t it0, int niter);
