/*
 * mini.h — CPU ORACLE (test infrastructure, NOT product code).
 * Restatement of src/MiniKernels.jl (finite differences, averages, harmonic
 * averages, gathers, mysum) with 1-based indices; pinned bit-for-bit by the
 * values in test/test_mini_kernels.jl:4-117 (tests/test_oracle_mini.py).
 */
#ifndef JR_ORACLE_MINI_H
#define JR_ORACLE_MINI_H
#include "jr_oracle.h"

typedef struct { const double *p; int n1, n2, n3; } arr;   /* n3 = 1 for 2-D */
#define AT3(A, i, j, k) ((A).p[IX3((A).n1, (A).n2, i, j, k)])
#define AT2(A, i, j) ((A).p[IX2((A).n1, i, j)])

/* mysum  MiniKernels.jl:208-233: accumulation order k -> j -> i (i fastest) starting from 0.0 */
static inline double mysum3(int do_inv, arr A, int ia, int ib, int ja, int jb, int ka, int kb)
{
    double s = 0.0;
    for (int k = ka; k <= kb; k++)
        for (int j = ja; j <= jb; j++)
            for (int i = ia; i <= ib; i++) s += do_inv ? orc_inv(AT3(A, i, j, k)) : AT3(A, i, j, k);
    return s;
}
static inline double mysum2(int do_inv, arr A, int ia, int ib, int ja, int jb) { return mysum3(do_inv, A, ia, ib, ja, jb, 1, 1); }

/* ---- 2D, MiniKernels.jl:36-98 ---- */
static inline double d_xa2(arr A, double _dx, int i, int j) { return (-AT2(A, i, j) + AT2(A, i + 1, j)) * _dx; }
static inline double d_ya2(arr A, double _dy, int i, int j) { return (-AT2(A, i, j) + AT2(A, i, j + 1)) * _dy; }
static inline double d_xi2(arr A, double _dx, int i, int j) { return (-AT2(A, i, j + 1) + AT2(A, i + 1, j + 1)) * _dx; }
static inline double d_yi2(arr A, double _dy, int i, int j) { return (-AT2(A, i + 1, j) + AT2(A, i + 1, j + 1)) * _dy; }
static inline double av2(arr A, int i, int j) { return 0.25 * mysum2(0, A, i + 1, i + 2, j + 1, j + 2); }
static inline double av_a2(arr A, int i, int j) { return 0.25 * mysum2(0, A, i, i + 1, j, j + 1); }
static inline double av_xa2(arr A, int i, int j) { return (AT2(A, i, j) + AT2(A, i + 1, j)) * 0.5; }
static inline double av_ya2(arr A, int i, int j) { return (AT2(A, i, j) + AT2(A, i, j + 1)) * 0.5; }
static inline double av_xi2(arr A, int i, int j) { return (AT2(A, i, j + 1) + AT2(A, i + 1, j + 1)) * 0.5; }
static inline double av_yi2(arr A, int i, int j) { return (AT2(A, i + 1, j) + AT2(A, i + 1, j + 1)) * 0.5; }
static inline double av_ai_clamped2(arr A, int i, int j)
{
    int i0 = orc_clamp(i - 1, 1, A.n1), i1 = orc_clamp(i, 1, A.n1);
    int j0 = orc_clamp(j - 1, 1, A.n2), j1 = orc_clamp(j, 1, A.n2);
    return 0.25 * (AT2(A, i0, j0) + AT2(A, i1, j0) + AT2(A, i0, j1) + AT2(A, i1, j1));
}
static inline double harm2(arr A, int i, int j) { return 4.0 * orc_inv(mysum2(1, A, i + 1, i + 2, j + 1, j + 2)); }
static inline double harm_a2(arr A, int i, int j) { return 4.0 * orc_inv(mysum2(1, A, i, i + 1, j, j + 1)); }
static inline double harm_xa2(arr A, int i, int j) { return 2.0 * orc_inv(orc_inv(AT2(A, i + 1, j)) + orc_inv(AT2(A, i, j))); }
static inline double harm_ya2(arr A, int i, int j) { return 2.0 * orc_inv(orc_inv(AT2(A, i, j + 1)) + orc_inv(AT2(A, i, j))); }

/* ---- 3D, MiniKernels.jl:43-45,53-55,100-204 ---- */
static inline double d_xa3(arr A, double _d, int i, int j, int k) { return (-AT3(A, i, j, k) + AT3(A, i + 1, j, k)) * _d; }
static inline double d_ya3(arr A, double _d, int i, int j, int k) { return (-AT3(A, i, j, k) + AT3(A, i, j + 1, k)) * _d; }
static inline double d_za3(arr A, double _d, int i, int j, int k) { return (-AT3(A, i, j, k) + AT3(A, i, j, k + 1)) * _d; }
static inline double d_xi3(arr A, double _d, int i, int j, int k) { return (-AT3(A, i, j + 1, k + 1) + AT3(A, i + 1, j + 1, k + 1)) * _d; }
static inline double d_yi3(arr A, double _d, int i, int j, int k) { return (-AT3(A, i + 1, j, k + 1) + AT3(A, i + 1, j + 1, k + 1)) * _d; }
static inline double d_zi3(arr A, double _d, int i, int j, int k) { return (-AT3(A, i + 1, j + 1, k) + AT3(A, i + 1, j + 1, k + 1)) * _d; }
static inline double av3(arr A, int i, int j, int k) { return 0.125 * mysum3(0, A, i, i + 1, j, j + 1, k, k + 1); }
static inline double av_x3(arr A, int i, int j, int k) { return 0.5 * (AT3(A, i, j, k) + AT3(A, i + 1, j, k)); }
static inline double av_y3(arr A, int i, int j, int k) { return 0.5 * (AT3(A, i, j, k) + AT3(A, i, j + 1, k)); }
static inline double av_z3(arr A, int i, int j, int k) { return 0.5 * (AT3(A, i, j, k) + AT3(A, i, j, k + 1)); }
static inline double av_xy3(arr A, int i, int j, int k) { return 0.25 * mysum3(0, A, i, i + 1, j, j + 1, k, k); }
static inline double av_xz3(arr A, int i, int j, int k) { return 0.25 * mysum3(0, A, i, i + 1, j, j, k, k + 1); }
static inline double av_yz3(arr A, int i, int j, int k) { return 0.25 * mysum3(0, A, i, i, j, j + 1, k, k + 1); }
static inline double av_xyi3(arr A, int i, int j, int k) { return 0.25 * mysum3(0, A, i - 1, i, j - 1, j, k, k); }
static inline double av_xzi3(arr A, int i, int j, int k) { return 0.25 * mysum3(0, A, i - 1, i, j, j, k - 1, k); }
static inline double av_yzi3(arr A, int i, int j, int k) { return 0.25 * mysum3(0, A, i, i, j - 1, j, k - 1, k); }
static inline double av_xyi_clamped3(arr A, int i, int j, int k)
{
    int i0 = orc_clamp(i - 1, 1, A.n1), i1 = orc_clamp(i, 1, A.n1), j0 = orc_clamp(j - 1, 1, A.n2), j1 = orc_clamp(j, 1, A.n2);
    return 0.25 * (AT3(A, i0, j0, k) + AT3(A, i1, j0, k) + AT3(A, i0, j1, k) + AT3(A, i1, j1, k));
}
static inline double av_xzi_clamped3(arr A, int i, int j, int k)
{
    int i0 = orc_clamp(i - 1, 1, A.n1), i1 = orc_clamp(i, 1, A.n1), k0 = orc_clamp(k - 1, 1, A.n3), k1 = orc_clamp(k, 1, A.n3);
    return 0.25 * (AT3(A, i0, j, k0) + AT3(A, i1, j, k0) + AT3(A, i0, j, k1) + AT3(A, i1, j, k1));
}
static inline double av_yzi_clamped3(arr A, int i, int j, int k)
{
    int j0 = orc_clamp(j - 1, 1, A.n2), j1 = orc_clamp(j, 1, A.n2), k0 = orc_clamp(k - 1, 1, A.n3), k1 = orc_clamp(k, 1, A.n3);
    return 0.25 * (AT3(A, i, j0, k0) + AT3(A, i, j1, k0) + AT3(A, i, j0, k1) + AT3(A, i, j1, k1));
}
static inline double harm_x3(arr A, int i, int j, int k) { return 2.0 * orc_inv(orc_inv(AT3(A, i, j, k)) + orc_inv(AT3(A, i + 1, j, k))); }
static inline double harm_y3(arr A, int i, int j, int k) { return 2.0 * orc_inv(orc_inv(AT3(A, i, j, k)) + orc_inv(AT3(A, i, j + 1, k))); }
static inline double harm_z3(arr A, int i, int j, int k) { return 2.0 * orc_inv(orc_inv(AT3(A, i, j, k)) + orc_inv(AT3(A, i, j, k + 1))); }
static inline double harm_xy3(arr A, int i, int j, int k) { return 4.0 * orc_inv(mysum3(1, A, i, i + 1, j, j + 1, k, k)); }
static inline double harm_xz3(arr A, int i, int j, int k) { return 4.0 * orc_inv(mysum3(1, A, i, i + 1, j, j, k, k + 1)); }
static inline double harm_yz3(arr A, int i, int j, int k) { return 4.0 * orc_inv(mysum3(1, A, i, i, j, j + 1, k, k + 1)); }
static inline double harm_xyi3(arr A, int i, int j, int k) { return 4.0 * orc_inv(mysum3(1, A, i - 1, i, j - 1, j, k, k)); }
static inline double harm_xzi3(arr A, int i, int j, int k) { return 4.0 * orc_inv(mysum3(1, A, i - 1, i, j, j, k - 1, k)); }
static inline double harm_yzi3(arr A, int i, int j, int k) { return 4.0 * orc_inv(mysum3(1, A, i, i, j - 1, j, k - 1, k)); }
static inline double harm_xyi_clamped3(arr A, int i, int j, int k)
{
    int i0 = orc_clamp(i - 1, 1, A.n1), i1 = orc_clamp(i, 1, A.n1), j0 = orc_clamp(j - 1, 1, A.n2), j1 = orc_clamp(j, 1, A.n2);
    return 4.0 * orc_inv(orc_inv(AT3(A, i0, j0, k)) + orc_inv(AT3(A, i1, j0, k)) + orc_inv(AT3(A, i0, j1, k)) + orc_inv(AT3(A, i1, j1, k)));
}
static inline double harm_xzi_clamped3(arr A, int i, int j, int k)
{
    int i0 = orc_clamp(i - 1, 1, A.n1), i1 = orc_clamp(i, 1, A.n1), k0 = orc_clamp(k - 1, 1, A.n3), k1 = orc_clamp(k, 1, A.n3);
    return 4.0 * orc_inv(orc_inv(AT3(A, i0, j, k0)) + orc_inv(AT3(A, i1, j, k0)) + orc_inv(AT3(A, i0, j, k1)) + orc_inv(AT3(A, i1, j, k1)));
}
static inline double harm_yzi_clamped3(arr A, int i, int j, int k)
{
    int j0 = orc_clamp(j - 1, 1, A.n2), j1 = orc_clamp(j, 1, A.n2), k0 = orc_clamp(k - 1, 1, A.n3), k1 = orc_clamp(k, 1, A.n3);
    return 4.0 * orc_inv(orc_inv(AT3(A, i, j0, k0)) + orc_inv(AT3(A, i, j1, k0)) + orc_inv(AT3(A, i, j0, k1)) + orc_inv(AT3(A, i, j1, k1)));
}
#endif
