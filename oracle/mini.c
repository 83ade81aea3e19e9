/*
 * mini.c — CPU ORACLE (test infrastructure, NOT product code).
 * Name-dispatched entry points onto mini.h so that tests/test_oracle_mini.py can
 * replay test/test_mini_kernels.jl:4-117 value by value.
 */
#include "mini.h"
#include <string.h>

double orc_mini3(const char *name, const double *p, int n1, int n2, int n3, double _d, int i, int j, int k)
{
    arr A = {p, n1, n2, n3};
#define CASE(s, expr) if (!strcmp(name, s)) return (expr)
    CASE("center", AT3(A, i, j, k)); CASE("left", AT3(A, i - 1, j, k)); CASE("right", AT3(A, i + 1, j, k));
    CASE("back", AT3(A, i, j - 1, k)); CASE("front", AT3(A, i, j + 1, k)); CASE("bot", AT3(A, i, j, k - 1));
    CASE("top", AT3(A, i, j, k + 1)); CASE("next", AT3(A, i + 1, j + 1, k + 1));
    CASE("_d_xa", d_xa3(A, _d, i, j, k)); CASE("_d_ya", d_ya3(A, _d, i, j, k)); CASE("_d_za", d_za3(A, _d, i, j, k));
    CASE("_d_xi", d_xi3(A, _d, i, j, k)); CASE("_d_yi", d_yi3(A, _d, i, j, k)); CASE("_d_zi", d_zi3(A, _d, i, j, k));
    CASE("_av", av3(A, i, j, k)); CASE("_av_x", av_x3(A, i, j, k)); CASE("_av_y", av_y3(A, i, j, k)); CASE("_av_z", av_z3(A, i, j, k));
    CASE("_av_xy", av_xy3(A, i, j, k)); CASE("_av_xz", av_xz3(A, i, j, k)); CASE("_av_yz", av_yz3(A, i, j, k));
    CASE("_av_xyi", av_xyi3(A, i, j, k)); CASE("_av_xzi", av_xzi3(A, i, j, k)); CASE("_av_yzi", av_yzi3(A, i, j, k));
    CASE("_av_xyi_clamped", av_xyi_clamped3(A, i, j, k)); CASE("_av_xzi_clamped", av_xzi_clamped3(A, i, j, k));
    CASE("_av_yzi_clamped", av_yzi_clamped3(A, i, j, k));
    CASE("_harm_x", harm_x3(A, i, j, k)); CASE("_harm_y", harm_y3(A, i, j, k)); CASE("_harm_z", harm_z3(A, i, j, k));
    CASE("_harm_xy", harm_xy3(A, i, j, k)); CASE("_harm_xz", harm_xz3(A, i, j, k)); CASE("_harm_yz", harm_yz3(A, i, j, k));
    CASE("_harm_xyi", harm_xyi3(A, i, j, k)); CASE("_harm_xzi", harm_xzi3(A, i, j, k)); CASE("_harm_yzi", harm_yzi3(A, i, j, k));
    CASE("_harm_xyi_clamped", harm_xyi_clamped3(A, i, j, k)); CASE("_harm_xzi_clamped", harm_xzi_clamped3(A, i, j, k));
    CASE("_harm_yzi_clamped", harm_yzi_clamped3(A, i, j, k));
    /* mysum over (i:i+1, j:j+1, k:k+1); _d != 0 selects mysum(inv, ...) */
    CASE("mysum", mysum3(_d != 0.0, A, i, i + 1, j, j + 1, k, k + 1));
    return NAN;
}

double orc_mini2(const char *name, const double *p, int n1, int n2, double _d, int i, int j)
{
    arr A = {p, n1, n2, 1};
    CASE("center", AT2(A, i, j)); CASE("left", AT2(A, i - 1, j)); CASE("right", AT2(A, i + 1, j));
    CASE("back", AT2(A, i, j - 1)); CASE("front", AT2(A, i, j + 1)); CASE("next", AT2(A, i + 1, j + 1));
    CASE("_d_xa", d_xa2(A, _d, i, j)); CASE("_d_ya", d_ya2(A, _d, i, j));
    CASE("_d_xi", d_xi2(A, _d, i, j)); CASE("_d_yi", d_yi2(A, _d, i, j));
    CASE("_av", av2(A, i, j)); CASE("_av_a", av_a2(A, i, j)); CASE("_av_xa", av_xa2(A, i, j)); CASE("_av_ya", av_ya2(A, i, j));
    CASE("_av_xi", av_xi2(A, i, j)); CASE("_av_yi", av_yi2(A, i, j)); CASE("_av_ai_clamped", av_ai_clamped2(A, i, j));
    CASE("_harm", harm2(A, i, j)); CASE("_harm_a", harm_a2(A, i, j)); CASE("_harm_xa", harm_xa2(A, i, j)); CASE("_harm_ya", harm_ya2(A, i, j));
    CASE("mysum", mysum2(_d != 0.0, A, i, i + 1, j, j + 1));
    /* 1-D mysum over i:j of a vector of length n1 */
    CASE("mysum1", mysum3(_d != 0.0, (arr){p, n1, 1, 1}, i, j, 1, 1, 1, 1));
    return NAN;
}
