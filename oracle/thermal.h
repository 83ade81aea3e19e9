/* thermal.h — CPU ORACLE (test infrastructure) types for heatdiffusion_PT!; see thermal.c. */
#ifndef JR_ORACLE_THERMAL_H
#define JR_ORACLE_THERMAL_H
#include <stdint.h>

/* ThermalArrays (src/types/heat_diffusion.jl:1-16) + PTThermalCoeffs arrays (:30-44) + the extra inputs of the solve */
typedef struct {
    int32_t ndim;
    int32_t n[3];
    double *T, *Told, *dT;                      /* (n+2)^d */
    double *qTx, *qTy, *qTz, *qTx2, *qTy2, *qTz2;
    double *H, *shear_heating, *adiabatic, *ResT;
    double *theta_r_dtau, *dtau_rho;            /* pt_thermal.θr_dτ, pt_thermal.dτ_ρ */
    double *K, *rhoCp;                          /* array form */
    double *P;                                  /* args.P (rheology form); args.T is T itself */
    double *dir_mask, *dir_value;               /* Dirichlet mask / value arrays on the ghosted grid, or NULL */
    double *phase_c, *phase_x, *phase_y, *phase_z; /* phase ratios [phase][node] (centre, Vx, Vy, Vz), or NULL */
} orc_thermal_fields;

typedef struct {
    int32_t rho_kind;   /* 0 ConstantDensity, 1 PT_Density, 2 T_Density */
    int32_t has_Hr;
    double rho0, alpha, beta, T0, P0, Cp, k, Hr;
    int32_t k_kind;   /* 0 ConstantConductivity, 1 TP_Conductivity: k = (k_a + k_b / (T + k_c)) (1 + k_d P)  (GeoParams, restated from its docstring) */
    int32_t _pad;
    double k_a, k_b, k_c, k_d;
} orc_thermal_phase;

typedef struct {
    double _di[3], dt, eps;
    int64_t iterMax, nout;
    double max_lxyz, Vpdtau;
    int32_t form;       /* 0: K, ρCp arrays; 1: rheology table */
    int32_t nphase;
    const orc_thermal_phase *phases;
    double dir_const;
    /* faces in the order left,right,front,back,top,bot */
    int32_t no_flux[6], cv_active[6], cf_active[6], periodic[6];
    double cv_value[6], cf_value[6];
} orc_thermal_opts;

typedef struct {
    int64_t iter, nhist, cap;
    double err;
    double *norm_ResT;
    int64_t *iter_count;
} orc_thermal_result;

void orc_thermal_pt_arrays(const orc_thermal_fields *f, const orc_thermal_opts *o);
void orc_thermal_flux(const orc_thermal_fields *f, const orc_thermal_opts *o);
void orc_thermal_update_T(const orc_thermal_fields *f, const orc_thermal_opts *o);
void orc_thermal_check_res(const orc_thermal_fields *f, const orc_thermal_opts *o);
void orc_thermal_bcs(const orc_thermal_fields *f, const orc_thermal_opts *o);
void orc_thermal_adiabatic(const orc_thermal_fields *f, const orc_thermal_opts *o, const double *P, const double *P0);
void orc_thermal_iterate_once(const orc_thermal_fields *f, const orc_thermal_opts *o);
int orc_heatdiffusion_PT(const orc_thermal_fields *f, const orc_thermal_opts *o, const double *stokes_P, const double *stokes_P0,
                         orc_thermal_result *res);
#endif
