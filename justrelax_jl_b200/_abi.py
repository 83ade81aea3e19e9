"""ctypes binding of include/jrb200.h (libjrb200.so, CUDA sm_100a).

This is the Python twin of the `ccall` stubs a Julia `JustRelaxB200Ext` would hold
(INTEGRATION.md).  Loading fails loudly when the CUDA library is missing: there is no
CPU fallback behind this module.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjrb200.so")


class JRError(RuntimeError):
    """Error raised for a negative jr_status (message = jr_last_error())."""

    def __init__(self, status: int, msg: str):
        super().__init__(f"libjrb200 status {status}: {msg}")
        self.status = status
        self.msg = msg


JR_OK, JR_ERR_CUDA, JR_ERR_SHAPE, JR_ERR_NAN, JR_ERR_UNSUPPORTED, JR_ERR_NCCL, JR_ERR_ARG = 0, -1, -2, -3, -4, -5, -6
JR_FLAG_UNFUSED = 1
JR_FLAG_DIAG_EVERY_ITER = 2

_lib = None
_field_names = None


def lib():
    """Load libjrb200.so (built by `make -C justrelax_jl_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). justrelax_jl_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.jr_last_error.restype = C.c_char_p
    L.jr_field_name.restype = C.c_char_p
    L.jr_field_name.argtypes = [C.c_int]
    _lib = L
    _declare(L)
    return L


def field_names():
    global _field_names
    if _field_names is None:
        L = lib()
        _field_names = [L.jr_field_name(i).decode() for i in range(L.jr_field_count())]
    return _field_names


def check(status: int):
    if status != 0:
        raise JRError(status, lib().jr_last_error().decode())


def make_fields_struct(nfields: int):
    class Fields(C.Structure):
        _fields_ = [("ndim", C.c_int32), ("n", C.c_int32 * 3), ("f", C.c_void_p * nfields)]

    return Fields


class StokesOpts(C.Structure):
    _fields_ = [
        ("r", C.c_double), ("theta_dtau", C.c_double), ("eta_dtau", C.c_double),
        ("eps_rel", C.c_double), ("eps_abs", C.c_double),
        ("_di", C.c_double * 3),
        ("dt", C.c_double),
        ("iterMax", C.c_int64), ("nout", C.c_int64),
        ("n_g", C.c_int32 * 3),
        ("free_slip", C.c_int32 * 6), ("no_slip", C.c_int32 * 6), ("periodic", C.c_int32 * 6),
        ("viscosity_relaxation", C.c_double), ("lambda_relaxation", C.c_double),
        ("visc_cutoff_lo", C.c_double), ("visc_cutoff_hi", C.c_double),
        ("iterMin", C.c_int64),
        ("strain_rate_ni_only", C.c_int32), ("strain_increment", C.c_int32), ("displacement_bcs", C.c_int32), ("dT_ghosted", C.c_int32),
    ]


class StokesResult(C.Structure):
    _fields_ = [
        ("iter", C.c_int64), ("nhist", C.c_int64), ("err", C.c_double),
        ("err_evo1", C.POINTER(C.c_double)), ("err_evo2", C.POINTER(C.c_int64)),
        ("norm_Rx", C.POINTER(C.c_double)), ("norm_Ry", C.POINTER(C.c_double)),
        ("norm_Rz", C.POINTER(C.c_double)), ("norm_divV", C.POINTER(C.c_double)),
        ("time_s", C.c_double), ("kernel_launches", C.c_int64),
    ]


class OracleStokesResult(C.Structure):
    """orc_stokes_result (oracle/jr_oracle.h) — same head, no timing tail."""

    _fields_ = StokesResult._fields_[:9]


class ThermalFields(C.Structure):
    """jr_thermal_fields"""
    names = ("T", "Told", "dT", "qTx", "qTy", "qTz", "qTx2", "qTy2", "qTz2", "H", "shear_heating", "adiabatic", "ResT",
             "theta_r_dtau", "dtau_rho", "K", "rhoCp", "P", "dir_mask", "dir_value", "phase_c", "phase_x", "phase_y", "phase_z")
    _fields_ = [("ndim", C.c_int32), ("n", C.c_int32 * 3)] + [(nm, C.c_void_p) for nm in names]


class ThermalPhase(C.Structure):
    _fields_ = [("rho_kind", C.c_int32), ("has_Hr", C.c_int32), ("rho0", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
                ("T0", C.c_double), ("P0", C.c_double), ("Cp", C.c_double), ("k", C.c_double), ("Hr", C.c_double),
                ("k_kind", C.c_int32), ("_pad", C.c_int32), ("k_a", C.c_double), ("k_b", C.c_double), ("k_c", C.c_double), ("k_d", C.c_double)]


class ThermalOpts(C.Structure):
    _fields_ = [("_di", C.c_double * 3), ("dt", C.c_double), ("eps", C.c_double), ("iterMax", C.c_int64), ("nout", C.c_int64),
                ("max_lxyz", C.c_double), ("Vpdtau", C.c_double), ("form", C.c_int32), ("nphase", C.c_int32),
                ("phases", C.POINTER(ThermalPhase)), ("dir_const", C.c_double),
                ("no_flux", C.c_int32 * 6), ("cv_active", C.c_int32 * 6), ("cf_active", C.c_int32 * 6), ("periodic", C.c_int32 * 6),
                ("cv_value", C.c_double * 6), ("cf_value", C.c_double * 6)]


class ThermalResult(C.Structure):
    _fields_ = [("iter", C.c_int64), ("nhist", C.c_int64), ("cap", C.c_int64), ("err", C.c_double),
                ("norm_ResT", C.POINTER(C.c_double)), ("iter_count", C.POINTER(C.c_int64)),
                ("time_s", C.c_double), ("kernel_launches", C.c_int64)]


class StokesPhase(C.Structure):
    """jr_stokes_phase"""
    _fields_ = [("eta", C.c_double), ("G", C.c_double), ("Kb", C.c_double), ("has_pl", C.c_int32), ("rho_kind", C.c_int32),
                ("C", C.c_double), ("sinphi", C.c_double), ("cosphi", C.c_double), ("sinpsi", C.c_double), ("eta_vp", C.c_double),
                ("rho0", C.c_double), ("alpha", C.c_double), ("beta", C.c_double), ("T0", C.c_double), ("P0", C.c_double),
                ("soft_C_kind", C.c_int32), ("_pad", C.c_int32), ("soft_C", C.c_double * 6)]


class VcInputs(C.Structure):
    """jr_vc_inputs"""
    _fields_ = [("nphase", C.c_int32), ("g_scalar", C.c_int32), ("phases", C.POINTER(StokesPhase)), ("g", C.c_double * 3),
                ("ph_center", C.c_void_p), ("ph_vertex", C.c_void_p), ("ph_xy", C.c_void_p), ("ph_yz", C.c_void_p), ("ph_xz", C.c_void_p),
                ("free_surface", C.c_double)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def _declare(L):
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    L.jr_context_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.jr_context_destroy.argtypes = [vp]
    L.jr_context_set_flags.argtypes = [vp, C.c_uint32]
    L.jr_context_synchronize.argtypes = [vp]
    L.jr_malloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.jr_free.argtypes = [vp, vp]
    L.jr_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.jr_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.jr_memcpy_d2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.jr_fill_f64.argtypes = [vp, vp, C.c_double, C.c_size_t]
    L.jr_scale_copy.argtypes = [vp, vp, vp, C.c_double, C.c_size_t]
    L.jr_sumsq.argtypes = [vp, vp, i32p, C.c_int, C.POINTER(C.c_double)]
    L.jr_flow_bcs3d.argtypes = [vp, vp, vp, vp, i32p, i32p, i32p, i32p]
    L.jr_maxloc3d.argtypes = [vp, vp, vp, i32p, i32p]
    L.jr_stokes3d_solve_VA.argtypes = [vp, vp, C.POINTER(StokesOpts), C.POINTER(StokesResult)]
    L.jr_stokes3d_iterate_VA.argtypes = [vp, vp, C.POINTER(StokesOpts), C.c_int64, C.POINTER(StokesResult)]
    L.jr_stokes3d_VA_plan_info.argtypes = [vp, i32p]
    L.jr_stokes3d_VA_begin.argtypes = [vp, vp, C.POINTER(StokesOpts)]
    L.jr_stokes3d_VA_step.argtypes = [vp, C.c_int64, C.c_int, C.POINTER(StokesResult)]
    L.jr_stokes3d_VA_end.argtypes = [vp]
    so, sr, vc = C.POINTER(StokesOpts), C.POINTER(StokesResult), C.POINTER(VcInputs)
    L.jr_stokes2d_solve_V2.argtypes = [vp, vp, so, sr]
    L.jr_stokes2d_iterate_V2.argtypes = [vp, vp, so, C.c_int64, sr]
    L.jr_stokes2d_solve_VC.argtypes = [vp, vp, so, vc, sr]
    L.jr_stokes2d_iterate_VC.argtypes = [vp, vp, so, vc, C.c_int64, C.c_int, sr]
    L.jr_flow_bcs2d.argtypes = [vp, vp, vp, i32p, i32p, i32p, i32p]
    L.jr_compute_viscosity2d.argtypes = [vp, vp, so, vc, C.c_double]
    L.jr_compute_rhog2d.argtypes = [vp, vp, vc]
    L.jr_tensor_invariant2d.argtypes = [vp, vp, vp, vp, vp, i32p]
    L.jr_stokes3d_solve_VC.argtypes = [vp, vp, so, vc, sr]
    L.jr_stokes3d_iterate_VC.argtypes = [vp, vp, so, vc, C.c_int64, C.c_int, sr]
    L.jr_compute_viscosity3d.argtypes = [vp, vp, so, vc, C.c_double]
    L.jr_compute_rhog3d.argtypes = [vp, vp, vc]
    L.jr_tensor_invariant3d.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32p]
    L.jr_shear2center3d.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32p]
    L.jr_phase_ratios_from_arrays.argtypes = [vp, C.c_int32, i32p, C.c_int32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp, vp, vp, vp]
    L.jr_accumulate_tensor2d.argtypes = [vp, vp, vp, vp, vp, i32p, C.c_double]
    L.jr_accumulate_tensor3d.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32p, C.c_double]
    L.jr_accumulate_vol.argtypes = [vp, vp, vp, C.c_size_t, C.c_double]
    L.jr_absmax.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    L.jr_velocity2vertex.argtypes = [vp, C.c_int32, i32p, i32p, vp, vp, vp, vp, vp, vp]
    L.jr_velocity2center.argtypes = [vp, C.c_int32, i32p, i32p, vp, vp, vp, vp, vp, vp]
    L.jr_lithostatic_pressure.argtypes = [vp, C.c_int32, i32p, vp, vp, C.c_double, vp, C.c_int, C.c_int32]
    L.jr_compute_shear_heating.argtypes = [vp, vp, vc, C.POINTER(C.c_double), C.c_double, vp]
    L.jr_heatdiffusion_PT.argtypes = [vp, C.POINTER(ThermalFields), C.POINTER(ThermalOpts), vp, vp, C.POINTER(ThermalResult)]
    L.jr_thermal_iterate.argtypes = [vp, C.POINTER(ThermalFields), C.POINTER(ThermalOpts), C.c_int64, C.POINTER(ThermalResult)]
    L.jr_thermal_bcs.argtypes = [vp, vp, C.c_int32, i32p, C.POINTER(ThermalOpts)]
    L.jr_thermal_pt_arrays.argtypes = [vp, C.POINTER(ThermalFields), C.POINTER(ThermalOpts)]
    L.jr_comm_create.argtypes = [vp, C.c_int, C.c_int, i32p, i32p, ALLGATHER_FN, vp, C.POINTER(vp)]
    L.jr_comm_create_periodic.argtypes = [vp, C.c_int, C.c_int, i32p, i32p, i32p, ALLGATHER_FN, vp, C.POINTER(vp)]
    L.jr_comm_destroy.argtypes = [vp]
    L.jr_context_set_comm.argtypes = [vp, vp]
    L.jr_comm_barrier.argtypes = [vp]
    L.jr_update_halo3d.argtypes = [vp, C.c_int, C.POINTER(vp), i32p, i32p]
    L.jr_allreduce_f64.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_int]
    L.jr_halo_source.argtypes = [i32p, i32p, i32p, i32p, i32p, i32p, i32p]
    L.jr_halo_source_periodic.argtypes = [i32p, i32p, i32p, i32p, i32p, i32p, i32p, i32p]


def i32x(vals):
    return (C.c_int32 * len(vals))(*[int(v) for v in vals])
