#!/usr/bin/env python
"""Generates the committed golden vectors of tests/golden/ (test infrastructure).

The Julia reference cannot run in this environment (no julia binary, no network: SURVEY.md §8c), so there are two kinds of goldens:
  reference_goldens.json   values TRANSCRIBED from the reference's own test-suite (file:line cited per entry) — the pins of the oracle;
  oracle_*.npz             seeded inputs → outputs of the CPU oracle (oracle/, after it passed the pins above) for every path of the
                           hot loop at small sizes.  They freeze the oracle (any later edit that changes a bit shows up) and give the
                           GPU tests committed vectors to compare with in addition to the live oracle run.
Usage: python tests/golden/make_fixtures.py        (rewrites the npz files; review the diff before committing)
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FS6 = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)


def case_va3d(po):
    from justrelax_jl_b200 import setups
    s = setups.random_stokes3d((9, 8, 7), seed=20261017)
    d = po.alloc_stokes(s.ni, s.fields)
    opts = po.make_opts(s.pt_stokes, s.grid._di.center, s.dt, FS6, s.ni, iterMax=4, nout=4)
    po.iterate3d_VA(d, s.ni, opts, 4)
    return {k: d[k] for k in ("Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "Rx", "Ry", "Rz", "RP", "etatau")}


def case_vc3d(po):
    from justrelax_jl_b200 import rheology as R, setups
    s = setups.random_vc3d((9, 8, 7), seed=20261017)
    d = po.alloc_stokes(s.ni, s.fields)
    d["Pargs"] = d["P"]
    vc = po.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), s.ratios)
    opts = po.make_opts(s.pt_stokes, s.grid._di.center, s.dt, FS6, s.ni, iterMax=4, nout=4, viscosity_relaxation=0.3, viscosity_cutoff=s.kwargs["viscosity_cutoff"])
    po.iterate3d_VC(d, s.ni, opts, vc, 4, finish=True)
    return {k: d[k] for k in ("Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "tyz_c", "txz_c", "txy_c", "eta", "etatau", "lam",
                              "Rx", "Ry", "Rz", "RP", "pyz", "pxz", "pxy", "tII", "eta_vep", "EII_pl", "rhogz", "wxy", "pxy_c")}


def case_2d(po):
    import test_gpu_stokes2d as t2
    from justrelax_jl_b200 import rheology as R
    out = {}
    ni = (12, 10)
    f, grid, pt, dt = t2.random_stokes2d(ni, 20261017)
    flags = dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6)
    d = po.alloc_stokes(ni, f)
    po.iterate2d_V2(d, ni, po.make_opts(pt, grid._di.center, dt, flags, ni, iterMax=4, nout=4), 4)
    out.update({"v2_" + k: d[k] for k in ("Vx", "Vy", "P", "txx", "tyy", "txy", "Rx", "Ry", "RP")})
    f, grid, pt, dt, rat, rheo = t2.random_vc2d(ni, 20261017, rho_var=True)
    d = po.alloc_stokes(ni, f)
    d["Pargs"] = d["P"]
    vc = po.vc_inputs(R.lower_stokes(rheo), R.gravity_of(rheo), rat)
    po.iterate2d_VC(d, ni, po.make_opts(pt, grid._di.center, dt, flags, ni, iterMax=4, nout=4, viscosity_relaxation=0.3), vc, 4, finish=True)
    out.update({"vc_" + k: d[k] for k in ("Vx", "Vy", "P", "txx", "tyy", "txy", "txy_c", "eta", "etav", "lam", "lamv", "Rx", "Ry", "RP", "tII", "EII_pl")})
    return out


def case_thermal(po):
    import test_gpu_thermal as tg
    from justrelax_jl_b200.types import Geometry
    out = {}
    for tag, ni in (("2d", (11, 9)), ("3d", (9, 8, 7))):
        li = tuple(1.0e5 * (1 + 0.1 * q) for q in range(len(ni)))
        grid = Geometry(ni, li)
        full = po.alloc_thermal(ni, tg.random_thermal(ni, 20261017, 3))
        o = po.thermal_opts(_di=grid._di.center, dt=1.0e11, eps=1e-8, iterMax=10, nout=3, max_lxyz=max(li), Vpdtau=min(grid.di.center) * 0.5, form=1,
                            phases=tg.PHASES, bc=tg.bc_variants(len(ni))[0])
        fs = po.thermal_fields(full, ni)
        for _ in range(3):
            po.lib().orc_thermal_iterate_once(C.byref(fs), C.byref(o))
        po.lib().orc_thermal_check_res(C.byref(fs), C.byref(o))
        names = ["T", "qTx", "qTy", "qTx2", "qTy2", "ResT", "theta_r_dtau", "dtau_rho"] + (["qTz", "qTz2"] if len(ni) == 3 else [])
        out.update({f"{tag}_{k}": (tg.face_only(full[k]) if k == "T" else full[k]) for k in names})
    return out


CASES = dict(oracle_va3d=case_va3d, oracle_vc3d=case_vc3d, oracle_stokes2d=case_2d, oracle_thermal=case_thermal)


def main():
    from oracle import pyoracle as po
    po.build()
    for name, fn in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fn(po))
        print("wrote", name)


if __name__ == "__main__":
    main()
