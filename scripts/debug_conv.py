import math, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from justrelax_jl_b200 import B200Backend, PhaseRatios, PTArray, StokesArrays, setups, stokes as jst, thermal as jth
from justrelax_jl_b200.stokes3d_vc import iterate3d_VC_
from justrelax_jl_b200.types import IGG, ThermalArrays
torch.cuda.set_device(0)
dev = lambda a: PTArray(B200Backend)(a)
fin = lambda name, a: print(f"   {name}: finite={bool(torch.isfinite(a).all().item())} max={float(a.abs().max().item()):.4g}", flush=True)
for n in (int(v) for v in sys.argv[1:]):
    print("n =", n, flush=True)
    s = setups.convection3d(n, n, n)
    st = StokesArrays(B200Backend, n, n, n, vertex_normals=False)
    th = ThermalArrays(B200Backend, n, n, n)
    th.T.copy_(dev(s.T)); th.Told.copy_(th.T)
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    a = dict(T=th.T, P=st.P)
    z = lambda: dev(np.zeros(s.ni, order="F"))
    ρg = (z(), z(), z())
    jst.flow_bcs_(st, s.flow_bcs)
    pt_th = jth.PTThermalCoeffs(B200Backend, s.rheology, pr, a, s.dt, s.ni, s.di, s.li, ϵ=1e-5, CFL=s.thermal_CFL)
    fin("θr_dτ", pt_th.θr_dτ); fin("dτ_ρ", pt_th.dτ_ρ)
    for step in range(2):
        jst.compute_ρg_(ρg, pr, s.rheology, a, st); fin("ρgz", ρg[2])
        jst.compute_viscosity_(st, pr, a, s.rheology, s.kwargs["viscosity_cutoff"]); fin("η", st.viscosity.η)
        for chunk in range(2):
            iterate3d_VC_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, a, s.dt, 25, finish=(chunk == 1), kwargs=dict(viscosity_cutoff=s.kwargs["viscosity_cutoff"]))
            fin("Vz", st.V.Vz); fin("P", st.P); fin("τxx", st.τ.xx); fin("τxy", st.τ.xy); fin("η", st.viscosity.η); fin("ητ", st.viscosity.ητ)
        jth.thermal_iterate_(th, pt_th, s.thermal_bc, s.rheology, a, s.dt, s.grid, 50, kwargs=dict(phase=pr, verbose=False))
        fin("T", th.T)
