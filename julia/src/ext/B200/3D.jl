# src/ext/B200/3D.jl — the B200 methods of JustRelax.JustRelax3D (mirror of the hot-path subset of src/ext/CUDA/3D.jl:38-108,
# 195-299, 319-327, 358-390, 519-539).  Everything forwards to libjrb200 through the helpers of common.jl.
module JustRelax3D

using JustRelax: JustRelax
using JustRelaxB200
using JustRelaxB200: B200Array, B200Backend, b200zeros, b200ones
using JustPIC
using GeoParams
using ImplicitGlobalGrid
using MPI

import JustRelax.JustRelax3D as JR3D
import JustRelax: IGG, Geometry, PTArray, backend
import JustRelax: TemperatureBoundaryConditions, AbstractFlowBoundaryConditions, DisplacementBoundaryConditions, VelocityBoundaryConditions
import ..B200BackendTrait

const JRND = JR3D
const ND = 3
include("common.jl")

# Types ------------------------------------------------------------------------------------------------------------
JR3D.StokesArrays(::Type{B200Backend}, ni::NTuple{N, Integer}) where {N} = b200_stokes_arrays(ni)                 # 3D.jl:38-40
JR3D.StokesArrays(::Type{B200Backend}, ni::Vararg{Integer, N}) where {N} = b200_stokes_arrays(ni)
JR3D.ThermalArrays(::Type{B200Backend}, ni::NTuple{N, Number}) where {N} = b200_thermal_arrays(map(Int, ni))       # :55-57
JR3D.ThermalArrays(::Type{B200Backend}, ni::Vararg{Number, N}) where {N} = b200_thermal_arrays(map(Int, ni))       # :59-61

function JR3D.PTThermalCoeffs(::Type{B200Backend}, K, ρCp, dt, di::NTuple, li::NTuple; ϵ = 1.0e-8, CFL = 0.9 / √3)   # :75-79
    return pt_thermal_coeffs(K, ρCp, dt, di, li; ϵ, CFL)
end
function JR3D.PTThermalCoeffs(::Type{B200Backend}, rheology, phase_ratios, args, dt, ni, di::NTuple{nDim, T}, li::NTuple{nDim, Any};
                              ϵ = 1.0e-8, CFL = 0.9 / √3) where {nDim, T}                                           # :81-94
    rows = API.ThermalPhase[lower_thermal_phase(p) for p in rheology]
    return pt_thermal_coeffs(nothing, nothing, dt, di, li; ϵ, CFL, rows, args, phase = phase_ratios)
end
function JR3D.PTThermalCoeffs(::Type{B200Backend}, rheology::GeoParams.MaterialParams, args, dt, ni, di::NTuple, li::NTuple; ϵ = 1.0e-8, CFL = 0.9 / √3)   # :96-108
    return pt_thermal_coeffs(nothing, nothing, dt, di, li; ϵ, CFL, rows = API.ThermalPhase[lower_thermal_phase(rheology)], args)
end
function JR3D.update_pt_thermal_arrays!(pt_thermal::JustRelax.PTThermalCoeffs{T, <:B200Array}, phase_ratios::JustPIC.PhaseRatios, rheology, args, _dt) where {T}   # :110-120
    rows = API.ThermalPhase[lower_thermal_phase(p) for p in rheology]
    # the library recomputes θr_dτ, dτ_ρ from dt and the grid spacing it is given: di = Vpdτ-consistent spacing of the solve
    update_pt_arrays!(pt_thermal, inv(_dt), ntuple(_ -> pt_thermal.Vpdτ / pt_thermal.CFL, 3); rows, args, phase = phase_ratios)
    return nothing
end

# Boundary conditions ----------------------------------------------------------------------------------------------
JR3D.flow_bcs!(::B200BackendTrait, stokes::JustRelax.StokesArrays, bcs::VelocityBoundaryConditions) = flow_bcs_b200!(stokes, bcs)        # :195-206
JR3D.flow_bcs!(::B200BackendTrait, stokes::JustRelax.StokesArrays, bcs::DisplacementBoundaryConditions) = flow_bcs_b200!(stokes, bcs)    # :209-218
JR3D.thermal_bcs!(::B200BackendTrait, thermal::JustRelax.ThermalArrays, bcs) = thermal_bcs_b200!(thermal.T, bcs)                         # :220-226

# Rheology -----------------------------------------------------------------------------------------------------------
function JR3D.compute_viscosity!(::B200BackendTrait, stokes, ν, phase_ratios, args, rheology, cutoff, fn_viscosity::F) where {F}           # :237-241
    return compute_viscosity_b200!(stokes, ν, phase_ratios, args, rheology, cutoff)
end
JR3D.tensor_invariant!(::B200BackendTrait, A::JustRelax.SymmetricTensor) = tensor_invariant_b200!(A)                                      # :266-268
JR3D.accumulate_tensor!(::B200BackendTrait, II, A::JustRelax.SymmetricTensor, dt) = accumulate_tensor_b200!(II, A, dt)                    # :270-272
JR3D.accumulate_vol!(::B200BackendTrait, EVol_pl, ε_vol_pl, dt) = accumulate_vol_b200!(EVol_pl, ε_vol_pl, dt)                              # :278-280
function JR3D.compute_ρg!(ρg::NTuple{N, B200Array}, phase_ratios::JustPIC.PhaseRatios, rheology, args) where {N}                           # :290-294
    return compute_ρg_b200!(ρg, phase_ratios, rheology, args)
end
function JR3D.shear2center!(::B200BackendTrait, A::JustRelax.SymmetricTensor)                                                              # :319-322
    LIB.check(ccall(jrsym(:jr_shear2center3d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                    ctx(), A.yz_c.ptr, A.xz_c.ptr, A.xy_c.ptr, A.yz.ptr, A.xz.ptr, A.xy.ptr, Int32[size(A.xx)...]))
    return nothing
end
JR3D.velocity2displacement!(::B200BackendTrait, stokes::JustRelax.StokesArrays, dt) = velocity2displacement_b200!(stokes, dt)              # :358-360
JR3D.displacement2velocity!(::B200BackendTrait, stokes::JustRelax.StokesArrays, dt) = displacement2velocity_b200!(stokes, dt)              # :366-368
function JR3D.compute_maxloc!(B::B200Array{3}, A::B200Array{3}; window = (1, 1, 1))                                                        # Utils.jl:409-461
    LIB.check(ccall(jrsym(:jr_maxloc3d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}), ctx(), B.ptr, A.ptr, Int32[size(A)...], Int32[window...]))
    return nothing
end

# Solvers ------------------------------------------------------------------------------------------------------------
# solve!(::CUDABackendTrait, stokes, args...; kwargs) = _solve!(stokes, args...; kwargs...)   3D.jl:375-377 — the method table of Stokes3D.jl:
#   3D-VA  (stokes, pt_stokes, grid|di, flow_bcs, ρg, K, G, dt, igg)                            :25-41, 188-202
#   3D-VC  (stokes, pt_stokes, grid|di, flow_bcs, ρg, phase_ratios, rheology, args, dt, igg)    :447-466
function JR3D.solve!(::B200BackendTrait, stokes, pt_stokes, grid, flow_bcs, ρg, K::AbstractArray, G::AbstractArray, dt, igg::IGG; kwargs)
    return solve_arrays!(stokes, pt_stokes, grid, flow_bcs, ρg, K, G, dt, igg; kwargs...)
end
function JR3D.solve!(::B200BackendTrait, stokes, pt_stokes, grid, flow_bcs, ρg, phase_ratios::JustPIC.PhaseRatios, rheology::NTuple, args, dt, igg::IGG; kwargs)
    return solve_phases!(stokes, pt_stokes, grid, flow_bcs, ρg, phase_ratios, rheology, args, dt, igg; kwargs...)
end
# the legacy single-MaterialParams variants (Stokes3D.jl:204-441) are outside the backend: fail loudly
function JR3D.solve!(::B200BackendTrait, stokes, pt_stokes, grid, flow_bcs, ρg, rheology::GeoParams.MaterialParams, args...; kwargs)
    throw(ArgumentError("solve! with a single MaterialParams is outside the B200 backend; pass PhaseRatios and a 1-tuple rheology"))
end

# heatdiffusion_PT!(::CUDABackendTrait, thermal, args...; kwargs) = _heatdiffusion_PT!(thermal, args...; kwargs...)   3D.jl:383-385
function JR3D.heatdiffusion_PT!(::B200BackendTrait, thermal, pt_thermal, thermal_bc, K::AbstractArray, ρCp::AbstractArray, dt, grid; kwargs)
    return heatdiffusion_arrays!(thermal, pt_thermal, thermal_bc, K, ρCp, dt, grid; kwargs...)
end
function JR3D.heatdiffusion_PT!(::B200BackendTrait, thermal, pt_thermal, thermal_bc, rheology, args::NamedTuple, dt, grid; kwargs)
    return heatdiffusion_rheology!(thermal, pt_thermal, thermal_bc, rheology, args, dt, grid; kwargs...)
end

# Utils --------------------------------------------------------------------------------------------------------------
JR3D.compute_dt(::B200BackendTrait, S::JustRelax.StokesArrays, args...) = compute_dt_b200(S, args...)                                      # :388-390
function JR3D.update_phase_ratios_3D!(phase_ratios::JustPIC.PhaseRatios, phase_arrays::NTuple{N, B200Array{3}}, xci, xvi) where {N}        # :519-539
    return update_phase_ratios_b200!(phase_ratios, phase_arrays, xci, xvi)
end

# the stand-alone halo update on B200 arrays (the solvers exchange their own halos inside the library; the communicator is
# attached on first use: ensure_comm! in common.jl)
ImplicitGlobalGrid.update_halo!(A::Vararg{B200Array{3}, N}) where {N} = update_halo_b200!(A...)

end # module
