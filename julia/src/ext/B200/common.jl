# common.jl — dimension-independent host code of the B200 backend, included by the JustRelax2D / JustRelax3D extension
# modules (the place src/ext/CUDA/{2D,3D}.jl include src/common.jl from).  Expects in scope: `JRND` (the owning
# JustRelax2D / JustRelax3D module), `ND` (2 or 3), JustRelax, JustRelaxB200 (+ API), JustPIC, GeoParams, MPI,
# ImplicitGlobalGrid.
#
# Everything here marshals arguments into the C ABI; there is no numerical kernel on the Julia side.

const API = JustRelaxB200.API
const LIB = JustRelaxB200
@inline jrsym(s::Symbol) = JustRelaxB200.sym(s)
@inline ctx() = JustRelaxB200.context()

# =====================================================================================================================
# constructors: the reference builds every container with ParallelStencil's @zeros/@ones (src/types/constructors/*.jl);
# here the same shapes are allocated as B200Arrays and handed to the reference's own struct constructors.
z(dims...) = b200zeros(dims...)
o(dims...) = b200ones(dims...)

function b200_velocity(ni::NTuple{2})
    nx, ny = ni
    return JustRelax.Velocity(z(nx + 1, ny + 2), z(nx + 2, ny + 1), nothing)                 # constructors/stokes.jl:10-16
end
function b200_velocity(ni::NTuple{3})
    nx, ny, nz = ni
    return JustRelax.Velocity(z(nx + 1, ny + 2, nz + 2), z(nx + 2, ny + 1, nz + 2), z(nx + 2, ny + 2, nz + 1))   # :26-33
end
function b200_displacement(ni::NTuple{N}) where {N}
    v = b200_velocity(ni)
    return JustRelax.Displacement(v.Vx, v.Vy, v.Vz)                                          # :42-66
end
b200_vorticity(ni::NTuple{2}) = JustRelax.Vorticity(nothing, nothing, z(ni[1] + 1, ni[2] + 1))                  # :74-78
b200_vorticity(ni::NTuple{3}) = JustRelax.Vorticity(z(ni[1], ni[2] + 1, ni[3] + 1), z(ni[1] + 1, ni[2], ni[3] + 1), z(ni[1] + 1, ni[2] + 1, ni[3]))
b200_viscosity(ni::NTuple) = JustRelax.Viscosity(o(ni...), o((ni .+ 1)...), o(ni...), z(ni...))               # :99-105
function b200_tensor(ni::NTuple{2})
    nx, ny = ni
    return JustRelax.SymmetricTensor(z(nx, ny), z(nx, ny), z(nx + 1, ny + 1), z(nx + 1, ny + 1), z(nx + 1, ny + 1), z(nx, ny), z(nx, ny))   # :164-174
end
function b200_tensor(ni::NTuple{3})
    nx, ny, nz = ni
    v() = z(nx + 1, ny + 1, nz + 1)
    return JustRelax.SymmetricTensor(z(nx, ny, nz), z(nx, ny, nz), z(nx, ny, nz), v(), v(), v(), z(nx + 1, ny + 1, nz), z(nx, ny + 1, nz + 1),
                                     z(nx + 1, ny, nz + 1), z(nx, ny, nz), z(nx, ny, nz), z(nx, ny, nz), z(nx, ny, nz))                   # :196-212
end
b200_residual(ni::NTuple{2}) = JustRelax.Residual(z(ni...), z(ni[1] - 1, ni[2]), z(ni[1], ni[2] - 1))          # :224-229
b200_residual(ni::NTuple{3}) = JustRelax.Residual(z(ni...), z(ni[1] - 1, ni[2], ni[3]), z(ni[1], ni[2] - 1, ni[3]), z(ni[1], ni[2], ni[3] - 1))

"StokesArrays(B200Backend, ni) — src/ext/CUDA/3D.jl:38-40 → constructors/stokes.jl:278-302 (same field order)"
function b200_stokes_arrays(ni::NTuple{N, Integer}) where {N}
    ni = map(Int, ni)
    return JustRelax.StokesArrays(
        z(ni...), z(ni...), b200_velocity(ni), z(ni...), z(ni...), b200_tensor(ni), b200_tensor(ni), b200_tensor(ni), z(ni...), z(ni...),
        z(ni...), b200_viscosity(ni), b200_tensor(ni), b200_residual(ni), b200_displacement(ni), b200_vorticity(ni), b200_tensor(ni),
        z(ni...), z(ni...), z((ni .+ 1)...), z(ni...),
    )
end

"ThermalArrays(B200Backend, ni...) — src/ext/CUDA/3D.jl:55-61 → constructors/heat_diffusion.jl:38-120"
function b200_thermal_arrays(ni::NTuple{N, Integer}) where {N}
    ni = map(Int, ni)
    g = ni .+ 2
    face(d) = z(ntuple(q -> ni[q] + (q == d ? 1 : 0), Val(N))...)
    qz = N == 3 ? face(3) : nothing
    qz2 = N == 3 ? face(3) : nothing
    return JustRelax.ThermalArrays(z(g...), z(g...), z(g...), z(ni...), z(ni...), face(1), face(2), qz, face(1), face(2), qz2, z(ni...), z(ni...), z(ni...))
end

# =====================================================================================================================
# field table
const TENSOR_SLOTS = (
    (:xx, "xx"), (:yy, "yy"), (:zz, "zz"), (:yz, "yz"), (:xz, "xz"), (:xy, "xy"), (:yz_c, "yz_c"), (:xz_c, "xz_c"), (:xy_c, "xy_c"), (:II, "II"),
)

"name → array of every StokesArrays member the library knows (slot names: include/jrb200.h JR_STOKES_FIELDS)"
function stokes_slots(stokes::JustRelax.StokesArrays)
    d = Dict{String, Any}()
    put!(k, v) = (v === nothing || (d[k] = v); nothing)
    put!("P", stokes.P); put!("P0", stokes.P0); put!("divV", stokes.∇V); put!("Q", stokes.Q)
    put!("Vx", stokes.V.Vx); put!("Vy", stokes.V.Vy); put!("Vz", stokes.V.Vz)
    put!("Ux", stokes.U.Ux); put!("Uy", stokes.U.Uy); put!("Uz", stokes.U.Uz)
    for (pre, T, suf) in (("t", stokes.τ, ""), ("e", stokes.ε, ""), ("p", stokes.ε_pl, ""), ("d", stokes.Δε, ""))
        for (fld, nm) in TENSOR_SLOTS
            put!(pre * nm * suf, getfield(T, fld))
        end
    end
    if stokes.τ_o !== nothing
        To = stokes.τ_o
        put!("txx_o", To.xx); put!("tyy_o", To.yy); put!("tzz_o", To.zz); put!("tyz_o", To.yz); put!("txz_o", To.xz); put!("txy_o", To.xy)
        put!("tyz_o_c", To.yz_c); put!("txz_o_c", To.xz_c); put!("txy_o_c", To.xy_c); put!("tII_o", To.II)
        ND == 2 && (put!("txx_o_v", To.xx_v); put!("tyy_o_v", To.yy_v))
    end
    ND == 2 && (put!("txx_v", stokes.τ.xx_v); put!("tyy_v", stokes.τ.yy_v))
    put!("EII_pl", stokes.EII_pl); put!("EVol_pl", stokes.EVol_pl); put!("e_vol_pl", stokes.ε_vol_pl)
    put!("eta", stokes.viscosity.η); put!("etav", stokes.viscosity.ηv); put!("eta_vep", stokes.viscosity.η_vep); put!("etatau", stokes.viscosity.ητ)
    put!("Rx", stokes.R.Rx); put!("Ry", stokes.R.Ry); put!("Rz", stokes.R.Rz); put!("RP", stokes.R.RP)
    put!("wyz", stokes.ω.yz); put!("wxz", stokes.ω.xz); put!("wxy", stokes.ω.xy)
    put!("divU", stokes.∇U); put!("lam", stokes.λ); put!("lamv", stokes.λv); put!("dPpsi", stokes.ΔPψ)
    return d
end

function add_ρg!(d, ρg)
    names = ("rhogx", "rhogy", "rhogz")
    for q in 1:length(ρg)
        d[names[q]] = LIB.ondevice(ρg[q])          # host arrays (PS-Threads @zeros in user scripts) are uploaded here
    end
    return d
end

"""
    add_args!(d, args)

`args::NamedTuple` of the multiphase solves (Stokes2D.jl:577-599, Stokes3D.jl:447-466): `T` (ni.+2), `P` (ni), `ΔT` (ni,
selects the thermal-stress form of compute_P!: PressureKernels.jl:128-149), `dt` (never read by the lowered laws).  Any
other key selects behaviour the backend does not have (melt_fraction: PressureKernels.jl:151-176, perturbation_C:
StressUpdate.jl:146-176, …) and throws instead of being dropped.
"""
function add_args!(d, args::NamedTuple)
    for (k, v) in pairs(args)
        (k === :dt || k === :perturbation_C) && continue     # perturbation_C: unused by the reference's stress kernels of this version too
        v === nothing && continue
        slot = k === :T ? "T" : k === :P ? "Pargs" : k === :ΔT ? "dTargs" :
            throw(ArgumentError("args.$k is not supported by the B200 backend (supported keys: T, P, ΔT, dt, perturbation_C); refusing to ignore it"))
        d[slot] = LIB.ondevice(v)
    end
    return d
end

# =====================================================================================================================
# options
"inverse spacings of a uniform grid: Geometry (_di.center, src/grid/Cartesian.jl:42-58) or the legacy `di` tuple / NamedTuple"
_inv_spacing(grid::JustRelax.Geometry) = map(Float64, grid._di.center)
_inv_spacing(di::NTuple) = map(x -> 1.0 / Float64(x), di)
_inv_spacing(di::NamedTuple) = _inv_spacing(di.center)
function _require_uniform(grid::JustRelax.Geometry)
    # non-uniform spacing (vector `_di`, src/grid/Cartesian.jl:76-99) is outside the supported subset: fail loudly
    all(x -> x isa Number, grid._di.center) || throw(ArgumentError("the B200 backend supports uniform grids only"))
    return grid
end
_require_uniform(di) = di

"nx_g(), ny_g(), nz_g() (ImplicitGlobalGrid), or the local size without an initialised global grid"
function global_size(ni::NTuple{N}) where {N}
    ImplicitGlobalGrid.grid_is_initialized() || return map(Int32, API.tuple3(ni, 1))
    f = (ImplicitGlobalGrid.nx_g, ImplicitGlobalGrid.ny_g, ImplicitGlobalGrid.nz_g)
    return ntuple(d -> Int32(d <= N ? f[d]() : 1), Val(3))
end

function stokes_opts(pt::JustRelax.PTStokesCoeffs, grid, dt, bcs, ni; iterMax = 10.0e3, nout = 500, viscosity_relaxation = 1.0e-2,
                     λ_relaxation = 0.2, viscosity_cutoff = (-Inf, Inf), iterMin = 0, strain_increment = false, dT_ghosted = false, kw...)
    _require_uniform(grid)
    return API.StokesOpts(
        pt.r, pt.θ_dτ, pt.ηdτ, pt.ϵ_rel, pt.ϵ_abs, API.tuple3(_inv_spacing(grid), 0.0), Float64(dt), Int64(floor(iterMax)), Int64(floor(nout)),
        global_size(ni), API.flags6(bcs.free_slip), API.flags6(bcs.no_slip), API.flags6(bcs.periodic),
        Float64(viscosity_relaxation), Float64(λ_relaxation), Float64(viscosity_cutoff[1]), Float64(viscosity_cutoff[2]), Int64(floor(iterMin)), Int32(0),
        Int32(strain_increment === true), Int32(bcs isa JustRelax.DisplacementBoundaryConditions), Int32(dT_ghosted === true),
    )
end

function check_flow_bcs_type(bcs; allow_displacement = false)
    bcs isa JustRelax.AbstractFlowBoundaryConditions || throw(ArgumentError("Unknown boundary conditions type: $(typeof(bcs))"))   # types/displacement.jl:68-70
    (bcs isa JustRelax.DisplacementBoundaryConditions && !allow_displacement) &&
        throw(ArgumentError("DisplacementBoundaryConditions are supported by the multiphase 2D solve (2D-VC) only"))
    return bcs
end

# =====================================================================================================================
# rheology lowering: rheology::NTuple{N,MaterialParams} → flat per-phase rows (SURVEY.md Appendix C).  Anything outside the
# subset throws at lowering time; nothing falls back.
_val(x) = x isa GeoParams.GeoUnit ? Float64(GeoParams.NumValue(x)) : Float64(x)
_elements(p) = isempty(p.CompositeRheology) ? () : p.CompositeRheology[1].elements       # StressUpdate.jl:128-129
_is_plastic(e) = e isa GeoParams.DruckerPrager || e isa GeoParams.DruckerPrager_regularised
function _modulus(x)                                                                    # rheology/GeoParams.jl:1-15
    v = Float64(x)
    return (isnan(v) || iszero(v)) ? Inf : v
end

function lower_density(p)
    isempty(p.Density) && return (Int32(0), 0.0, 0.0, 0.0, 0.0, 0.0)
    ρ = p.Density[1]                                                                    # rheology/GeoParams.jl:50-51
    ρ isa GeoParams.ConstantDensity && return (Int32(0), _val(ρ.ρ), 0.0, 0.0, 0.0, 0.0)
    ρ isa GeoParams.PT_Density && return (Int32(1), _val(ρ.ρ0), _val(ρ.α), _val(ρ.β), _val(ρ.T0), _val(ρ.P0))
    ρ isa GeoParams.T_Density && return (Int32(2), _val(ρ.ρ0), _val(ρ.α), 0.0, _val(ρ.T0), 0.0)
    throw(ArgumentError("density law $(nameof(typeof(ρ))) is outside the B200 backend's supported subset"))
end

function lower_phase(p::GeoParams.MaterialParams)
    els = _elements(p)
    visc = filter(e -> e isa GeoParams.LinearViscous, collect(els))
    other = filter(e -> !(e isa GeoParams.LinearViscous || e isa GeoParams.ConstantElasticity || _is_plastic(e)), collect(els))
    isempty(other) || throw(ArgumentError("rheological element $(nameof(typeof(first(other)))) is outside the B200 backend's supported subset"))
    length(visc) == 1 || throw(ArgumentError("exactly one LinearViscous element per phase is supported"))
    G = _modulus(JustRelax.get_shear_modulus(p))                                        # rheology/GeoParams.jl:1-15 (Inf when absent)
    Kb = _modulus(JustRelax.get_bulk_modulus(p))
    pls = filter(_is_plastic, collect(els))
    has_pl, C, sϕ, cϕ, sψ, ηvp = Int32(0), 0.0, 0.0, 0.0, 0.0, 0.0
    skind, spar = Int32(0), (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    if !isempty(pls)
        pl = first(pls)                                                                 # the FIRST plastic element wins: StressUpdate.jl:131-144
        pl.softening_ϕ isa GeoParams.NoSoftening || throw(ArgumentError("friction-angle softening is outside the B200 backend's supported subset"))
        sc = pl.softening_C
        if sc isa GeoParams.LinearSoftening
            ND == 2 || throw(ArgumentError("cohesion softening is supported by the 2D multiphase solve only"))
            skind = Int32(1)
            spar = (Float64(sc.lo), Float64(sc.hi), Float64(sc.max_value), Float64(sc.min_value), Float64(sc.slope), Float64(sc.ordinate))
        elseif sc isa GeoParams.NonLinearSoftening
            ND == 2 || throw(ArgumentError("cohesion softening is supported by the 2D multiphase solve only"))
            skind = Int32(2)
            spar = (Float64(sc.ξ₀), Float64(sc.Δ), Float64(sc.μ), Float64(sc.σ), 0.0, 0.0)
        elseif !(sc isa GeoParams.NoSoftening)
            throw(ArgumentError("softening law $(nameof(typeof(sc))) is outside the B200 backend's supported subset"))
        end
        ηvp_val = pl isa GeoParams.DruckerPrager_regularised ? _val(pl.η_vp) : 0.0
        has_pl, C, sϕ, cϕ, sψ, ηvp = Int32(1), _val(pl.C), _val(pl.sinϕ), _val(pl.cosϕ), _val(pl.sinΨ), ηvp_val     # StressUpdate.jl:139
    end
    kind, ρ0, α, β, T0, P0 = lower_density(p)
    return API.StokesPhase(_val(first(visc).η), G, Kb, has_pl, kind, C, sϕ, cϕ, sψ, ηvp, ρ0, α, β, T0, P0, skind, Int32(0), spar)
end

function lower_thermal_phase(p::GeoParams.MaterialParams)
    kind, ρ0, α, β, T0, P0 = lower_density(p)
    (length(p.HeatCapacity) == 1 && p.HeatCapacity[1] isa GeoParams.ConstantHeatCapacity) ||
        throw(ArgumentError("heat-capacity law outside the B200 backend's supported subset (ConstantHeatCapacity)"))
    length(p.Conductivity) == 1 || throw(ArgumentError("exactly one conductivity law per phase"))
    κ = p.Conductivity[1]
    k, k_kind, ka, kb, kc, kd = if κ isa GeoParams.ConstantConductivity
        _val(κ.k), Int32(0), 0.0, 0.0, 0.0, 0.0
    elseif κ isa GeoParams.TP_Conductivity           # k = (a + b / (T + c)) (1 + d P)   Layered_rheology.jl:45-57
        0.0, Int32(1), _val(κ.a), _val(κ.b), _val(κ.c), _val(κ.d)
    else
        throw(ArgumentError("conductivity law outside the B200 backend's supported subset (ConstantConductivity, TP_Conductivity)"))
    end
    has_Hr, Hr = Int32(0), 0.0
    if !isempty(p.RadioactiveHeat)                                                      # DiffusionPT_GeoParams.jl:145
        p.RadioactiveHeat[1] isa GeoParams.ConstantRadioactiveHeat ||
            throw(ArgumentError("radioactive-heat law outside the B200 backend's supported subset (ConstantRadioactiveHeat)"))
        has_Hr, Hr = Int32(1), _val(p.RadioactiveHeat[1].H_r)
    end
    return API.ThermalPhase(kind, has_Hr, ρ0, α, β, T0, P0, _val(p.HeatCapacity[1].Cp), k, Hr, k_kind, Int32(0), ka, kb, kc, kd)
end

"compute_gravity(first(rheology)) (BuoyancyForces.jl:25,56): a Number acting along the last axis"
function gravity3(rheology)
    g = GeoParams.compute_gravity(first(rheology))
    return g isa Number ? (0.0, 0.0, Float64(g)) : API.tuple3(map(Float64, g), 0.0)
end

# =====================================================================================================================
# phase ratios: JustPIC.PhaseRatios holds CellArrays.  On this backend their `data` is a B200Array laid out like CellArrays'
# GPU layout (blocklength 0): size(data) == (nodes, nphases, 1), node index fastest — which is exactly the library's
# [phase][node] layout, so the pointer is passed as it is.  A host CellArray (blocklength 1: size(data) == (1, nphases, nodes))
# is transposed and uploaded.
function flat(ca, keep::Vector{Any})
    ca === nothing && return Ptr{Float64}(C_NULL)
    data = ca.data
    if data isa B200Array
        size(data, 3) == 1 || throw(ArgumentError("B200 phase-ratio CellArrays must use the struct-of-arrays layout (blocklength 0)"))
        return data.ptr
    end
    h = Array(data)
    dev = size(h, 1) == 1 ? B200Array(permutedims(reshape(h, size(h, 2), size(h, 3)))) : B200Array(reshape(h, size(h, 1), size(h, 2)))
    push!(keep, dev)
    return dev.ptr
end
_maybe(pr, f::Symbol) = hasproperty(pr, f) ? getproperty(pr, f) : nothing

"jr_vc_inputs + the Julia objects that must outlive the call"
function vc_inputs(rheology::NTuple{N, GeoParams.MaterialParams}, phase_ratios; free_surface = 0.0) where {N}
    N <= 8 || throw(ArgumentError("at most 8 phases are supported"))
    rows = API.StokesPhase[lower_phase(p) for p in rheology]
    keep = Any[rows]
    vc = API.VcInputs(
        Int32(N), Int32(GeoParams.compute_gravity(first(rheology)) isa Number), pointer(rows), gravity3(rheology),
        flat(phase_ratios.center, keep), flat(_maybe(phase_ratios, :vertex), keep), flat(_maybe(phase_ratios, :xy), keep),
        flat(_maybe(phase_ratios, :yz), keep), flat(_maybe(phase_ratios, :xz), keep), Float64(free_surface),
    )
    return Ref(vc), keep
end

# =====================================================================================================================
# results
function stokes_named_tuple(h::API.History)
    r = h.res[]
    n = Int(r.nhist)
    base = (iter = Int(r.iter), err_evo1 = h.err_evo1[1:n], err_evo2 = h.err_evo2[1:n], norm_Rx = h.norm_Rx[1:n], norm_Ry = h.norm_Ry[1:n])
    if ND == 3                                                                          # Stokes3D.jl:175-185, 657-667
        av = r.time_s / max(r.iter - 1, 1)
        return merge(base, (norm_Rz = h.norm_Rz[1:n], norm_∇V = h.norm_divV[1:n], time = r.time_s, av_time = av))
    end
    return merge(base, (norm_∇V = h.norm_divV[1:n],))                                   # Stokes2D.jl:858-865
end

function print_history(out, igg, verbose)
    (verbose && igg.me == 0) || return nothing
    for c in eachindex(out.err_evo1)
        if ND == 3
            JustRelax.@printf("iter = %d, abs_err = %1.3e [norm_Rx=%1.3e, norm_Ry=%1.3e, norm_Rz=%1.3e, norm_∇V=%1.3e] \n",
                              out.err_evo2[c], out.err_evo1[c], out.norm_Rx[c], out.norm_Ry[c], out.norm_Rz[c], out.norm_∇V[c])
        else
            JustRelax.@printf("Iteration = %d, err = %1.3e [norm_Rx=%1.3e, norm_Ry=%1.3e, norm_∇V=%1.3e] \n",
                              out.err_evo2[c], out.err_evo1[c], out.norm_Rx[c], out.norm_Ry[c], out.norm_∇V[c])
        end
    end
    return nothing
end

# =====================================================================================================================
# solve!  (src/ext/CUDA/3D.jl:375-377 → the `_solve!` method table of src/stokes/Stokes{2,3}D.jl)
"visco-elastic variants with K, G arrays: 3D-VA (Stokes3D.jl:25-41: …, ρg, K, G, dt, igg) and 2D-V2 (Stokes2D.jl:181-196: …, ρg, G, K, dt, igg)"
function solve_arrays!(stokes, pt_stokes, grid, flow_bcs, ρg, K, G, dt, igg; kwargs...)
    kw = merge((; iterMax = 10.0e3, nout = 500, b_width = ND == 3 ? (4, 4, 4) : (4, 4, 1), verbose = true), values(kwargs))
    check_flow_bcs_type(flow_bcs)
    ensure_comm!(igg)
    ni = size(stokes.P)
    d = add_ρg!(stokes_slots(stokes), ρg)
    d["K"], d["G"] = LIB.ondevice(K), LIB.ondevice(G)
    f = API.Fields(ni, d)
    opt = Ref(stokes_opts(pt_stokes, grid, dt, flow_bcs, ni; kw...))
    h = API.History(kw.iterMax, kw.nout)
    fn = ND == 3 ? :jr_stokes3d_solve_VA : :jr_stokes2d_solve_V2
    GC.@preserve f h begin
        LIB.check(ccall(jrsym(fn), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{API.StokesOpts}, Ref{API.StokesResult}), ctx(), pointer(f), opt, h.res))
    end
    out = stokes_named_tuple(h)
    print_history(out, igg, kw.verbose)
    return out
end

"multiphase visco-elasto-plastic variants: 3D-VC (Stokes3D.jl:447-466) and 2D-VC (Stokes2D.jl:577-599)"
function solve_phases!(stokes, pt_stokes, grid, flow_bcs, ρg, phase_ratios, rheology::NTuple, args::NamedTuple, dt, igg; kwargs...)
    defaults = ND == 3 ?
        (; iterMax = 10.0e3, nout = 500, b_width = (4, 4, 4), verbose = true, viscosity_relaxation = 1.0e-2, λ_relaxation = 0.2,
         viscosity_cutoff = (-Inf, Inf), iterMin = 0) :
        (; iterMax = 50.0e3, iterMin = 1.0e2, viscosity_relaxation = 1.0e-2, λ_relaxation = 0.2, free_surface = false, nout = 500,
         b_width = (4, 4, 0), verbose = true, viscosity_cutoff = (-Inf, Inf), strain_increment = false)
    kw = merge(defaults, values(kwargs))
    check_flow_bcs_type(flow_bcs; allow_displacement = ND == 2)
    ensure_comm!(igg)
    ni = size(stokes.P)
    d = add_args!(add_ρg!(stokes_slots(stokes), ρg), args)
    f = API.Fields(ni, d)
    ghosted = haskey(args, :ΔT) && args.ΔT !== nothing && size(args.ΔT) == ni .+ 2     # thermal.ΔT, as the reference's scripts pass it
    (haskey(args, :ΔT) && args.ΔT !== nothing && !ghosted && size(args.ΔT) != ni) &&
        throw(ArgumentError("args.ΔT must have the size of the cell grid $(ni) or of thermal.ΔT $(ni .+ 2)"))
    opt = Ref(stokes_opts(pt_stokes, grid, dt, flow_bcs, ni; kw..., dT_ghosted = ghosted))
    fs = (ND == 2 && kw.free_surface !== false) ? Float64(dt) * Float64(kw.free_surface) : 0.0       # VelocityKernels.jl:134-180
    vc, keep = vc_inputs(rheology, phase_ratios; free_surface = fs)
    h = API.History(kw.iterMax, kw.nout)
    fn = ND == 3 ? :jr_stokes3d_solve_VC : :jr_stokes2d_solve_VC
    GC.@preserve f h keep begin
        LIB.check(ccall(jrsym(fn), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{API.StokesOpts}, Ref{API.VcInputs}, Ref{API.StokesResult}),
                        ctx(), pointer(f), opt, vc, h.res))
    end
    out = stokes_named_tuple(h)
    print_history(out, igg, kw.verbose)
    return out
end

# =====================================================================================================================
# per-time-step kernels either side of the loops
function flow_bcs_b200!(stokes, bcs::JustRelax.VelocityBoundaryConditions)              # 3D.jl:195-206 → BoundaryConditions.jl:65-100
    n3 = Int32[API.tuple3(size(stokes.P), 1)...]
    fs, ns, pe = Int32[API.flags6(bcs.free_slip)...], Int32[API.flags6(bcs.no_slip)...], Int32[API.flags6(bcs.periodic)...]
    if ND == 3
        LIB.check(ccall(jrsym(:jr_flow_bcs3d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                        ctx(), stokes.V.Vx.ptr, stokes.V.Vy.ptr, stokes.V.Vz.ptr, n3, fs, ns, pe))
    else
        LIB.check(ccall(jrsym(:jr_flow_bcs2d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                        ctx(), stokes.V.Vx.ptr, stokes.V.Vy.ptr, n3, fs, ns, pe))
    end
    return nothing
end
function flow_bcs_b200!(stokes, bcs::JustRelax.DisplacementBoundaryConditions)          # 3D.jl:209-218: the same kernels on U
    n3 = Int32[API.tuple3(size(stokes.P), 1)...]
    fs, ns, pe = Int32[API.flags6(bcs.free_slip)...], Int32[API.flags6(bcs.no_slip)...], Int32[API.flags6(bcs.periodic)...]
    if ND == 3
        LIB.check(ccall(jrsym(:jr_flow_bcs3d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                        ctx(), stokes.U.Ux.ptr, stokes.U.Uy.ptr, stokes.U.Uz.ptr, n3, fs, ns, pe))
    else
        LIB.check(ccall(jrsym(:jr_flow_bcs2d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                        ctx(), stokes.U.Ux.ptr, stokes.U.Uy.ptr, n3, fs, ns, pe))
    end
    return nothing
end

function scale_copy!(dst::B200Array, src::B200Array, factor)
    LIB.check(ccall(jrsym(:jr_scale_copy), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Csize_t), ctx(), dst.ptr, src.ptr, Float64(factor), length(src)))
end
function velocity2displacement_b200!(stokes, dt)                                        # 3D.jl:358-364 → types/displacement.jl:1-29
    scale_copy!(stokes.U.Ux, stokes.V.Vx, dt); scale_copy!(stokes.U.Uy, stokes.V.Vy, dt)
    ND == 3 && scale_copy!(stokes.U.Uz, stokes.V.Vz, dt)
    return nothing
end
function displacement2velocity_b200!(stokes, dt)                                        # 3D.jl:366-372 → types/displacement.jl:33-60
    scale_copy!(stokes.V.Vx, stokes.U.Ux, inv(dt)); scale_copy!(stokes.V.Vy, stokes.U.Uy, inv(dt))
    ND == 3 && scale_copy!(stokes.V.Vz, stokes.U.Uz, inv(dt))
    return nothing
end

"compute_viscosity!(stokes, phase_ratios, args, rheology, cutoff; relaxation)  3D.jl:231-263 → rheology/Viscosity.jl:67-106,282-323"
function compute_viscosity_b200!(stokes, ν, phase_ratios, args, rheology::NTuple, cutoff)
    ni = size(stokes.P)
    dummy = ND == 3 ? (stokes.P, stokes.P, stokes.P) : (stokes.P, stokes.P)
    f = API.Fields(ni, add_args!(add_ρg!(stokes_slots(stokes), dummy), args))
    opt = Ref(API.StokesOpts(0, 0, 0, 0, 0, (0.0, 0.0, 0.0), Inf, 0, 1, global_size(ni), ntuple(_ -> Int32(0), 6), ntuple(_ -> Int32(0), 6),
                             ntuple(_ -> Int32(0), 6), 0.0, 0.0, Float64(cutoff[1]), Float64(cutoff[2]), 0, Int32(0), Int32(0), Int32(0), Int32(0)))
    vc, keep = vc_inputs(rheology, phase_ratios)
    fn = ND == 3 ? :jr_compute_viscosity3d : :jr_compute_viscosity2d
    GC.@preserve f keep LIB.check(ccall(jrsym(fn), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{API.StokesOpts}, Ref{API.VcInputs}, Cdouble), ctx(), pointer(f), opt, vc, Float64(ν)))
    return nothing
end

"compute_ρg!(ρg, phase_ratios, rheology, args)  3D.jl:287-299 → rheology/BuoyancyForces.jl:38-60,74-95"
function compute_ρg_b200!(ρg::NTuple, phase_ratios, rheology::NTuple, args::NamedTuple)
    ni = size(ρg[1])
    d = Dict{String, Any}()
    add_args!(add_ρg!(d, ρg), args)
    f = API.Fields(ni, d)
    vc, keep = vc_inputs(rheology, phase_ratios)
    fn = ND == 3 ? :jr_compute_rhog3d : :jr_compute_rhog2d
    GC.@preserve f keep LIB.check(ccall(jrsym(fn), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{API.VcInputs}), ctx(), pointer(f), vc))
    return nothing
end

function tensor_invariant_b200!(A::JustRelax.SymmetricTensor)                           # 3D.jl:266-272 → StressKernels.jl:442-500
    n3 = Int32[API.tuple3(size(A.xx), 1)...]
    if ND == 3
        LIB.check(ccall(jrsym(:jr_tensor_invariant3d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                        ctx(), A.II.ptr, A.xx.ptr, A.yy.ptr, A.zz.ptr, A.yz.ptr, A.xz.ptr, A.xy.ptr, n3))
    else
        LIB.check(ccall(jrsym(:jr_tensor_invariant2d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                        ctx(), A.II.ptr, A.xx.ptr, A.yy.ptr, A.xy.ptr, n3))
    end
    return nothing
end

function accumulate_tensor_b200!(II::B200Array, A::JustRelax.SymmetricTensor, dt)       # 3D.jl:274-279 → StressKernels.jl:364-408
    n3 = Int32[API.tuple3(size(A.xx), 1)...]
    if ND == 3
        LIB.check(ccall(jrsym(:jr_accumulate_tensor3d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Cdouble),
                        ctx(), II.ptr, A.xx.ptr, A.yy.ptr, A.zz.ptr, A.yz.ptr, A.xz.ptr, A.xy.ptr, n3, Float64(dt)))
    else
        LIB.check(ccall(jrsym(:jr_accumulate_tensor2d), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Cdouble),
                        ctx(), II.ptr, A.xx.ptr, A.yy.ptr, A.xy.ptr, n3, Float64(dt)))
    end
    return nothing
end

accumulate_vol_b200!(EVol_pl::B200Array, ε_vol_pl::B200Array, dt) =                     # 3D.jl:281-284 → StressKernels.jl:422-438
    LIB.check(ccall(jrsym(:jr_accumulate_vol), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Csize_t, Cdouble), ctx(), EVol_pl.ptr, ε_vol_pl.ptr, length(EVol_pl), Float64(dt)))

"maximum(abs.(A)) (maximum_mpi with an initialised global grid): the reduction of compute_dt  3D.jl:388-390 → Utils.jl:492-519"
function absmax(A::B200Array)
    out = Ref{Float64}(0.0)
    allreduce = Cint(ImplicitGlobalGrid.grid_is_initialized() && COMM[] != C_NULL)
    LIB.check(ccall(jrsym(:jr_absmax), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t, Cint, Ref{Float64}), ctx(), A.ptr, length(A), allreduce, out))
    return out[]
end
function compute_dt_b200(S::JustRelax.StokesArrays, di, dt_diff = Inf)
    V = ND == 3 ? (S.V.Vx, S.V.Vy, S.V.Vz) : (S.V.Vx, S.V.Vy)
    dt_adv = mapreduce(x -> x[1] * inv(absmax(x[2])), min, zip(di, V)) * 0.9          # Utils.jl:507-519 (C = 0.9)
    return min(dt_diff, dt_adv)
end

# =====================================================================================================================
# thermal: heatdiffusion_PT!  (3D.jl:383-385 → thermal_diffusion/DiffusionPT_solver.jl:34-149 [K, ρCp] and :181-305 [rheology])
function thermal_opts(pt, grid, dt, bc::JustRelax.TemperatureBoundaryConditions, rows::Vector{API.ThermalPhase}, form; iterMax = 50.0e3, nout = 1.0e3, kw...)
    _require_uniform(grid)
    cva, cvv = API.valued6(bc.constant_value)
    cfa, cfv = API.valued6(bc.constant_flux)
    dirc = (hasproperty(bc.dirichlet, :values) && bc.dirichlet.values isa Number) ? Float64(bc.dirichlet.values) : 0.0
    return API.ThermalOpts(API.tuple3(_inv_spacing(grid), 0.0), Float64(dt), pt.ϵ, Int64(floor(iterMax)), Int64(floor(nout)), pt.max_lxyz, pt.Vpdτ,
                           Int32(form), Int32(length(rows)), isempty(rows) ? Ptr{API.ThermalPhase}(C_NULL) : pointer(rows), dirc,
                           API.flags6(bc.no_flux), cva, cfa, API.flags6(bc.periodic), cvv, cfv)
end

function thermal_fields(thermal::JustRelax.ThermalArrays, pt, keep::Vector{Any}; K = nothing, ρCp = nothing, P = nothing, phase = nothing,
                        bc::JustRelax.TemperatureBoundaryConditions)
    ni = size(thermal.H)
    dev(x) = x === nothing ? nothing : (y = LIB.ondevice(x); push!(keep, y); y)
    K, ρCp, P = dev(K), dev(ρCp), dev(P)
    mask = hasproperty(bc.dirichlet, :mask) ? bc.dirichlet.mask : nothing
    dval = (hasproperty(bc.dirichlet, :values) && bc.dirichlet.values isa AbstractArray) ? bc.dirichlet.values : nothing
    maskp = (mask === nothing || !hasproperty(mask, :mask)) ? nothing : dev(mask.mask)            # mask/mask.jl:5-22 (0/1 Float64)
    dvalp = dev(dval)
    pc = phase === nothing ? Ptr{Float64}(C_NULL) : flat(phase.center, keep)
    px = phase === nothing ? Ptr{Float64}(C_NULL) : flat(phase.Vx, keep)
    py = phase === nothing ? Ptr{Float64}(C_NULL) : flat(phase.Vy, keep)
    pz = (phase === nothing || ND == 2) ? Ptr{Float64}(C_NULL) : flat(phase.Vz, keep)
    p(x) = LIB.ptr_or_null(x)
    return API.ThermalFields(Int32(ND), map(Int32, API.tuple3(ni, 1)), p(thermal.T), p(thermal.Told), p(thermal.ΔT), p(thermal.qTx), p(thermal.qTy), p(thermal.qTz),
                             p(thermal.qTx2), p(thermal.qTy2), p(thermal.qTz2), p(thermal.H), p(thermal.shear_heating), p(thermal.adiabatic), p(thermal.ResT),
                             p(pt.θr_dτ), p(pt.dτ_ρ), p(K), p(ρCp), p(P), p(maskp), p(dvalp), pc, px, py, pz)
end

function thermal_call(thermal, f::API.ThermalFields, o::API.ThermalOpts, stokes, keep)
    cap = Int(floor(o.iterMax / max(o.nout, 1))) + 3
    norm_ResT, iter_count = zeros(cap), zeros(Int64, cap)
    res = Ref(API.ThermalResult(0, 0, cap, NaN, pointer(norm_ResT), pointer(iter_count), 0.0, 0))
    sP = stokes === nothing ? Ptr{Float64}(C_NULL) : stokes.P.ptr
    sP0 = stokes === nothing ? Ptr{Float64}(C_NULL) : stokes.P0.ptr
    GC.@preserve norm_ResT iter_count keep thermal begin
        LIB.check(ccall(jrsym(:jr_heatdiffusion_PT), Cint, (Ptr{Cvoid}, Ref{API.ThermalFields}, Ref{API.ThermalOpts}, Ptr{Float64}, Ptr{Float64}, Ref{API.ThermalResult}),
                        ctx(), Ref(f), Ref(o), sP, sP0, res))
    end
    n = Int(res[].nhist)
    return (iter_count = iter_count[1:n], norm_ResT = norm_ResT[1:n])                   # DiffusionPT_solver.jl:148, 304
end

"heatdiffusion_PT!(thermal, pt_thermal, thermal_bc, K, ρCp, dt, grid|di; kwargs)  DiffusionPT_solver.jl:34-48"
function heatdiffusion_arrays!(thermal, pt_thermal, thermal_bc, K::AbstractArray, ρCp::AbstractArray, dt, grid; kwargs...)
    kw = merge((; igg = nothing, b_width = (4, 4, 4), iterMax = 50.0e3, nout = 1.0e3, verbose = true), values(kwargs))
    kw.igg === nothing || ensure_comm!(kw.igg)
    keep = Any[]
    f = thermal_fields(thermal, pt_thermal, keep; K, ρCp, bc = thermal_bc)
    o = thermal_opts(pt_thermal, grid, dt, thermal_bc, API.ThermalPhase[], 0; kw...)
    return thermal_call(thermal, f, o, nothing, keep)
end

"heatdiffusion_PT!(thermal, pt_thermal, thermal_bc, rheology, args, dt, grid|di; kwargs = (; igg, phase, stokes, …))  :181-197"
function heatdiffusion_rheology!(thermal, pt_thermal, thermal_bc, rheology, args::NamedTuple, dt, grid; kwargs...)
    kw = merge((; igg = nothing, phase = nothing, stokes = nothing, b_width = (4, 4, 4), iterMax = 50.0e3, nout = 1.0e3, verbose = true), values(kwargs))
    kw.igg === nothing || ensure_comm!(kw.igg)
    rh = rheology isa Tuple ? rheology : (rheology,)
    rows = API.ThermalPhase[lower_thermal_phase(p) for p in rh]
    (length(rows) > 1 && kw.phase === nothing) && throw(ArgumentError("a multi-phase rheology needs kwargs.phase (PhaseRatios)"))
    keep = Any[rows]
    f = thermal_fields(thermal, pt_thermal, keep; P = args.P, phase = kw.phase, bc = thermal_bc)
    o = thermal_opts(pt_thermal, grid, dt, thermal_bc, rows, 1; kw...)
    return thermal_call(thermal, f, o, kw.stokes, keep)
end

function thermal_bcs_b200!(T::B200Array, bc::JustRelax.TemperatureBoundaryConditions)   # 3D.jl:220-226 → BoundaryConditions.jl:39-54
    ni = size(T) .- 2
    pt = (; ϵ = 0.0, max_lxyz = 0.0, Vpdτ = 0.0)
    o = thermal_opts(pt, ntuple(_ -> 1.0, ND), 0.0, bc, API.ThermalPhase[], 0)
    LIB.check(ccall(jrsym(:jr_thermal_bcs), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Int32}, Ref{API.ThermalOpts}), ctx(), T.ptr, Int32(ND), Int32[API.tuple3(ni, 1)...], Ref(o)))
    return nothing
end

"""
PTThermalCoeffs(B200Backend, …) — 3D.jl:75-108 → DiffusionPT_coefficients.jl:17-26, 53-65, 91-151.  The scalars are host
arithmetic (max_lxyz, Vpdτ = min(di)·CFL); the arrays θr_dτ, dτ_ρ are filled by the library (jr_thermal_pt_arrays).
"""
function pt_thermal_coeffs(K, ρCp, dt, di::NTuple, li::NTuple; ϵ = 1.0e-8, CFL = 0.9 / √(ND + 0.1), thermal = nothing, rows = API.ThermalPhase[], args = nothing, phase = nothing)
    ni = K === nothing ? size(args.P) : size(K)
    max_lxyz = max(li...)
    Vpdτ = min(di...) * CFL
    θr_dτ, dτ_ρ = z(ni...), z(ni...)
    pt = JustRelax.PTThermalCoeffs(CFL, ϵ, max_lxyz, max_lxyz^2, Vpdτ, θr_dτ, dτ_ρ)
    update_pt_arrays!(pt, dt, di; K, ρCp, rows, args, phase)
    return pt
end
function update_pt_arrays!(pt, dt, di; K = nothing, ρCp = nothing, rows = API.ThermalPhase[], args = nothing, phase = nothing)
    ni = size(pt.θr_dτ)
    keep = Any[rows]
    dev(x) = x === nothing ? nothing : (y = LIB.ondevice(x); push!(keep, y); y)
    K, ρCp = dev(K), dev(ρCp)
    T = args === nothing ? nothing : dev(args.T)
    P = args === nothing ? nothing : dev(args.P)
    p(x) = LIB.ptr_or_null(x)
    nul = Ptr{Float64}(C_NULL)
    pc = phase === nothing ? nul : flat(phase.center, keep)
    f = API.ThermalFields(Int32(ND), map(Int32, API.tuple3(ni, 1)), p(T), nul, nul, nul, nul, nul, nul, nul, nul, nul, nul, nul, nul, p(pt.θr_dτ), p(pt.dτ_ρ),
                          p(K), p(ρCp), p(P), nul, nul, pc, nul, nul, nul)
    no6, z6 = ntuple(_ -> Int32(0), 6), ntuple(_ -> 0.0, 6)
    o = API.ThermalOpts(API.tuple3(map(x -> 1.0 / x, di), 0.0), Float64(dt), pt.ϵ, 1, 1, pt.max_lxyz, pt.Vpdτ, Int32(isempty(rows) ? 0 : 1), Int32(length(rows)),
                        isempty(rows) ? Ptr{API.ThermalPhase}(C_NULL) : pointer(rows), 0.0, no6, no6, no6, no6, z6, z6)
    GC.@preserve keep LIB.check(ccall(jrsym(:jr_thermal_pt_arrays), Cint, (Ptr{Cvoid}, Ref{API.ThermalFields}, Ref{API.ThermalOpts}), ctx(), Ref(f), Ref(o)))
    return nothing
end

# =====================================================================================================================
# grid-based phase ratios: update_phase_ratios_{2,3}D!(phase_ratios, phase_arrays, xci, xvi)  3D.jl:519-539 → phases/PhaseRatios.jl:21-78
function update_phase_ratios_b200!(pr, phase_arrays::NTuple{NP, B200Array}, xci, xvi) where {NP}
    ni = size(phase_arrays[1])
    keep = Any[]
    ptrs = Ptr{Float64}[a.ptr for a in phase_arrays]
    xc = [collect(Float64, x) for x in xci]
    xv = [collect(Float64, x) for x in xvi]
    xcp, xvp = Ptr{Float64}[pointer(x) for x in xc], Ptr{Float64}[pointer(x) for x in xv]
    GC.@preserve xc xv ptrs keep begin
        LIB.check(ccall(jrsym(:jr_phase_ratios_from_arrays), Cint,
                        (Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                        ctx(), Int32(ND), Int32[API.tuple3(ni, 1)...], Int32(NP), ptrs, xcp, xvp, flat(pr.center, keep), flat(pr.vertex, keep), flat(pr.Vx, keep),
                        flat(pr.Vy, keep), flat(_maybe(pr, :Vz), keep), flat(_maybe(pr, :xy), keep), flat(_maybe(pr, :yz), keep), flat(_maybe(pr, :xz), keep)))
    end
    isempty(keep) || throw(ArgumentError("update_phase_ratios!: the PhaseRatios must live on the B200 (CellArrays over B200Array data)"))
    return nothing
end

# =====================================================================================================================
# multi-GPU: ImplicitGlobalGrid topology + CUDA-IPC peer memory inside the library (replaces update_halo! / MPI.Allreduce on this path)
const COMM = Ref{Ptr{Cvoid}}(C_NULL)

"the one host collective the library needs for bootstrap: an all-gather of small byte blobs (CUDA-IPC handles) on igg.comm_cart"
function _allgather_cb(send::Ptr{Cvoid}, recv::Ptr{Cvoid}, nbytes::Csize_t, user::Ptr{Cvoid})::Cint
    try
        comm = unsafe_pointer_to_objref(user)::MPI.Comm
        n = Int(nbytes)
        sbuf = unsafe_wrap(Array, Ptr{UInt8}(send), n)
        rbuf = unsafe_wrap(Array, Ptr{UInt8}(recv), n * MPI.Comm_size(comm))
        MPI.Allgather!(sbuf, MPI.UBuffer(rbuf, n), comm)
        return Cint(0)
    catch err
        @error "libjrb200 all-gather callback failed" err
        return Cint(1)
    end
end

"attach a libjrb200 communicator to the context: call once after `igg = IGG(init_global_grid(nx, ny, nz; …)...)`"
function init_b200_comm!(igg::JustRelax.IGG)
    igg.nprocs > 1 || return nothing
    cb = @cfunction(_allgather_cb, Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}))
    comm = igg.comm_cart
    out = Ref{Ptr{Cvoid}}(C_NULL)
    dims, coords = Int32[API.tuple3(igg.dims, 1)...], Int32[API.tuple3(igg.coords, 0)...]
    # periodx / periody / periodz of init_global_grid (ImplicitGlobalGrid keeps them in its global grid object)
    periods = Int32[ImplicitGlobalGrid.global_grid().periods...]
    GC.@preserve comm begin
        LIB.check(ccall(jrsym(:jr_comm_create_periodic), Cint,
                        (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                        ctx(), igg.me, igg.nprocs, dims, coords, periods, cb, pointer_from_objref(comm), out))
    end
    LIB.check(ccall(jrsym(:jr_context_set_comm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), out[]))
    COMM[] = out[]
    return nothing
end
"attach the communicator the first time a multi-rank IGG reaches a solver / halo update"
ensure_comm!(igg::JustRelax.IGG) =
    ((igg.nprocs > 1 || any(!iszero, ImplicitGlobalGrid.global_grid().periods)) && COMM[] == C_NULL) ? init_b200_comm!(igg) : nothing
function finalize_b200_comm!()
    COMM[] == C_NULL && return nothing
    LIB.check(ccall(jrsym(:jr_context_set_comm), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), C_NULL))
    LIB.check(ccall(jrsym(:jr_comm_destroy), Cint, (Ptr{Cvoid},), COMM[]))
    COMM[] = C_NULL
    return nothing
end

"update_halo!(A...) on B200 arrays (ImplicitGlobalGrid's call sites outside the solvers, e.g. the setup smoothing of SolVi3D.jl:38-42)"
function update_halo_b200!(A::Vararg{B200Array{3}, NA}) where {NA}
    g = ImplicitGlobalGrid.global_grid()
    if g.nprocs > 1 && COMM[] == C_NULL
        init_b200_comm!(JustRelax.IGG(g.me, collect(g.dims), g.nprocs, collect(g.coords), g.comm))
    end
    ptrs = Ptr{Float64}[a.ptr for a in A]
    ext = Int32[s for a in A for s in size(a)]
    LIB.check(ccall(jrsym(:jr_update_halo3d), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Float64}}, Ptr{Int32}, Ptr{Int32}), ctx(), NA, ptrs, ext, Int32[g.nxyz...]))
    return nothing
end
