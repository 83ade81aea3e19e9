# api.jl — bit-compatible mirrors of the POD structs of include/jrb200.h and the marshalling helpers the extension uses.
# Every struct below is `isbits` and laid out by Julia with C alignment rules, so `Ref(x)` can be passed where the C
# function takes a pointer to the struct.  Sizes are asserted against the header's layout in `__init__`-time tests
# (`JustRelaxB200.API.selfcheck()`), the Python twin of which is tests/test_abi.py::test_struct_sizes_match_header.
module API

using ..JustRelaxB200: JustRelaxB200, B200Array, sym, check, context, ptr_or_null, ondevice

# ---- jr_fields: {int32 ndim; int32 n[3]; double *f[JR_F_COUNT]} -----------------------------------------------------
# The slot table is read from the library at run time (jr_field_count / jr_field_name), so this file never hard-codes
# an index; the struct is assembled as a byte buffer.
const FIELD_NAMES = String[]
function field_names()
    if isempty(FIELD_NAMES)
        n = ccall(sym(:jr_field_count), Cint, ())
        for i in 0:(n - 1)
            push!(FIELD_NAMES, unsafe_string(ccall(sym(:jr_field_name), Cstring, (Cint,), i)))
        end
    end
    return FIELD_NAMES
end

"""
    Fields(ni, slots::Dict{String,<:Any})

A `jr_fields` value: `slots` maps slot names of `JR_STOKES_FIELDS` to `B200Array`s (missing / `nothing` → NULL).
Keeps the arrays alive; `pointer(f)` is what the C functions take.
"""
struct Fields
    buf::Vector{UInt8}
    keep::Vector{Any}
end
function Fields(ni::NTuple{N, Integer}, slots::AbstractDict) where {N}
    names = field_names()
    unknown = setdiff(keys(slots), names)
    isempty(unknown) || throw(ArgumentError("unknown field slots $(collect(unknown))"))
    buf = zeros(UInt8, 16 + 8 * length(names))
    GC.@preserve buf begin
        p32 = Ptr{Int32}(pointer(buf))
        unsafe_store!(p32, Int32(N), 1)
        for d in 1:3
            unsafe_store!(p32, Int32(d <= N ? ni[d] : 1), 1 + d)
        end
        pp = Ptr{Ptr{Float64}}(pointer(buf) + 16)
        for (i, nm) in enumerate(names)
            a = get(slots, nm, nothing)
            unsafe_store!(pp, a === nothing ? Ptr{Float64}(C_NULL) : (a::B200Array).ptr, i)
        end
    end
    return Fields(buf, Any[v for v in values(slots)])
end
Base.pointer(f::Fields) = Ptr{Cvoid}(pointer(f.buf))

# ---- jr_stokes_opts ----------------------------------------------------------------------------------------------
struct StokesOpts
    r::Cdouble
    theta_dtau::Cdouble
    eta_dtau::Cdouble
    eps_rel::Cdouble
    eps_abs::Cdouble
    _di::NTuple{3, Cdouble}
    dt::Cdouble
    iterMax::Int64
    nout::Int64
    n_g::NTuple{3, Int32}
    free_slip::NTuple{6, Int32}      # left, right, front, back, top, bot
    no_slip::NTuple{6, Int32}
    periodic::NTuple{6, Int32}
    viscosity_relaxation::Cdouble
    lambda_relaxation::Cdouble
    visc_cutoff_lo::Cdouble
    visc_cutoff_hi::Cdouble
    iterMin::Int64
    strain_rate_ni_only::Int32
    strain_increment::Int32          # 2D-VC kwarg strain_increment (Δε form)
    displacement_bcs::Int32          # flow_bcs isa DisplacementBoundaryConditions
    dT_ghosted::Int32                # args.ΔT has extents ni.+2 (thermal.ΔT) and is indexed ΔT[I...] without offset, like the reference
end

# ---- jr_stokes_result (history arrays are HOST pointers) --------------------------------------------------------------
struct StokesResult
    iter::Int64
    nhist::Int64
    err::Cdouble
    err_evo1::Ptr{Cdouble}
    err_evo2::Ptr{Int64}
    norm_Rx::Ptr{Cdouble}
    norm_Ry::Ptr{Cdouble}
    norm_Rz::Ptr{Cdouble}
    norm_divV::Ptr{Cdouble}
    time_s::Cdouble
    kernel_launches::Int64
end

"history vectors of the NamedTuple solve! returns (Stokes3D.jl:64-69) + the Ref the C call fills"
mutable struct History
    err_evo1::Vector{Float64}
    err_evo2::Vector{Int64}
    norm_Rx::Vector{Float64}
    norm_Ry::Vector{Float64}
    norm_Rz::Vector{Float64}
    norm_divV::Vector{Float64}
    res::Base.RefValue{StokesResult}
end
function History(iterMax, nout)
    cap = Int(floor(iterMax / max(nout, 1))) + 3
    h = History(zeros(cap), zeros(Int64, cap), zeros(cap), zeros(cap), zeros(cap), zeros(cap), Ref{StokesResult}())
    h.res[] = StokesResult(0, 0, NaN, pointer(h.err_evo1), pointer(h.err_evo2), pointer(h.norm_Rx), pointer(h.norm_Ry),
                           pointer(h.norm_Rz), pointer(h.norm_divV), 0.0, 0)
    return h
end

# ---- jr_stokes_phase / jr_vc_inputs --------------------------------------------------------------------------------
struct StokesPhase
    eta::Cdouble
    G::Cdouble
    Kb::Cdouble
    has_pl::Int32
    rho_kind::Int32      # 0 ConstantDensity, 1 PT_Density, 2 T_Density
    C::Cdouble
    sinphi::Cdouble
    cosphi::Cdouble
    sinpsi::Cdouble
    eta_vp::Cdouble
    rho0::Cdouble
    alpha::Cdouble
    beta::Cdouble
    T0::Cdouble
    P0::Cdouble
    soft_C_kind::Int32               # 0 none, 1 LinearSoftening, 2 NonLinearSoftening (cohesion, 2D solves)
    _pad::Int32
    soft_C::NTuple{6, Cdouble}
end

struct VcInputs
    nphase::Int32
    g_scalar::Int32
    phases::Ptr{StokesPhase}         # HOST pointer
    g::NTuple{3, Cdouble}
    ph_center::Ptr{Cdouble}
    ph_vertex::Ptr{Cdouble}
    ph_xy::Ptr{Cdouble}
    ph_yz::Ptr{Cdouble}
    ph_xz::Ptr{Cdouble}
    free_surface::Cdouble
end

# ---- thermal ------------------------------------------------------------------------------------------------------
struct ThermalFields
    ndim::Int32
    n::NTuple{3, Int32}
    T::Ptr{Cdouble}
    Told::Ptr{Cdouble}
    dT::Ptr{Cdouble}
    qTx::Ptr{Cdouble}
    qTy::Ptr{Cdouble}
    qTz::Ptr{Cdouble}
    qTx2::Ptr{Cdouble}
    qTy2::Ptr{Cdouble}
    qTz2::Ptr{Cdouble}
    H::Ptr{Cdouble}
    shear_heating::Ptr{Cdouble}
    adiabatic::Ptr{Cdouble}
    ResT::Ptr{Cdouble}
    theta_r_dtau::Ptr{Cdouble}
    dtau_rho::Ptr{Cdouble}
    K::Ptr{Cdouble}
    rhoCp::Ptr{Cdouble}
    P::Ptr{Cdouble}
    dir_mask::Ptr{Cdouble}
    dir_value::Ptr{Cdouble}
    phase_c::Ptr{Cdouble}
    phase_x::Ptr{Cdouble}
    phase_y::Ptr{Cdouble}
    phase_z::Ptr{Cdouble}
end

struct ThermalPhase
    rho_kind::Int32
    has_Hr::Int32
    rho0::Cdouble
    alpha::Cdouble
    beta::Cdouble
    T0::Cdouble
    P0::Cdouble
    Cp::Cdouble
    k::Cdouble
    Hr::Cdouble
    k_kind::Int32                    # 0 ConstantConductivity (k), 1 TP_Conductivity: (k_a + k_b / (T + k_c)) (1 + k_d P)
    _pad::Int32
    k_a::Cdouble
    k_b::Cdouble
    k_c::Cdouble
    k_d::Cdouble
end

struct ThermalOpts
    _di::NTuple{3, Cdouble}
    dt::Cdouble
    eps::Cdouble
    iterMax::Int64
    nout::Int64
    max_lxyz::Cdouble
    Vpdtau::Cdouble
    form::Int32                      # 0: K, ρCp arrays; 1: rheology table
    nphase::Int32
    phases::Ptr{ThermalPhase}        # HOST pointer
    dir_const::Cdouble
    no_flux::NTuple{6, Int32}
    cv_active::NTuple{6, Int32}
    cf_active::NTuple{6, Int32}
    periodic::NTuple{6, Int32}
    cv_value::NTuple{6, Cdouble}
    cf_value::NTuple{6, Cdouble}
end

struct ThermalResult
    iter::Int64
    nhist::Int64
    cap::Int64
    err::Cdouble
    norm_ResT::Ptr{Cdouble}
    iter_count::Ptr{Int64}
    time_s::Cdouble
    kernel_launches::Int64
end

"layout check against the sizes the C compiler produces for include/jrb200.h (same numbers as tests/test_abi.py)"
function selfcheck()
    @assert sizeof(StokesOpts) == 5 * 8 + 3 * 8 + 8 + 16 + 12 + 3 * 24 + 4 + 4 * 8 + 8 + 16
    @assert sizeof(StokesResult) == 11 * 8
    @assert sizeof(StokesPhase) == 21 * 8                  # 13 + 6 doubles + 2 × two int32 sharing one 8-byte slot
    @assert sizeof(VcInputs) == 8 + 8 + 24 + 5 * 8 + 8
    @assert sizeof(ThermalFields) == 16 + 24 * 8
    @assert sizeof(ThermalPhase) == 8 + 8 * 8 + 8 + 4 * 8
    @assert sizeof(ThermalOpts) == 24 + 16 + 16 + 16 + 8 + 8 + 8 + 4 * 24 + 2 * 48
    @assert sizeof(ThermalResult) == 8 * 8
    return true
end

# ---- boundary-condition flags: NamedTuple(left, right, front, back, top, bot) → 6 × Int32 ----------------------------
const FACES = (:left, :right, :front, :back, :top, :bot)
"flags of a Velocity/DisplacementBoundaryConditions field (src/boundaryconditions/types.jl:139-157); 2D tuples lack front/back"
flags6(nt::NamedTuple) = ntuple(q -> Int32(haskey(nt, FACES[q]) && nt[FACES[q]] === true), Val(6))
"constant_value / constant_flux entries are `false` or a number (types.jl:65-99): (active flags, values)"
function valued6(nt::NamedTuple)
    act = ntuple(q -> Int32(haskey(nt, FACES[q]) && !(nt[FACES[q]] === false)), Val(6))
    val = ntuple(q -> (haskey(nt, FACES[q]) && !(nt[FACES[q]] === false)) ? Float64(nt[FACES[q]]) : 0.0, Val(6))
    return act, val
end

tuple3(x, fillv) = ntuple(d -> d <= length(x) ? x[d] : fillv, Val(3))

end # module API
