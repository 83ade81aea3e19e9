"""
    JustRelaxB200

Trigger package of the `JustRelaxB200Ext` extension of JustRelax.jl: the role CUDA.jl plays for
`ext/JustRelaxCUDAExt.jl`.  It owns

  * the handle of `libjrb200.so` (C ABI: `include/jrb200.h`, ABI version 2) and one `jr_context` per process,
  * `B200Array{N} <: DenseArray{Float64,N}`: a dense column-major Float64 device array allocated with `jr_malloc`
    (freed by a finalizer), with `Array(::B200Array)`, `B200Array(::Array)`, `copyto!`, `fill!`, `copy`, `similar`,
  * bit-compatible mirrors of the POD structs of the header and raw `ccall` wrappers (`JustRelaxB200.API`),
  * the backend tag `B200Backend`.

No CUDA.jl, no ParallelStencil: every kernel lives in the shared library.  There is no CPU fallback — loading fails
when the library or a CUDA device is missing.
"""
module JustRelaxB200

using Libdl

export B200Backend, B200Array, b200zeros, b200ones, b200fill

struct B200Backend end

# ---------------------------------------------------------------------------------------------------------------
# library handle, error handling, context
const LIBPATH = Ref{String}("")
const LIBHANDLE = Ref{Ptr{Cvoid}}(C_NULL)
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
const ABI_VERSION = 3

const JR_OK = Cint(0)
const JR_ERR_CUDA = Cint(-1)
const JR_ERR_SHAPE = Cint(-2)
const JR_ERR_NAN = Cint(-3)
const JR_ERR_UNSUPPORTED = Cint(-4)
const JR_ERR_NCCL = Cint(-5)
const JR_ERR_ARG = Cint(-6)

function __init__()
    LIBPATH[] = get(ENV, "JRB200_LIB", "libjrb200.so")
    LIBHANDLE[] = Libdl.dlopen(LIBPATH[]; throw_error = true)    # no library ⇒ load error, never a fallback
    v = ccall(sym(:jr_abi_version), Cint, ())
    v == ABI_VERSION || error("libjrb200 ABI version $v, this package binds version $ABI_VERSION")
    return nothing
end

@inline sym(name::Symbol) = Libdl.dlsym(LIBHANDLE[], name)
last_error() = unsafe_string(ccall(sym(:jr_last_error), Cstring, ()))

"""
    check(status)

Map a `jr_status` to the exception the reference throws in the same situation (INTEGRATION.md §5):
`JR_ERR_NAN` → `ErrorException("NaN(s)")` (Stokes3D.jl:162, Stokes2D.jl:836); `JR_ERR_UNSUPPORTED`, `JR_ERR_SHAPE`,
`JR_ERR_ARG` → `ArgumentError`; CUDA errors → `ErrorException` with the CUDA string.
"""
function check(st::Integer)
    st == JR_OK && return nothing
    st == JR_ERR_NAN && error("NaN(s)")
    msg = last_error()
    (st == JR_ERR_UNSUPPORTED || st == JR_ERR_SHAPE || st == JR_ERR_ARG) && throw(ArgumentError(msg))
    error("libjrb200 status $st: $msg")
end

"""
    context(; device = parse(Int, get(ENV, "JRB200_DEVICE", "0")))

The process-wide `jr_context` (one GPU per process, like one MPI rank per GPU in the reference's multi-GPU runs).
"""
function context(; device::Integer = parse(Int, get(ENV, "JRB200_DEVICE", "0")))
    if CTX[] == C_NULL
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall(sym(:jr_context_create), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, out))
        CTX[] = out[]
    end
    return CTX[]
end

synchronize() = check(ccall(sym(:jr_context_synchronize), Cint, (Ptr{Cvoid},), context()))

# ---------------------------------------------------------------------------------------------------------------
# B200Array
"""
    B200Array{N}

Dense column-major `Float64` array in the HBM of the context's GPU.  Element access from the host is deliberately
slow-path only (`getindex` copies one element; used by `show` and tests) — bulk data moves with `Array(a)`,
`copyto!` and the library's kernels.
"""
mutable struct B200Array{N} <: DenseArray{Float64, N}
    ptr::Ptr{Float64}
    dims::NTuple{N, Int}
    function B200Array{N}(::UndefInitializer, dims::NTuple{N, Int}) where {N}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall(sym(:jr_malloc), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), context(), 8 * max(prod(dims), 1), p))
        a = new{N}(Ptr{Float64}(p[]), dims)
        finalizer(a) do x
            x.ptr == C_NULL || ccall(sym(:jr_free), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], x.ptr)
            x.ptr = C_NULL
        end
        return a
    end
end
B200Array{N}(u::UndefInitializer, dims::Vararg{Integer, N}) where {N} = B200Array{N}(u, map(Int, dims))
B200Array(u::UndefInitializer, dims::NTuple{N, Integer}) where {N} = B200Array{N}(u, map(Int, dims))
B200Array(u::UndefInitializer, dims::Vararg{Integer, N}) where {N} = B200Array{N}(u, map(Int, dims))

Base.size(a::B200Array) = a.dims
Base.length(a::B200Array) = prod(a.dims)
Base.sizeof(a::B200Array) = 8 * length(a)
Base.pointer(a::B200Array) = a.ptr
Base.unsafe_convert(::Type{Ptr{Float64}}, a::B200Array) = a.ptr
Base.unsafe_convert(::Type{Ptr{Cvoid}}, a::B200Array) = Ptr{Cvoid}(a.ptr)
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
Base.elsize(::Type{<:B200Array}) = 8
Base.strides(a::B200Array{N}) where {N} = ntuple(d -> d == 1 ? 1 : prod(a.dims[1:(d - 1)]), Val(N))

function Base.fill!(a::B200Array, v::Real)
    check(ccall(sym(:jr_fill_f64), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cdouble, Csize_t), context(), a.ptr, Float64(v), length(a)))
    return a
end
b200fill(v::Real, dims::Vararg{Integer, N}) where {N} = fill!(B200Array{N}(undef, map(Int, dims)), v)
b200fill(v::Real, dims::NTuple{N, Integer}) where {N} = b200fill(v, dims...)
b200zeros(dims...) = b200fill(0.0, dims...)
b200ones(dims...) = b200fill(1.0, dims...)

# host → device, device → host, device → device
function Base.copyto!(dst::B200Array, src::Array{Float64})
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) ← $(size(src))"))
    GC.@preserve src check(ccall(sym(:jr_memcpy_h2d), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), context(), dst.ptr, pointer(src), sizeof(src)))
    return dst
end
Base.copyto!(dst::B200Array, src::AbstractArray{<:Real}) = copyto!(dst, convert(Array{Float64}, collect(src)))
function Base.copyto!(dst::Array{Float64}, src::B200Array)
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) ← $(size(src))"))
    GC.@preserve dst check(ccall(sym(:jr_memcpy_d2h), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), context(), pointer(dst), src.ptr, sizeof(dst)))
    return dst
end
function Base.copyto!(dst::B200Array, src::B200Array)
    length(dst) == length(src) || throw(DimensionMismatch("copyto!: $(size(dst)) ← $(size(src))"))
    check(ccall(sym(:jr_memcpy_d2d), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), context(), dst.ptr, src.ptr, sizeof(dst)))
    return dst
end
B200Array(h::AbstractArray{<:Real, N}) where {N} = copyto!(B200Array{N}(undef, size(h)), h)
B200Array(a::B200Array) = a
B200Array{N}(h::AbstractArray{<:Real, N}) where {N} = B200Array(h)
Base.Array(a::B200Array{N}) where {N} = copyto!(Array{Float64, N}(undef, a.dims), a)
Base.collect(a::B200Array) = Array(a)
Base.copy(a::B200Array{N}) where {N} = copyto!(B200Array{N}(undef, a.dims), a)
Base.similar(a::B200Array{N}) where {N} = B200Array{N}(undef, a.dims)
Base.similar(::B200Array, ::Type{Float64}, dims::Dims{N}) where {N} = B200Array{N}(undef, dims)
Base.zero(a::B200Array) = fill!(similar(a), 0.0)

# scalar access (debug / show / tests only): one element over PCIe
function Base.getindex(a::B200Array, i::Int)
    @boundscheck checkbounds(a, i)
    h = Ref{Float64}(0.0)
    check(ccall(sym(:jr_memcpy_d2h), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), context(), h, a.ptr + 8 * (i - 1), 8))
    return h[]
end
function Base.setindex!(a::B200Array, v, i::Int)
    @boundscheck checkbounds(a, i)
    h = Ref{Float64}(Float64(v))
    check(ccall(sym(:jr_memcpy_h2d), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), context(), a.ptr + 8 * (i - 1), h, 8))
    return a
end
# the one broadcast the solvers' callers rely on: `A .= scalar` / `A .= B` (e.g. `stokes.P .= θ`, `@copy`)
Base.Broadcast.materialize!(dst::B200Array, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{Real}}) = fill!(dst, bc.args[1])
Base.Broadcast.materialize!(dst::B200Array, bc::Base.Broadcast.Broadcasted{<:Any, <:Any, typeof(identity), <:Tuple{B200Array}}) = copyto!(dst, bc.args[1])

"pointer of an optional array (`nothing` → NULL)"
@inline ptr_or_null(::Nothing) = Ptr{Float64}(C_NULL)
@inline ptr_or_null(a::B200Array) = a.ptr

"""
    ondevice(x)

`x` itself when it is a `B200Array`; a device copy when it is a host `Array` (ρg, `args.T`, K, G created with the
ParallelStencil-Threads `@zeros/@fill` of a user script, e.g. test/test_shearband2D.jl:129-130, are uploaded at solve entry).
"""
ondevice(x::B200Array) = x
ondevice(x::AbstractArray{<:Real}) = B200Array(x)
ondevice(::Nothing) = nothing

include("api.jl")

end # module
