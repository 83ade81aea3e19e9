# ext/JustRelaxB200Ext.jl — the B200 backend of JustRelax.jl as a package extension.
#
# Replaces ext/JustRelaxCUDAExt.jl:1-15 (+ src/ext/CUDA/{2D,3D}.jl) for the pseudo-transient hot path: Julia host code
# drives libjrb200.so (C ABI include/jrb200.h) through ccall; no CUDA.jl, no ParallelStencil kernel, no CPU fallback.
# Loaded when the trigger package JustRelaxB200 is imported next to JustRelax:
#
#   Project.toml of JustRelax.jl gains
#       [weakdeps]    JustRelaxB200 = "6d2f3c1e-52a7-4b0b-9c1d-b200b200b200"
#       [extensions]  JustRelaxB200Ext = "JustRelaxB200"
#   and src/types/traits.jl:2-7 gains nothing: the trait below subtypes the existing abstract GPUBackendTrait.
#
# User scripts change exactly like they do for CUDA: `const backend = B200Backend` instead of `CUDABackend`.
module JustRelaxB200Ext

using JustRelaxB200
using JustRelaxB200: B200Array, B200Backend
using JustRelax: JustRelax
import JustRelax: PTArray, backend, GPUBackendTrait

struct B200BackendTrait <: GPUBackendTrait end                    # src/types/traits.jl:2

PTArray(::Type{B200Backend}) = B200Array                           # ext/JustRelaxCUDAExt.jl:7

@inline backend(::B200Array) = B200BackendTrait()                  # :9
@inline backend(::Type{<:B200Array}) = B200BackendTrait()          # :10

include("../src/ext/B200/2D.jl")                                   # :12
include("../src/ext/B200/3D.jl")                                   # :13

end
