"""justrelax_jl_b200 — B200 (sm_100a) backend for the pseudo-transient hot path of JustRelax.jl.

Host-side mirror of the reference's public API for that path (solve!, heatdiffusion_PT!, StokesArrays,
ThermalArrays, PTStokesCoeffs, boundary-condition and grid types) over the C ABI of libjrb200.so
(include/jrb200.h).  PyTorch is used for device memory / streams / process-group plumbing only.
"""
from .types import (AbstractBackend, CPUBackend, B200Backend, CPUBackendTrait, B200BackendTrait, PTArray, backend, zeros,
                    to_host, StokesArrays, ThermalArrays, PTStokesCoeffs, VelocityBoundaryConditions,
                    DisplacementBoundaryConditions, TemperatureBoundaryConditions, Geometry, IGG, legacy_uniform_grid, PhaseRatios)

__all__ = [n for n in dir() if not n.startswith("_")]
