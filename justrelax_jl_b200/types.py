"""Host-side mirror of the JustRelax.jl data types that cross the drop-in boundary.

Reference (all paths relative to /root/reference):
  StokesArrays / Velocity / SymmetricTensor / Residual / Viscosity ...  src/types/stokes.jl:1-183,
                                                                        src/types/constructors/stokes.jl:10-302
  PTStokesCoeffs                                                        src/types/stokes.jl:203-229
  ThermalArrays / PTThermalCoeffs                                       src/types/heat_diffusion.jl, constructors/heat_diffusion.jl
  Velocity/Displacement/TemperatureBoundaryConditions                   src/boundaryconditions/types.jl:65-203
  Geometry / IGG                                                        src/grid/Cartesian.jl:9-58, src/grid/Grid.jl:18-24
  backends / PTArray / backend()                                        src/JustRelax.jl:170-178, src/types/traits.jl:1-33,
                                                                        ext/JustRelaxCUDAExt.jl:7-13

Arrays keep Julia's memory layout: dense column-major Float64.  Host arrays are numpy
(order='F'); device arrays are torch CUDA tensors whose strides are Fortran ordered, so
`A[i, j, k]` indexes exactly like the Julia array (0-based) and `A.data_ptr()` is what the
C ABI receives.  Field names follow the reference (τ, ε, η, ητ, λ ...); `∇V`/`∇U` are not legal
Python identifiers and are spelled `divV`/`divU`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np


# ----------------------------------------------------------------------------- backends
class AbstractBackend:
    pass


class CPUBackend(AbstractBackend):
    """Host container only (numpy).  The solvers of this package refuse to run on it:
    the CPU implementation is JustRelax.jl's own ParallelStencil-Threads backend."""


class B200Backend(AbstractBackend):
    """Device arrays on an NVIDIA B200 driven by libjrb200 (replaces CUDABackend)."""


class CPUBackendTrait:
    pass


class B200BackendTrait:
    pass


def _torch():
    import torch

    return torch


def zeros(backend, *shape, fill: float = 0.0, device=None):
    """`@zeros/@ones/@fill` equivalent producing a column-major Float64 array."""
    shape = tuple(int(s) for s in (shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape))
    if backend is CPUBackend:
        a = np.empty(shape, dtype=np.float64, order="F")
        a[...] = fill
        return a
    if backend is B200Backend:
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("B200Backend needs a CUDA device; justrelax_jl_b200 has no CPU fallback")
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        t = torch.full(tuple(reversed(shape)), float(fill), dtype=torch.float64, device=dev)
        return t.permute(*reversed(range(len(shape))))
    raise ValueError(f"Backend {backend} not supported")  # ArgumentError in types/traits.jl:33


def PTArray(backend, x=None):
    """PTArray(backend)(x): move/convert an array to the backend's array type
    (src/types/type_conversions.jl:50-63)."""
    if x is None:
        return lambda y: PTArray(backend, y)
    if backend is CPUBackend:
        return to_host(x)
    if backend is B200Backend:
        torch = _torch()
        h = np.asfortranarray(to_host(x), dtype=np.float64)
        t = zeros(B200Backend, *h.shape)
        # contiguous view of the same memory in reversed-dims order
        t.permute(*reversed(range(h.ndim))).copy_(torch.from_numpy(np.ascontiguousarray(h.T)))
        return t
    raise ValueError(f"Backend {backend} not supported")


def to_host(x) -> np.ndarray:
    """Array(x): column-major numpy copy/view of any backend array."""
    if isinstance(x, np.ndarray):
        return x
    torch = _torch()
    if isinstance(x, torch.Tensor):
        nd = x.dim()
        c = x.permute(*reversed(range(nd))).contiguous().cpu().numpy()  # C-order of reversed dims
        return c.T  # F-order view with the original shape
    return np.asfortranarray(np.asarray(x, dtype=np.float64))


def is_device_array(x) -> bool:
    if isinstance(x, np.ndarray):
        return False
    torch = _torch()
    return isinstance(x, torch.Tensor) and x.is_cuda


def backend(x):
    """backend(x) trait dispatch (src/types/traits.jl:11-33)."""
    if isinstance(x, (StokesArrays,)):
        return backend(x.P)
    if isinstance(x, ThermalArrays):
        return backend(x.T)
    if isinstance(x, np.ndarray):
        return CPUBackendTrait()
    if is_device_array(x):
        return B200BackendTrait()
    raise ValueError(f"Backend {type(x)} not supported")


def data_ptr(x) -> int:
    """Raw address of a column-major Float64 array (device pointer for B200 arrays)."""
    if isinstance(x, np.ndarray):
        if not x.flags.f_contiguous or x.dtype != np.float64:
            raise ValueError("host arrays must be Float64 column-major")
        return x.ctypes.data
    torch = _torch()
    if isinstance(x, torch.Tensor):
        nd = x.dim()
        if x.dtype != torch.float64 or not x.permute(*reversed(range(nd))).is_contiguous():
            raise ValueError("device arrays must be Float64 column-major")
        return x.data_ptr()
    raise TypeError(type(x))


# ----------------------------------------------------------------------------- Stokes containers
class _NS(SimpleNamespace):
    def __iter__(self):
        return iter(self.__dict__.values())


def Velocity(backend_t, *ni):
    if len(ni) == 2:
        nx, ny = ni
        return _NS(Vx=zeros(backend_t, nx + 1, ny + 2), Vy=zeros(backend_t, nx + 2, ny + 1), Vz=None)
    nx, ny, nz = ni
    return _NS(Vx=zeros(backend_t, nx + 1, ny + 2, nz + 2), Vy=zeros(backend_t, nx + 2, ny + 1, nz + 2),
               Vz=zeros(backend_t, nx + 2, ny + 2, nz + 1))


def Displacement(backend_t, *ni):
    v = Velocity(backend_t, *ni)
    return _NS(Ux=v.Vx, Uy=v.Vy, Uz=v.Vz)


def Vorticity(backend_t, *ni):
    if len(ni) == 2:
        nx, ny = ni
        return _NS(yz=None, xz=None, xy=zeros(backend_t, nx + 1, ny + 1))
    nx, ny, nz = ni
    return _NS(yz=zeros(backend_t, nx, ny + 1, nz + 1), xz=zeros(backend_t, nx + 1, ny, nz + 1),
               xy=zeros(backend_t, nx + 1, ny + 1, nz))


def Viscosity(backend_t, ni):
    ni1 = tuple(n + 1 for n in ni)
    return _NS(η=zeros(backend_t, *ni, fill=1.0), ηv=zeros(backend_t, *ni1, fill=1.0),
               η_vep=zeros(backend_t, *ni, fill=1.0), ητ=zeros(backend_t, *ni))


def SymmetricTensor(backend_t, *ni, vertex_normals: bool = True):
    """SymmetricTensor (constructors/stokes.jl:139-198).  `vertex_normals=False` skips the three
    (n+1)^3 `xx_v,yy_v,zz_v` arrays in 3D, which no solver on the hot path touches."""
    z = lambda *s: zeros(backend_t, *s)
    if len(ni) == 2:
        nx, ny = ni
        return _NS(xx=z(nx, ny), yy=z(nx, ny), xx_v=z(nx + 1, ny + 1), yy_v=z(nx + 1, ny + 1),
                   xy=z(nx + 1, ny + 1), xy_c=z(nx, ny), II=z(nx, ny))
    nx, ny, nz = ni
    v = (lambda: z(nx + 1, ny + 1, nz + 1)) if vertex_normals else (lambda: None)
    return _NS(xx=z(nx, ny, nz), yy=z(nx, ny, nz), zz=z(nx, ny, nz), xx_v=v(), yy_v=v(), zz_v=v(),
               xy=z(nx + 1, ny + 1, nz), yz=z(nx, ny + 1, nz + 1), xz=z(nx + 1, ny, nz + 1),
               yz_c=z(nx, ny, nz), xz_c=z(nx, ny, nz), xy_c=z(nx, ny, nz), II=z(nx, ny, nz))


def Residual(backend_t, *ni):
    z = lambda *s: zeros(backend_t, *s)
    if len(ni) == 2:
        nx, ny = ni
        return _NS(RP=z(nx, ny), Rx=z(nx - 1, ny), Ry=z(nx, ny - 1), Rz=None)
    nx, ny, nz = ni
    return _NS(RP=z(nx, ny, nz), Rx=z(nx - 1, ny, nz), Ry=z(nx, ny - 1, nz), Rz=z(nx, ny, nz - 1))


class StokesArrays:
    """StokesArrays(backend, ni) — src/types/stokes.jl:161-197, constructors/stokes.jl:277-302."""

    def __init__(self, backend_t, *ni, vertex_normals: bool = True):
        if len(ni) == 1 and isinstance(ni[0], (tuple, list)):
            ni = tuple(ni[0])
        for n in ni:
            if not isinstance(n, (int, np.integer)):
                raise TypeError("StokesArrays dimensions must be given as integers")  # types/stokes.jl:192-197
        if len(ni) not in (2, 3):
            raise ValueError("StokesArrays needs 2 or 3 dimensions")
        ni = tuple(int(n) for n in ni)
        z = lambda *s: zeros(backend_t, *s)
        self.backend_type = backend_t
        self.ni = ni
        self.P = z(*ni)
        self.P0 = z(*ni)
        self.V = Velocity(backend_t, *ni)
        self.divV = z(*ni)
        self.Q = z(*ni)
        self.τ = SymmetricTensor(backend_t, *ni, vertex_normals=vertex_normals)
        self.ε = SymmetricTensor(backend_t, *ni, vertex_normals=vertex_normals)
        self.ε_pl = SymmetricTensor(backend_t, *ni, vertex_normals=vertex_normals)
        self.EII_pl = z(*ni)
        self.EVol_pl = z(*ni)
        self.ε_vol_pl = z(*ni)
        self.viscosity = Viscosity(backend_t, ni)
        self.τ_o = SymmetricTensor(backend_t, *ni, vertex_normals=vertex_normals)
        self.R = Residual(backend_t, *ni)
        self.U = Displacement(backend_t, *ni)
        self.ω = Vorticity(backend_t, *ni)
        self.Δε = SymmetricTensor(backend_t, *ni, vertex_normals=vertex_normals)
        self.divU = z(*ni)
        self.λ = z(*ni)
        self.λv = z(*(n + 1 for n in ni))
        self.ΔPψ = z(*ni)

    # ABI slot name -> array (include/jrb200.h JR_STOKES_FIELDS)
    def slots(self) -> dict:
        d = dict(P=self.P, P0=self.P0, divV=self.divV, Q=self.Q,
                 Vx=self.V.Vx, Vy=self.V.Vy, Vz=self.V.Vz, Ux=self.U.Ux, Uy=self.U.Uy, Uz=self.U.Uz,
                 EII_pl=self.EII_pl, EVol_pl=self.EVol_pl, e_vol_pl=self.ε_vol_pl,
                 eta=self.viscosity.η, etav=self.viscosity.ηv, eta_vep=self.viscosity.η_vep, etatau=self.viscosity.ητ,
                 Rx=self.R.Rx, Ry=self.R.Ry, Rz=self.R.Rz, RP=self.R.RP,
                 wyz=self.ω.yz, wxz=self.ω.xz, wxy=self.ω.xy,
                 divU=self.divU, lam=self.λ, lamv=self.λv, dPpsi=self.ΔPψ)
        for pre, T, suf in (("t", self.τ, ""), ("t", self.τ_o, "_o"), ("e", self.ε, ""), ("p", self.ε_pl, ""), ("d", self.Δε, "")):
            for comp in ("xx", "yy", "zz", "yz", "xz", "xy"):
                d[f"{pre}{comp}{suf}"] = getattr(T, comp, None)
            for comp in ("yz", "xz", "xy"):
                d[f"{pre}{comp}{suf}_c"] = getattr(T, comp + "_c", None)
            d[f"{pre}II{suf}"] = T.II
        if len(self.ni) == 2:
            d["txx_v"], d["tyy_v"] = self.τ.xx_v, self.τ.yy_v
            d["txx_o_v"], d["tyy_o_v"] = self.τ_o.xx_v, self.τ_o.yy_v
        return d

    def to_host(self) -> dict:
        return {k: (None if v is None else np.array(to_host(v), order="F", copy=True)) for k, v in self.slots().items()}


@dataclass
class PTStokesCoeffs:
    """PTStokesCoeffs(li, di; ϵ_rel=1e-6, ϵ_abs=1e-12, Re=3π, CFL, r=0.7) — src/types/stokes.jl:203-229."""

    CFL: float
    ϵ_rel: float
    ϵ_abs: float
    Re: float
    r: float
    Vpdτ: float
    θ_dτ: float
    ηdτ: float

    def __init__(self, li: Sequence[float], di: Sequence[float], *, ϵ_rel: float = 1.0e-6, ϵ_abs: float = 1.0e-12,
                 Re: float = 3 * math.pi, CFL: Optional[float] = None, r: float = 0.7):
        N = len(li)
        if CFL is None:
            CFL = 0.9 / math.sqrt(2.1) if N == 2 else 0.9 / math.sqrt(3.1)
        lτ = min(li)
        Vpdτ = min(di) * CFL
        θ_dτ = lτ * (r + 4 / 3) / (Re * Vpdτ)
        ηdτ = Vpdτ * lτ / Re
        self.CFL, self.ϵ_rel, self.ϵ_abs, self.Re, self.r = float(CFL), float(ϵ_rel), float(ϵ_abs), float(Re), float(r)
        self.Vpdτ, self.θ_dτ, self.ηdτ = float(Vpdτ), float(θ_dτ), float(ηdτ)


# ----------------------------------------------------------------------------- boundary conditions
_FACES2 = ("left", "right", "top", "bot")
_FACES3 = ("left", "right", "front", "back", "top", "bot")


def _expand(bc: Optional[dict], nD: int, default=False) -> dict:
    faces = _FACES2 if nD == 2 else _FACES3
    out = {f: default for f in faces}
    if bc:
        for k, v in bc.items():
            if k not in _FACES3:
                raise KeyError(k)
            out[k] = v
    return out


def _check_periodic_pairs(periodic: dict, nD: int):
    pairs = (("left", "right"), ("bot", "top")) if nD == 2 else (("left", "right"), ("front", "back"), ("bot", "top"))
    for a, b in pairs:
        if bool(periodic.get(a, False)) != bool(periodic.get(b, False)):
            raise RuntimeError(f"Periodic boundary conditions must be paired: {a} and {b}")


class AbstractFlowBoundaryConditions:
    def __init__(self, *, no_slip=None, free_slip=None, periodic=None, free_surface: bool = False):
        given = [b for b in (no_slip, free_slip, periodic) if b is not None]
        nD = 3 if any(len(b) == 6 for b in given) else 2
        self.nD = nD
        self.no_slip = _expand(no_slip, nD, False)
        self.free_slip = _expand(free_slip, nD, True) if free_slip is None else _expand(free_slip, nD, False)
        self.periodic = _expand(periodic, nD, False)
        self.free_surface = bool(free_surface)
        # check_flow_bcs  src/boundaryconditions/types.jl:167-186
        _check_periodic_pairs(self.periodic, nD)
        for k in self.no_slip:
            if sum(1 for v in (self.no_slip[k], self.free_slip[k], self.periodic[k]) if v is True) > 1:
                raise RuntimeError(f"Incompatible boundary conditions on the {k} boundary")
        if self.free_surface and self.periodic.get("top", False):
            raise RuntimeError("Incompatible boundary conditions: the top can't be both periodic and free_surface")

    def flags(self, name: str):
        d = getattr(self, name)
        return [int(bool(d.get(f, False))) for f in _FACES3]


class VelocityBoundaryConditions(AbstractFlowBoundaryConditions):
    """src/boundaryconditions/types.jl:139-157"""


class DisplacementBoundaryConditions(AbstractFlowBoundaryConditions):
    """src/boundaryconditions/types.jl:110-128"""


class TemperatureBoundaryConditions:
    """src/boundaryconditions/types.jl:65-99.  Values: False = inactive, True/number = active."""

    def __init__(self, *, no_flux=None, constant_flux=None, constant_value=None, periodic=None, dirichlet=None):
        given = [b for b in (no_flux, constant_flux, constant_value, periodic) if b is not None]
        nD = 3 if any(len(b) == 6 for b in given) else 2
        self.nD = nD
        self.no_flux = _expand(no_flux if no_flux is not None else dict(left=True), 3)
        self.constant_flux = _expand(constant_flux, 3)
        self.constant_value = _expand(constant_value, 3)
        self.periodic = _expand(periodic, 3)
        self.dirichlet = dirichlet  # (constant, mask) or None
        _check_periodic_pairs(self.periodic, nD)
        for k, v in self.periodic.items():
            if v and any(c[k] is not False for c in (self.no_flux, self.constant_flux, self.constant_value)):
                raise RuntimeError(f"Incompatible boundary conditions on the {k} boundary")


# ----------------------------------------------------------------------------- grid
def _linrange(a: float, b: float, n: int) -> np.ndarray:
    """Julia LinRange(a, b, n): element i = (1 - t) a + t b with t = i/(n-1)."""
    if n == 1:
        return np.array([a], dtype=np.float64)
    t = np.arange(n, dtype=np.float64) / (n - 1)
    return (1.0 - t) * a + t * b


@dataclass
class IGG:
    """IGG(me, dims, nprocs, coords, comm_cart) — src/grid/Grid.jl:18-24 (+ overlaps of ImplicitGlobalGrid)."""

    me: int = 0
    dims: Sequence[int] = (1, 1, 1)
    nprocs: int = 1
    coords: Sequence[int] = (0, 0, 0)
    comm_cart: object = None
    overlaps: Sequence[int] = (2, 2, 2)
    nxyz: Optional[Sequence[int]] = None   # local cell counts given to init_global_grid (IGG keeps them as nxyz)
    periods: Sequence[int] = (0, 0, 0)     # periodx, periody, periodz of init_global_grid

    def n_g(self, ni: Sequence[int]):
        """nx_g(), ny_g(), nz_g(): dims·(n − overlap) + overlap·(period == 0)  (ImplicitGlobalGrid init_global_grid)."""
        return tuple(int(self.dims[d] * (ni[d] - self.overlaps[d]) + (0 if self.periods[d] else self.overlaps[d])) for d in range(len(ni)))


class Geometry:
    """Staggered Cartesian grid — src/grid/Cartesian.jl:9-19.

    Geometry(ni, li; origin): uniform grid (src/grid/Cartesian.jl:42-58, geometry_nonMPI / geometry_MPI src/grid/Grid.jl:56-116).
    Geometry.from_vertices(xv1, xv2[, xv3]) ≙ Geometry(TA, xvi...) / Geometry((xv1, xv2, …)): grid from explicit vertex coordinates
    (refined / non-uniform meshes, src/grid/Cartesian.jl:76-99): di / _di are VECTORS per direction (`center` = spacing of the cell
    centres, n−1 entries; `vertex` = cell widths, n entries; `velocity[c][d]` = spacing of component c's staggered grid along d).
    `uniform` tells the two apart; the B200 kernels take scalar spacings, so the solvers refuse a non-uniform grid loudly
    (`require_uniform`) — SURVEY §8f-3."""

    def __init__(self, ni: Sequence[int], li: Sequence[float], *, origin: Optional[Sequence[float]] = None, igg: Optional[IGG] = None):
        nD = len(ni)
        self.ni = tuple(int(n) for n in ni)
        self.li = tuple(float(l) for l in li)
        self.origin = tuple(float(o) for o in (origin if origin is not None else (0.0,) * nD))
        self.max_li = max(self.li)
        self.uniform = True
        igg = igg or IGG()
        ni_g = igg.n_g(self.ni)
        self.ni_g = ni_g
        di = tuple(self.li[d] / ni_g[d] for d in range(nD))
        self.di = SimpleNamespace(center=di, vertex=di, velocity=tuple(di for _ in range(nD)))
        _di = tuple(1.0 / d for d in di)
        self._di = SimpleNamespace(center=_di, vertex=_di, velocity=tuple(_di for _ in range(nD)))
        xci, xvi = [], []
        for d in range(nD):
            n, dx = self.ni[d], di[d]
            if igg.nprocs > 1 or any(k > 1 for k in igg.dims):
                # x_g(i, dx, n) = (coord·(n − overlap) + (i − 1))·dx  (grid/Utils.jl:1-75)
                x0 = igg.coords[d] * (n - igg.overlaps[d]) * dx + self.origin[d]
                xe_c = x0 + (n - 1) * dx
                xe_v = x0 + n * dx
                xci.append(_linrange(x0 + dx / 2, xe_c + dx / 2, n))
                xvi.append(_linrange(x0, xe_v, n + 1))
            else:
                xci.append(_linrange(self.origin[d] + dx / 2, self.origin[d] + self.li[d] - dx / 2, n))
                xvi.append(_linrange(self.origin[d], self.origin[d] + self.li[d], n + 1))
        self.xci, self.xvi = tuple(xci), tuple(xvi)
        # velocity_grids(xci, xvi, di) for scalar spacings  src/grid/Grid.jl:161-169, 184-200: component c lives on the vertices along c and on
        # the centres extended by one ghost point on either side along the transverse directions
        ghost = tuple(_linrange(self.xci[d][0] - di[d], self.xci[d][-1] + di[d], self.ni[d] + 2) for d in range(nD))
        self.xi_vel = tuple(tuple(self.xvi[d] if d == c else ghost[d] for d in range(nD)) for c in range(nD))

    @classmethod
    def from_vertices(cls, *xvi):
        """Geometry(TA, xvi...) / Geometry((xv1, xv2, …))  src/grid/Cartesian.jl:76-99; velocity_grids for vector spacings Grid.jl:171-182, 202-215."""
        if len(xvi) == 1 and isinstance(xvi[0], (tuple, list)) and len(xvi[0]) in (2, 3) and np.ndim(xvi[0][0]) == 1:
            xvi = tuple(xvi[0])
        xvi = tuple(np.asarray(x, dtype=np.float64) for x in xvi)
        nD = len(xvi)
        if nD not in (2, 3) or any(x.ndim != 1 or x.size < 2 for x in xvi):
            raise ValueError("Geometry.from_vertices: one 1-D vertex-coordinate vector (≥ 2 entries) per dimension, 2 or 3 dimensions")
        g = cls.__new__(cls)
        g.uniform = False
        g.ni = tuple(int(x.size) - 1 for x in xvi)
        g.ni_g = g.ni
        g.xvi = xvi
        g.xci = tuple((x[:-1] + x[1:]) / 2 for x in xvi)
        lims = tuple((float(x.min()), float(x.max())) for x in xvi)
        g.li = tuple(hi - lo for lo, hi in lims)
        g.max_li = max(g.li)
        g.origin = tuple(lo for lo, _ in lims)
        di_vertex = tuple(np.diff(x) for x in xvi)
        di_center = tuple(np.diff(x) for x in g.xci)
        # ghost points: the first / last centre spacing outwards (dxW = di_center[d][1], dxE = di_center[d][end])
        ghost = tuple(np.concatenate(([g.xci[d][0] - di_center[d][0]], g.xci[d], [g.xci[d][-1] + di_center[d][-1]])) for d in range(nD))
        g.xi_vel = tuple(tuple(g.xvi[d] if d == c else ghost[d] for d in range(nD)) for c in range(nD))
        di_vel = tuple(tuple(np.diff(x) for x in g.xi_vel[c]) for c in range(nD))
        g.di = SimpleNamespace(center=di_center, vertex=di_vertex, velocity=di_vel)
        g._di = SimpleNamespace(center=tuple(1.0 / x for x in di_center), vertex=tuple(1.0 / x for x in di_vertex),
                                velocity=tuple(tuple(1.0 / x for x in comp) for comp in di_vel))
        return g


def require_uniform(grid, what: str):
    """The B200 kernels take scalar grid spacings (vector `_di` is a next-row): refuse a non-uniform Geometry loudly."""
    if isinstance(grid, Geometry) and not grid.uniform:
        raise NotImplementedError(f"{what}: non-uniform grids (Geometry.from_vertices, vector spacings) are outside the B200 backend's subset")


def legacy_uniform_grid(ni, di, igg: Optional[IGG] = None) -> Geometry:
    """src/grid/Grid.jl:41-54"""
    igg = igg or IGG()
    ni_g = igg.n_g(ni)
    li = tuple(float(di[d]) * ni_g[d] for d in range(len(ni)))
    return Geometry(ni, li, igg=igg)


# ----------------------------------------------------------------------------- thermal containers
class ThermalArrays:
    """ThermalArrays(backend, ni) — src/types/heat_diffusion.jl:1-16, constructors/heat_diffusion.jl:38-120."""

    def __init__(self, backend_t, *ni):
        if len(ni) == 1 and isinstance(ni[0], (tuple, list)):
            ni = tuple(ni[0])
        ni = tuple(int(n) for n in ni)
        z = lambda *s: zeros(backend_t, *s)
        self.backend_type = backend_t
        self.ni = ni
        g = tuple(n + 2 for n in ni)
        self.T, self.Told, self.ΔT = z(*g), z(*g), z(*g)
        self.Tc, self.ΔTc = z(*ni), z(*ni)
        self.H, self.shear_heating, self.adiabatic, self.dT_dt, self.ResT = z(*ni), z(*ni), z(*ni), z(*ni), z(*ni)
        if len(ni) == 2:
            nx, ny = ni
            self.qTx, self.qTy, self.qTz = z(nx + 1, ny), z(nx, ny + 1), None
            self.qTx2, self.qTy2, self.qTz2 = z(nx + 1, ny), z(nx, ny + 1), None
        else:
            nx, ny, nz = ni
            self.qTx, self.qTy, self.qTz = z(nx + 1, ny, nz), z(nx, ny + 1, nz), z(nx, ny, nz + 1)
            self.qTx2, self.qTy2, self.qTz2 = z(nx + 1, ny, nz), z(nx, ny + 1, nz), z(nx, ny, nz + 1)


# ----------------------------------------------------------------------------- phase ratios
class PhaseRatios:
    """JustPIC.PhaseRatios as the B200 backend takes it: one array per staggered location of shape (nodes..., nphases),
    column-major, i.e. [phase][node] in memory (the CellArray layout flattened).  2D: center, vertex (+ Vx, Vy
    for the thermal solver); 3D: center, vertex, xy, yz, xz (+ Vx, Vy, Vz).  PhaseRatios(backend, nphases, ni) —
    src/phases/PhaseRatios.jl / JustPIC — allocates zeros; `from_arrays` wraps existing host/device arrays."""

    _names = ("center", "vertex", "xy", "yz", "xz", "Vx", "Vy", "Vz")

    def __init__(self, backend_t=None, nphases: int = 1, ni: Sequence[int] = ()):
        for nm in self._names:
            setattr(self, nm, None)
        self.nphases = int(nphases)
        if backend_t is None or not ni:
            return
        ni = tuple(int(n) for n in ni)
        nd = len(ni)
        shp = dict(center=ni, vertex=tuple(n + 1 for n in ni))
        for a, nm in enumerate(("Vx", "Vy", "Vz")[:nd]):
            shp[nm] = tuple(n + (1 if b == a else 0) for b, n in enumerate(ni))
        if nd == 3:
            nx, ny, nz = ni
            shp.update(xy=(nx + 1, ny + 1, nz), yz=(nx, ny + 1, nz + 1), xz=(nx + 1, ny, nz + 1))
        for nm, sh in shp.items():
            setattr(self, nm, zeros(backend_t, *sh, self.nphases))

    @classmethod
    def from_arrays(cls, backend_t, **arrays):
        pr = cls()
        for nm, a in arrays.items():
            if nm not in cls._names:
                raise KeyError(nm)
            setattr(pr, nm, None if a is None else PTArray(backend_t)(a))
            if a is not None:
                pr.nphases = int(a.shape[-1])
        return pr


def update_phase_ratios_(phase_ratios: PhaseRatios, phase_arrays, xci, xvi):
    """update_phase_ratios_2D!/update_phase_ratios_3D!(phase_ratios, phase_arrays, xci, xvi) — src/phases/PhaseRatios.jl:21-78 (GPU method
    src/ext/CUDA/3D.jl:519-539): fills every allocated location of `phase_ratios` from N phase arrays (B200 arrays with values in [0, 1])."""
    import ctypes as C

    from . import _abi
    from .stokes import context

    nd = phase_arrays[0].dim()
    ni = tuple(int(v) for v in phase_arrays[0].shape)
    for a in phase_arrays:
        if not is_device_array(a) or tuple(a.shape) != ni:
            raise ValueError("update_phase_ratios_: phase arrays must be B200 arrays of identical shape")
    if len(phase_arrays) != phase_ratios.nphases:
        raise ValueError(f"{len(phase_arrays)} phase arrays for PhaseRatios with {phase_ratios.nphases} phases")
    vp3 = lambda ptrs: (C.c_void_p * max(len(ptrs), 3))(*ptrs)
    xc = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in xci]
    xv = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in xvi]
    out = [getattr(phase_ratios, nm) for nm in ("center", "vertex", "Vx", "Vy", "Vz", "xy", "yz", "xz")]
    _abi.check(_abi.lib().jr_phase_ratios_from_arrays(context(), nd, _abi.i32x(list(ni) + [1] * (3 - nd)), len(phase_arrays),
                                                        vp3([data_ptr(a) for a in phase_arrays]), vp3([x.ctypes.data for x in xc]),
                                                        vp3([x.ctypes.data for x in xv]), *[None if a is None else data_ptr(a) for a in out]))


update_phase_ratios_2D_ = update_phase_ratios_3D_ = update_phase_ratios_
