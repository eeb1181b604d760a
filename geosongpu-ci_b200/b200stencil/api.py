"""NDSL-shaped host API: the call surface the reference patterns use, backed by sm_100a kernels.

Surface kept:
  get_factories_single_tile_numpy(nx, ny, nz, nhalo)      Do__get_top_of_the_column.py:28-30
  StencilFactory.from_dims_halo(func, compute_dims)        Do__get_top_of_the_column.py:49-52
  StencilFactory.config.dace_config                        Do__get_top_of_the_column.py:47
  QuantityFactory.zeros(dims, units, dtype=Float)          Do__get_top_of_the_column.py:48
  Quantity.view[...] read/write, printable                 WIP__hybrid_index_2dout.py:72-90
  orchestrate(obj=..., config=...)                         Do__get_top_of_the_column.py:47
Semantics kept (recalled from NDSL, not vendored): compute origin (nhalo, nhalo, 0), domain
(nx, ny, nz); Quantity storage padded to (nx+2h+1, ny+2h+1, nz+1) with ``.view`` the compute window;
raw arrays of exactly the domain shape are accepted as stencil arguments; writes land only inside
origin...origin+domain.

Differences: storage is a torch CUDA tensor, i-fastest; ``func`` is not compiled but dispatched to a
hand-written kernel (registry.py); NumPy arguments are staged to the device and copied back into the
caller's array (the reference asserts read the caller's ``O``, Do__get_top_of_the_column.py:68).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import fields, registry
from .constants import X_DIM, X_INTERFACE_DIM, Y_DIM, Y_INTERFACE_DIM, Z_DIM, Z_INTERFACE_DIM
from .typing import Float

_TORCH = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
          np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32}  # fmt: skip


def _torch_dtype(dtype) -> torch.dtype:
    if isinstance(dtype, torch.dtype):
        return dtype
    return _TORCH[np.dtype(dtype)]


def default_device() -> torch.device:
    """Storage device: cuda when there is one.  Without a GPU, storage can still be built (layout tests)
    but every stencil call raises -- there is no CPU compute path."""
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


@dataclass
class DaceConfig:
    """Placeholder for ``stencil_factory.config.dace_config`` (orchestration is a no-op here)."""

    is_dace_orchestrated: bool = False


@dataclass
class StencilConfig:
    backend: str = "b200:cuda"
    dace_config: DaceConfig = field(default_factory=DaceConfig)


@dataclass
class GridIndexing:
    domain: Tuple[int, int, int]
    n_halo: int

    @property
    def origin(self) -> Tuple[int, int, int]:
        return (self.n_halo, self.n_halo, 0)


def orchestrate(obj=None, config=None, **kwargs) -> None:
    """No-op (reference: ``ndsl.orchestrate``; only meaningful for dace:* backends)."""
    return None


class _View:
    """``Quantity.view[...]``: read/write window on the compute domain."""

    def __init__(self, q: "Quantity"):
        self._q = q

    def _window(self) -> torch.Tensor:
        return self._q.compute_view()

    def __getitem__(self, idx):
        return self._window()[idx]

    def __setitem__(self, idx, value):
        w = self._window()
        if isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value)).to(w.device, dtype=w.dtype)
        elif isinstance(value, torch.Tensor):
            value = value.to(w.device, dtype=w.dtype)
        w[idx] = value

    def __repr__(self):
        return repr(self._window().cpu().numpy())


class Quantity:
    """Halo-padded field with named dims (reference: ``ndsl.Quantity`` as used in WIP__hybrid_index_2dout.py)."""

    def __init__(self, data: torch.Tensor, dims: Sequence[str], units: str, origin: Sequence[int], extent: Sequence[int]):
        self.data, self.dims, self.units = data, tuple(dims), units
        self.origin, self.extent = tuple(origin), tuple(extent)
        self.view = _View(self)

    def compute_view(self) -> torch.Tensor:
        idx = tuple(slice(o, o + e) for o, e in zip(self.origin, self.extent))
        return self.data[idx]

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def dtype(self):
        return self.data.dtype

    def __repr__(self):
        return f"Quantity(dims={self.dims}, units={self.units!r}, extent={self.extent}, device={self.data.device})"


class QuantityFactory:
    def __init__(self, indexing: GridIndexing, device: Optional[torch.device] = None):
        self.indexing = indexing
        self.device = device or default_device()

    def _geometry(self, dims: Sequence[str]):
        nx, ny, nz = self.indexing.domain
        h = self.indexing.n_halo
        shape, origin, extent = [], [], []
        for d in dims:
            if d in (X_DIM, X_INTERFACE_DIM):
                shape.append(nx + 2 * h + 1), origin.append(h), extent.append(nx + (d == X_INTERFACE_DIM))
            elif d in (Y_DIM, Y_INTERFACE_DIM):
                shape.append(ny + 2 * h + 1), origin.append(h), extent.append(ny + (d == Y_INTERFACE_DIM))
            elif d in (Z_DIM, Z_INTERFACE_DIM):
                shape.append(nz + 1), origin.append(0), extent.append(nz + (d == Z_INTERFACE_DIM))
            else:
                raise ValueError(f"unknown dimension {d!r}")
        return shape, origin, extent

    def _make(self, dims, units, dtype, fill) -> Quantity:
        if list(dims[:2]) not in ([X_DIM, Y_DIM], [X_INTERFACE_DIM, Y_DIM], [X_DIM, Y_INTERFACE_DIM]):
            raise ValueError("b200stencil fields are indexed [x, y(, z)]")
        shape, origin, extent = self._geometry(dims)
        data = fields.empty(shape, _torch_dtype(dtype), self.device, fill=fill)
        return Quantity(data, dims, units, origin, extent)

    def zeros(self, dims: Sequence[str], units: str, dtype=Float) -> Quantity:
        return self._make(dims, units, dtype, 0)

    def ones(self, dims: Sequence[str], units: str, dtype=Float) -> Quantity:
        return self._make(dims, units, dtype, 1)

    def empty(self, dims: Sequence[str], units: str, dtype=Float) -> Quantity:
        return self._make(dims, units, dtype, None)


class FrozenStencil:
    """Callable returned by the factory: marshals Quantity / tensor / ndarray arguments to device
    fields on the compute domain and runs the kernel the definition resolves to."""

    def __init__(self, func: Callable, kernel_name: str, origin, domain, device: torch.device):
        self.func, self.kernel_name = func, kernel_name
        self.origin, self.domain = tuple(origin), tuple(domain)
        self.device = device
        self._adapter = registry.adapter(kernel_name)

    def _to_device(self, a):
        """-> (device field on the compute domain, write-back closure or None)."""
        if isinstance(a, Quantity):
            return a.compute_view(), None
        if isinstance(a, torch.Tensor):
            return a, None
        if isinstance(a, np.ndarray):
            want = self.domain[: a.ndim]
            if tuple(a.shape) != tuple(want):
                raise ValueError(f"raw array argument has shape {a.shape}, the compute domain is {want}")
            dev = fields.from_numpy(a, device=self.device)

            def back(dev=dev, a=a):
                a[...] = dev.cpu().numpy()

            return dev, back
        return a, None  # scalars

    def __call__(self, *args, **kwargs):
        if self.device.type != "cuda":
            raise RuntimeError(
                f"stencil '{self.kernel_name}' needs a CUDA device: b200stencil has no CPU fallback "
                "(storage was created on the CPU because no GPU is visible)"
            )
        staged = [self._to_device(a) for a in args]
        kstaged = {k: self._to_device(v) for k, v in kwargs.items()}
        self._adapter(*[s[0] for s in staged], **{k: v[0] for k, v in kstaged.items()})
        for _, back in list(staged) + list(kstaged.values()):
            if back is not None:
                back()  # NumPy callers read their own arrays afterwards


class StencilFactory:
    def __init__(self, config: StencilConfig, indexing: GridIndexing, device: Optional[torch.device] = None):
        self.config, self.grid_indexing = config, indexing
        self.device = device or default_device()

    @property
    def backend(self) -> str:
        return self.config.backend

    def from_dims_halo(self, func: Callable, compute_dims: Sequence[str], compute_halos: Sequence[int] = (),
                       kernel: Optional[str] = None, **kwargs) -> FrozenStencil:
        nx, ny, nz = self.grid_indexing.domain
        domain = []
        for d in compute_dims:
            domain.append({X_DIM: nx, X_INTERFACE_DIM: nx + 1, Y_DIM: ny, Y_INTERFACE_DIM: ny + 1,
                           Z_DIM: nz, Z_INTERFACE_DIM: nz + 1}[d])  # fmt: skip
        return self.from_origin_domain(func, self.grid_indexing.origin[: len(domain)], tuple(domain), kernel=kernel)

    def from_origin_domain(self, func: Callable, origin, domain, kernel: Optional[str] = None, **kwargs) -> FrozenStencil:
        return FrozenStencil(func, registry.resolve(func, kernel), origin, domain, self.device)


def get_factories_single_tile(nx: int, ny: int, nz: int, nhalo: int, device=None) -> Tuple[StencilFactory, QuantityFactory]:
    """One tile, layout 1x1 (reference: ``ndsl.boilerplate.get_factories_single_tile_numpy``)."""
    idx = GridIndexing((int(nx), int(ny), int(nz)), int(nhalo))
    dev = torch.device(device) if device is not None else default_device()
    return StencilFactory(StencilConfig(), idx, dev), QuantityFactory(idx, dev)


# The reference spells the backend in the function name; kept so pattern files run unmodified.  The
# backend is the B200 one all the same.
get_factories_single_tile_numpy = get_factories_single_tile
