"""cffi (ABI mode) binding of libb200stencil.so -- the only door into the CUDA kernels.

Mirrors the way the reference glues languages with CFFI
(/root/reference/src/tcn/py_ftn_interface/templates/interface.py.jinja2:1-110), in the other
direction: the reference embeds Python under a C symbol, this dlopen()s a C symbol table whose
``cdef`` is generated from the same kind of YAML (bridge/b200stencil.yaml).

There is NO CPU fallback: a missing library, a missing symbol or a non-zero status raises.
"""
from __future__ import annotations

import os
import threading
from typing import Any, Dict, List, Optional, Sequence, Tuple

import cffi

from .bridge.generate import PRECISIONS, Bridge, Function

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libb200stencil.so")


class B200StencilError(RuntimeError):
    """A C-ABI call returned a non-zero status (message from b2s_last_error())."""

    def __init__(self, symbol: str, status: int, message: str):
        super().__init__(f"{symbol} failed with status {status}: {message}")
        self.symbol, self.status, self.message = symbol, status, message


class LibraryMissing(ImportError):
    pass


_lock = threading.Lock()
_state: Dict[str, Any] = {}


def bridge() -> Bridge:
    if "bridge" not in _state:
        _state["bridge"] = Bridge.from_yaml()
    return _state["bridge"]


def load() -> Tuple[cffi.FFI, Any]:
    """dlopen the library once; loud failure if it has not been built."""
    with _lock:
        if "lib" not in _state:
            if not os.path.exists(LIB_PATH):
                raise LibraryMissing(
                    f"{LIB_PATH} not found. Build it with `python -m b200stencil.build` "
                    "(or __graft_entry__.build()); b200stencil has no CPU fallback."
                )
            ffi = cffi.FFI()
            ffi.cdef(bridge().emit_cdef())
            lib = ffi.dlopen(LIB_PATH)
            if lib.b2s_abi_version() != 1:
                raise LibraryMissing(f"{LIB_PATH}: ABI version {lib.b2s_abi_version()} != 1, rebuild")
            _state["ffi"], _state["lib"] = ffi, lib
        return _state["ffi"], _state["lib"]


def last_error() -> str:
    ffi, lib = load()
    return ffi.string(lib.b2s_last_error()).decode()


def check(symbol: str, status: int) -> None:
    if status != 0:
        raise B200StencilError(symbol, status, last_error())


def init(device: int) -> None:
    """b2s_init: bind the library to a device; raises when there is no sm_100 GPU."""
    _, lib = load()
    check("b2s_init", lib.b2s_init(int(device)))
    _state["device"] = int(device)


def ensure_init(device: int) -> None:
    if _state.get("device") != int(device):
        init(device)


def launch_count() -> int:
    _, lib = load()
    return int(lib.b2s_launch_count())


def set_option(name: str, value: int) -> None:
    _, lib = load()
    if lib.b2s_set_option(name.encode(), int(value)) != 0:
        raise KeyError(f"unknown libb200stencil option {name!r}")


def get_option(name: str) -> int:
    _, lib = load()
    return int(lib.b2s_get_option(name.encode()))


# ---- argument marshalling ----------------------------------------------------------------------

_TORCH_DTYPES: Dict[str, Any] = {}


def _torch_dtype(ctype: str):
    import torch

    if not _TORCH_DTYPES:
        _TORCH_DTYPES.update(
            {"double": torch.float64, "float": torch.float32, "int64_t": torch.int64, "int32_t": torch.int32, "int": torch.int32}
        )
    return _TORCH_DTYPES[ctype]


def field_args(t, dims: int, name: str, origin: Optional[Sequence[int]] = None) -> List[Any]:
    """(ptr, strides...) of a torch CUDA tensor indexed [b,] i, j [, k] with i-stride 1.

    ``origin`` = index of compute cell (0, 0[, 0]) inside the tensor (halo offset).
    """
    ffi, _ = load()
    if t.dim() == dims:
        sb, core = 0, t
    elif t.dim() == dims + 1:
        sb, core = t.stride(0), t[0] if t.shape[0] > 0 else t
    else:
        raise ValueError(f"{name}: expected a tensor indexed [b,]i,j{',k' if dims == 3 else ''}; got {tuple(t.shape)}")
    if core.dim() >= 1 and core.shape[0] > 1 and core.stride(0) != 1:
        raise ValueError(
            f"{name}: fields must be i-fastest (stride(i) == 1, got {core.stride(0)}); "
            "allocate with b200stencil.fields or QuantityFactory"
        )
    off = 0
    if origin is not None:
        off = sum(int(o) * core.stride(d) for d, o in enumerate(origin))
    ptr = ffi.cast("void*", t.data_ptr() + off * t.element_size())
    strides = [core.stride(1), core.stride(2), sb] if dims == 3 else [core.stride(1), sb]
    return [ptr] + [int(s) for s in strides]


class PreparedCall:
    """A C-ABI call with its arguments already validated and marshalled.

    ``prepare()`` does the per-tensor checks once; ``__call__`` only appends the stream and calls the
    symbol (about a microsecond of host time instead of tens), which matters when a step is a handful
    of 50-microsecond kernels.  The tensors are kept alive by the object; it must be rebuilt if a
    field is reallocated.
    """

    __slots__ = ("symbol", "_fn", "_args", "_device", "_keep", "_ffi")

    def __init__(self, symbol, fn, args, device, keep, ffi):
        self.symbol, self._fn, self._args, self._device, self._keep, self._ffi = symbol, fn, args, device, keep, ffi

    def __call__(self, stream: Optional[int] = None) -> None:
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(self._device).cuda_stream
        status = self._fn(*self._args, self._ffi.cast("void*", int(stream)))
        if status != 0:
            raise B200StencilError(self.symbol, status, last_error())


def prepare(fn_name: str, precision: str, values: Dict[str, Any],
            origins: Optional[Dict[str, Sequence[int]]] = None) -> PreparedCall:
    """Validate and marshal the arguments of ``b2s_<fn>[_f32]_c`` once; returns a :class:`PreparedCall`."""
    import torch

    ffi, lib = load()
    fn: Function = bridge().functions[fn_name]
    symbol = fn.symbol(bridge().prefix, precision)
    args: List[Any] = []
    keep: List[Any] = []
    device = None
    for a in fn.arguments:
        v = values[a.name]
        if not a.is_array:
            args.append(v)
            continue
        ctype = a.element_ctype(precision)
        if v is None:  # optional output: NULL pointer (zero strides); the library rejects a NULL required field
            args.append(ffi.NULL)
            if (a.dims or 1) != 1:
                args += [0] * (3 if a.dims == 3 else 2)
            continue
        if not isinstance(v, torch.Tensor) or not v.is_cuda:
            raise TypeError(f"{symbol}: {a.name} must be a torch CUDA tensor (device storage only, no CPU path)")
        if v.dtype != _torch_dtype(ctype):
            raise TypeError(f"{symbol}: {a.name} has dtype {v.dtype}, the ABI wants {ctype} (no type casting at the boundary)")
        device = v.device if device is None else device
        if v.device != device:
            raise ValueError(f"{symbol}: {a.name} lives on {v.device}, other fields on {device}")
        keep.append(v)
        if (a.dims or 1) == 1:
            args.append(ffi.cast(f"{ctype}*", v.data_ptr()))
        else:
            fa = field_args(v, a.dims, a.name, (origins or {}).get(a.name))
            fa[0] = ffi.cast(f"{ctype}*", fa[0])
            args += fa
    if device is None:
        raise ValueError(f"{symbol}: no device field among the arguments")
    ensure_init(device.index if device.index is not None else torch.cuda.current_device())
    return PreparedCall(symbol, getattr(lib, symbol), args, device, keep, ffi)


def call(fn_name: str, precision: str, values: Dict[str, Any], stream: Optional[int] = None,
         origins: Optional[Dict[str, Sequence[int]]] = None) -> None:
    """Call ``b2s_<fn>[_f32]_c`` with YAML-ordered arguments taken from ``values``.

    Scalars are passed by value; tensors are expanded to (device pointer, strides) after checking
    device, dtype and layout.  ``stream`` is a raw cudaStream_t (default: torch's current stream).
    """
    prepare(fn_name, precision, values, origins)(stream)


def precision_of(t) -> str:
    import torch

    if t.dtype == torch.float64:
        return "double"
    if t.dtype == torch.float32:
        return "float"
    raise TypeError(f"unsupported field dtype {t.dtype}: the kernels compute in float64 or float32")


__all__ = [
    "B200StencilError", "LibraryMissing", "LIB_PATH", "PRECISIONS", "bridge", "call", "check", "ensure_init",
    "field_args", "get_option", "prepare", "PreparedCall", "init", "last_error", "launch_count", "load", "precision_of", "set_option",
]  # fmt: skip
