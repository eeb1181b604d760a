"""Device storage for fields: torch CUDA tensors, i-fastest.

A field indexed ``[i, j, k]`` (or ``[b, i, j, k]`` for a batch of tiles / sub-domains) is a
permuted view of contiguous ``(b, k, j, i)`` storage, so ``stride(i) == 1`` -- the layout the
reference obtains from Fortran memory without a copy
(/root/reference/src/tcn/py_ftn_interface/templates/data_conversion.py:134-148) and the one gt4py's
GPU backends use.  Rows are padded so that ``stride(j)`` is a multiple of 16 bytes, which lets
every kernel use 16-byte vector accesses and TMA tiles.  PyTorch is used for memory only.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

ROW_ALIGN_BYTES = 16


def _pad_i(ni: int, dtype: torch.dtype, align_rows: bool) -> int:
    if not align_rows:
        return ni
    per = ROW_ALIGN_BYTES // torch.empty((), dtype=dtype).element_size()
    return (ni + per - 1) // per * per


def empty(shape: Sequence[int], dtype=torch.float64, device="cuda", batch: Optional[int] = None,
          align_rows: bool = True, fill: Optional[float] = None) -> torch.Tensor:
    """Uninitialised (or filled) field indexed [i,j(,k)] / [b,i,j(,k)], stored i-fastest."""
    shape = tuple(int(s) for s in shape)
    ni = shape[0]
    nip = _pad_i(ni, dtype, align_rows)
    store = (nip,) + shape[1:]
    full = tuple(reversed(store))
    if batch is not None:
        full = (int(batch),) + full
    base = torch.empty(full, dtype=dtype, device=device) if fill is None else torch.full(full, fill, dtype=dtype, device=device)
    nd = len(shape)
    if batch is None:
        view = base.permute(*reversed(range(nd)))
        return view[:ni]
    view = base.permute(0, *reversed(range(1, nd + 1)))
    return view[:, :ni]


def zeros(shape, dtype=torch.float64, device="cuda", batch=None, align_rows=True) -> torch.Tensor:
    return empty(shape, dtype, device, batch, align_rows, fill=0)


def from_numpy(a, device="cuda", dtype=None, align_rows: bool = True, pinned_stage: Optional[torch.Tensor] = None,
               stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """Upload a NumPy array indexed [i,j(,k)] (any layout) into an i-fastest device field."""
    t = torch.from_numpy(a) if not isinstance(a, torch.Tensor) else a
    out = empty(t.shape, dtype or t.dtype, device, None, align_rows)
    out.copy_(t, non_blocking=False)
    return out


def to_numpy(t: torch.Tensor):
    return t.detach().cpu().numpy()


def is_ifirst(t: torch.Tensor) -> bool:
    return t.dim() >= 1 and (t.shape[0] <= 1 or t.stride(0) == 1)


def halo_view(t: torch.Tensor, halo: int) -> torch.Tensor:
    """Compute-domain window of a field that carries ``halo`` cells on every horizontal side."""
    if halo == 0:
        return t
    if t.dim() in (2, 3):
        return t[halo:-halo, halo:-halo]
    raise ValueError("halo_view expects a field indexed [i,j(,k)]")


def shape3(t: torch.Tensor) -> Tuple[int, int, int, int]:
    """(ni, nj, nk, nb) of a 3-D field indexed [b,]i,j,k."""
    if t.dim() == 3:
        return t.shape[0], t.shape[1], t.shape[2], 1
    if t.dim() == 4:
        return t.shape[1], t.shape[2], t.shape[3], t.shape[0]
    raise ValueError(f"expected a field indexed [b,]i,j,k, got shape {tuple(t.shape)}")
