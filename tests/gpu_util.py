"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch

from b200stencil import fields
from oracle import inputs as gen


def up(a: np.ndarray, align_rows=True) -> torch.Tensor:
    """NumPy [i,j(,k)] -> i-fastest device field."""
    return fields.from_numpy(np.ascontiguousarray(a), align_rows=align_rows)


def up_batch(arrs, align_rows=True) -> torch.Tensor:
    """list of NumPy [i,j(,k)] -> device field [b,i,j(,k)]."""
    t = fields.empty(arrs[0].shape, dtype=torch.from_numpy(np.empty(0, arrs[0].dtype)).dtype, batch=len(arrs), align_rows=align_rows)
    for b, a in enumerate(arrs):
        t[b].copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return t


def down(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def zeros_like_np(shape, dtype):
    a = gen.ifirst_empty(shape, dtype)
    a[...] = 0
    return a


def assert_close(got, want, rtol, name=""):
    """|got - want| <= rtol * max(|want|, max|want|): relative to the field's magnitude."""
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    scale = np.abs(want).max() if want.size else 0.0
    bound = rtol * np.maximum(np.abs(want), scale)
    err = np.abs(got - want)
    bad = err > bound
    assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} beyond rtol={rtol}; worst {err.max():.3e} vs bound {bound[err.argmax()] if bound.ndim else bound:.3e}"
