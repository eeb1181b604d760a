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


POINTWISE = []  # one record per assert_close call; tests/conftest.py writes them to gpurun_out/ at session end


def assert_close(got, want, rtol, name=""):
    """|got - want| <= rtol * max(|want|, max|want|): relative to the field's magnitude (the asserted bound).

    The POINTWISE relative error |got - want| / |want| is recorded beside it for every call (worst value, and the
    share of points within rtol pointwise): fields that span many decades (ql: 0 .. 1e-3) carry cancellation error
    near their zeros that no implementation can avoid, so the pointwise figure is reported, not asserted; for fields
    of uniform magnitude (q_out, x, q2) the two bounds coincide."""
    got_dtype = str(np.asarray(got).dtype)
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    scale = np.abs(want).max() if want.size else 0.0
    bound = rtol * np.maximum(np.abs(want), scale)
    err = np.abs(got - want)
    bad = err > bound
    if want.size:
        nz = np.abs(want) > 0
        rel = np.zeros_like(err)
        rel[nz] = err[nz] / np.abs(want[nz])
        rel[~nz & (err > 0)] = np.inf
        POINTWISE.append({
            "name": name, "dtype": got_dtype, "rtol": rtol, "points": int(want.size), "field_max": float(scale),
            "field_min_abs_nonzero": float(np.abs(want[nz]).min()) if nz.any() else 0.0,
            "worst_abs_err": float(err.max()), "worst_err_over_field_max": float(err.max() / scale) if scale else 0.0,
            "worst_pointwise_rel": float(rel.max()), "share_within_rtol_pointwise": float((rel <= rtol).mean()),
            "bit_identical": bool((err == 0).all()),
        })  # fmt: skip
    assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} beyond rtol={rtol}; worst {err.max():.3e} vs bound {bound[err.argmax()] if bound.ndim else bound:.3e}"
