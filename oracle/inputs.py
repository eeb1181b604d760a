"""Seeded synthetic inputs for every BASELINE config -- TEST INFRASTRUCTURE ONLY.

Follows SURVEY.md 8(d) "Synthetic inputs": ``np.random.default_rng(20240724 + cfg)``,
generated in fp64 and cast for fp32 runs.  Arrays come back indexed ``[i, j, k]`` but stored
i-fastest (the layout of the device fields, cf. the zero-copy i-fastest view the reference
builds from Fortran memory at src/tcn/py_ftn_interface/templates/data_conversion.py:141).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 20240724


def ifirst_empty(shape, dtype=np.float64) -> np.ndarray:
    """Uninitialised array indexed [i,j(,k)] whose memory is i-fastest."""
    return np.empty(tuple(reversed(shape)), dtype=dtype).transpose()


def as_ifirst(a: np.ndarray, dtype=None) -> np.ndarray:
    out = ifirst_empty(a.shape, dtype or a.dtype)
    out[...] = a
    return out


def rng_for(cfg: int) -> np.random.Generator:
    return np.random.default_rng(SEED_BASE + cfg)


# ---- golden vectors of the reference's own asserts ------------------------------------


def golden_column_input(domain=(3, 3, 4), dtype=np.float64) -> np.ndarray:
    """I = ones; I[:, :, -1] = 42  (Do__get_top_of_the_column.py:59-60, Do__while_in_gt_functions.py:53-54)."""
    I = np.ones(domain[0] * domain[1] * domain[2], dtype=dtype).reshape(domain)
    I[:, :, domain[2] - 1] = 42
    return I


# ---- cfg1 / cfg2: dsl_patterns ----------------------------------------------------------


def top_of_column_inputs(ni, nj, nk, dtype=np.float64, cfg=1):
    rng = rng_for(cfg)
    k = np.arange(nk, dtype=np.float64)
    PLEmb = 1000.0 * (k + 1.0)[None, None, :] / nk + rng.random((ni, nj, nk))
    return as_ifirst(PLEmb, dtype)


def while_inputs(ni, nj, nk, dtype=np.float64, cfg=2):
    """in = U(0, 3.999); 1-3 hit levels per column set to 4 + U(0, 38); in[..., nk-1] = 42."""
    rng = rng_for(cfg)
    f = rng.random((ni, nj, nk)) * 3.999
    nhits = rng.integers(1, 4, size=(ni, nj))
    ii, jj = np.meshgrid(np.arange(ni), np.arange(nj), indexing="ij")
    for h in range(3):
        lev = rng.integers(0, nk, size=(ni, nj))
        val = 4.0 + rng.random((ni, nj)) * 38.0
        sel = nhits > h
        f[ii[sel], jj[sel], lev[sel]] = val[sel]
    f[:, :, nk - 1] = 42.0
    return as_ifirst(f, dtype)


def hybrid_inputs(ni, nj, nk, dtype=np.float64, cfg=2, miss_fraction=0.0):
    """k_mask[..., k] = k; k_index = randint(0, nk); data = randint(800, 900) (WIP__hybrid_index_2dout.py:72-82)."""
    rng = rng_for(cfg + 100)
    k_mask = np.broadcast_to(np.arange(nk, dtype=np.float64)[None, None, :], (ni, nj, nk))
    k_index = rng.integers(0, nk, size=(ni, nj)).astype(np.float64)
    if miss_fraction > 0:
        k_index[rng.random((ni, nj)) < miss_fraction] = -1.0
    data = rng.integers(800, 900, size=(ni, nj, nk)).astype(np.float64)
    return as_ifirst(data, dtype), as_ifirst(k_mask, dtype), as_ifirst(k_index, dtype)


# ---- cfg3: moist-physics-style columns --------------------------------------------------


def _qs(T, p):
    es = 611.2 * np.exp(17.67 * (T - 273.15) / (T - 29.65))
    return 0.622 * es / (p - (1 - 0.622) * es)


def moist_inputs(ni, nj, nk, dtype=np.float64, cfg=3):
    """p [Pa] grows with k; T = 210 + 90 k/nk + N(0,2); q = U(0,1.2) qs; ql = max(0, N(0,1e-4)); PLCL = U(600,950) hPa."""
    rng = rng_for(cfg)
    k = np.arange(nk, dtype=np.float64)
    p = np.broadcast_to(100.0 * (100.0 + 900.0 * (k + 0.5) / nk)[None, None, :], (ni, nj, nk)).copy()
    T = 210.0 + 90.0 * (k / nk)[None, None, :] + rng.normal(0.0, 2.0, (ni, nj, nk))
    q = rng.random((ni, nj, nk)) * 1.2 * _qs(T, p)
    ql = np.maximum(0.0, rng.normal(0.0, 1e-4, (ni, nj, nk)))
    PLCL = 100.0 * (600.0 + 350.0 * rng.random((ni, nj)))
    return {
        "p": as_ifirst(p, dtype),
        "T": as_ifirst(T, dtype),
        "q": as_ifirst(q, dtype),
        "ql": as_ifirst(ql, dtype),
        "PLCL": as_ifirst(PLCL, dtype),
    }


# ---- cfg4: horizontal finite volume ------------------------------------------------------


def fv_inputs(ni, nj, nk, dtype=np.float64, cfg=4, halo=3):
    """q = 1 + 0.5 sin(2 pi i/ni) cos(2 pi j/nj) + 0.01 N(0,1), periodic halo; c = U(-0.9,0.9); xfx = c U(0.9,1.1)."""
    rng = rng_for(cfg)
    i = np.arange(ni, dtype=np.float64)
    j = np.arange(nj, dtype=np.float64)
    core = (
        1.0
        + 0.5 * np.sin(2 * np.pi * i / ni)[:, None, None] * np.cos(2 * np.pi * j / nj)[None, :, None]
        + 0.01 * rng.normal(0.0, 1.0, (ni, nj, nk))
    )
    q = np.pad(core, ((halo, halo), (halo, halo), (0, 0)), mode="wrap")
    crx = rng.uniform(-0.9, 0.9, (ni + 1, nj, nk))
    cry = rng.uniform(-0.9, 0.9, (ni, nj + 1, nk))
    xfx = crx * rng.uniform(0.9, 1.1, (ni + 1, nj, nk))
    yfx = cry * rng.uniform(0.9, 1.1, (ni, nj + 1, nk))
    rarea = rng.uniform(0.9, 1.1, (ni, nj))
    return {
        "q": as_ifirst(q, dtype),
        "crx": as_ifirst(crx, dtype),
        "xfx": as_ifirst(xfx, dtype),
        "cry": as_ifirst(cry, dtype),
        "yfx": as_ifirst(yfx, dtype),
        "rarea": as_ifirst(rarea, dtype),
    }


# ---- cfg5: vertical scans ---------------------------------------------------------------


def fv_split_inputs(ni, nj, nk, dtype=np.float64, cfg=8):
    """Inputs of fv_tp2d_split: q with a full 3-cell halo (corners included), Courant numbers and area fluxes
    on the halo-extended interface sets, cell areas with halo.  xfx = crx * dy * dx-ish so that the advected
    areas ra_x = area + xfx[i] - xfx[i+1] stay well away from zero (|c| <= 0.45 per direction)."""
    rng = rng_for(cfg)
    h = 3
    i = np.arange(-h, ni + h)[:, None, None]
    j = np.arange(-h, nj + h)[None, :, None]
    q = 1.0 + 0.5 * np.sin(2 * np.pi * i / max(ni, 1)) * np.cos(2 * np.pi * j / max(nj, 1)) + 0.01 * rng.standard_normal((ni + 2 * h, nj + 2 * h, nk))
    area = rng.uniform(0.9, 1.1, (ni + 2 * h, nj + 2 * h))
    crx = rng.uniform(-0.45, 0.45, (ni + 1, nj + 2 * h, nk))
    cry = rng.uniform(-0.45, 0.45, (ni + 2 * h, nj + 1, nk))
    xfx = crx * rng.uniform(0.9, 1.1, crx.shape)
    yfx = cry * rng.uniform(0.9, 1.1, cry.shape)
    rarea = 1.0 / area[h : h + ni, h : h + nj]
    f = {"q": q, "crx": crx, "xfx": xfx, "cry": cry, "yfx": yfx, "area": area, "rarea": rarea}
    return {k: as_ifirst(v.astype(dtype)) for k, v in f.items()}


def vertical_inputs(ni, nj, nk, dtype=np.float64, cfg=5, nk2=None, ptop=1.0):
    """delp = U(0.5,1.5) 1000e2/nk; pe1 = ptop + cumsum(delp); pe2 = uniform levels between ptop and pe1[nk]."""
    rng = rng_for(cfg)
    nk2 = nk if nk2 is None else nk2
    delp = (rng.uniform(0.5, 1.5, (ni, nj, nk)) * (1000.0e2 / nk)).astype(dtype)
    pe1 = np.empty((ni, nj, nk + 1), dtype=dtype)
    pe1[:, :, 0] = ptop
    for k in range(nk):  # same sequential order as the stencil, in the target dtype
        pe1[:, :, k + 1] = pe1[:, :, k] + delp[:, :, k]
    sig = (np.arange(nk2 + 1, dtype=np.float64) / nk2).astype(dtype)
    pe2 = (pe1[:, :, :1] + (pe1[:, :, -1:] - pe1[:, :, :1]) * sig[None, None, :]).astype(dtype)
    pe2[:, :, -1] = pe1[:, :, -1]
    q1 = (1.0 + rng.random((ni, nj, nk))).astype(dtype)
    return {
        "delp": as_ifirst(delp),
        "pe1": as_ifirst(pe1),
        "pe2": as_ifirst(pe2),
        "q1": as_ifirst(q1),
        "ptop": ptop,
    }


def ppm_inputs(ni, nj, nk, dtype=np.float64, cfg=7, nk2=None, smooth=True, positive=True):
    """Inputs of the PPM remap: the edges of :func:`vertical_inputs`; q1 either a smooth profile in pressure
    (two sines with column-dependent phase, plus 2 % noise so that limiters fire in places) or pure noise
    (every layer an extremum: the limiters flatten everything).  ``positive=False`` shifts the tracer so that
    it changes sign (the positive-definite limiter must then leave it alone where fmin >= 0 does not matter)."""
    v = vertical_inputs(ni, nj, nk, dtype, cfg=cfg, nk2=nk2)
    rng = rng_for(cfg + 100)
    pe1 = v["pe1"].astype(np.float64)
    pm = 0.5 * (pe1[:, :, 1:] + pe1[:, :, :-1]) / pe1[:, :, -1:]
    ph = rng.uniform(0, 2 * np.pi, (ni, nj, 1))
    if smooth:
        q = 1.0 + 0.6 * np.sin(2 * np.pi * pm + ph) + 0.3 * np.sin(7 * np.pi * pm * pm - ph) + 0.02 * rng.standard_normal((ni, nj, nk))
        q = np.maximum(q, 0.0) if positive else q - 1.0
    else:
        q = rng.uniform(0.0 if positive else -1.0, 1.0, (ni, nj, nk))
    v["q1"] = as_ifirst(q.astype(dtype))
    return v


def tridiag_inputs(ni, nj, nk, dtype=np.float64, cfg=6):
    """Diagonally dominant system: b = 2 + U(0,1), a, c = -U(0,1) (a[0] = c[nk-1] = 0), d = U(-1,1)."""
    rng = rng_for(cfg)
    a = -rng.random((ni, nj, nk))
    c = -rng.random((ni, nj, nk))
    a[:, :, 0] = 0.0
    c[:, :, nk - 1] = 0.0
    b = 2.0 + rng.random((ni, nj, nk))
    d = rng.uniform(-1.0, 1.0, (ni, nj, nk))
    return {k: as_ifirst(v, dtype) for k, v in dict(a=a, b=b, c=c, d=d).items()}
