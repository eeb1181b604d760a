"""NumPy restatement of the stencil hot path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function takes arrays indexed ``[i, j, k]`` (IJ fields ``[i, j]``) of exactly
the compute-domain shape unless stated otherwise, works for any memory layout
(C-order like the reference's numpy backend, or i-fastest views), and writes its
outputs in place, the way a gt4py stencil call does.

gt4py numpy-backend semantics honoured here (recalled from gt4py.cartesian):
  (i)   PARALLEL: each statement is applied to the whole interval before the next;
  (ii)  FORWARD/BACKWARD: sequential in k, all statements per level;
  (iii) ``interval(a, b)`` is a Python slice on k; ``interval(...)`` is all levels;
  (iv)  IJ fields broadcast over k and are written only in FORWARD/BACKWARD;
  (v)   scalar locals are per-gridpoint temporaries;
  (vi)  ``field[0, 0, expr]`` is a relative, unchecked variable-K read;
  (viii) assigning an int to a Float field casts.

Parity status: S1/S2 pinned by the reference asserts, everything else
"parity unpinned" (oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# S1  dsl_patterns/Do__get_top_of_the_column.py:33-38
# --------------------------------------------------------------------------------------


def top_of_column(PLEmb: np.ndarray, PLEmb_top: np.ndarray, out_field: np.ndarray) -> None:
    """``stencil`` of Do__get_top_of_the_column.py:33-38.

    with computation(FORWARD), interval(-1, None): PLEmb_top = PLEmb      (:34-35)
    with computation(PARALLEL), interval(...):     out_field = PLEmb_top  (:37-38)
    """
    nk = PLEmb.shape[2]
    for k in range(nk - 1, nk):  # interval(-1, None)
        PLEmb_top[:, :] = PLEmb[:, :, k]
    out_field[:, :, :] = PLEmb_top[:, :, None]


# --------------------------------------------------------------------------------------
# S2  dsl_patterns/Do__while_in_gt_functions.py:22-32
# --------------------------------------------------------------------------------------


class UndefinedBehaviour(IndexError):
    """A variable-K read left the column (undefined in gt4py)."""


def while_in_function(in_field: np.ndarray, out_field: np.ndarray, threshold: float = 4.0) -> None:
    """``stencil`` + ``while_in_function`` of Do__while_in_gt_functions.py:22-32, literally.

    Per point: ``lev = 0; while field[0,0,lev] < 4: lev += 1; return lev`` (:23-27),
    assigned (cast to Float) to ``out_field`` under PARALLEL (:30-32).  Restated the way
    the numpy backend runs a ``while``: ``while np.any(mask)`` with masked updates.
    Raises UndefinedBehaviour where the reference would read past the last level.
    """
    ni, nj, nk = in_field.shape
    ii, jj, kk = np.meshgrid(np.arange(ni), np.arange(nj), np.arange(nk), indexing="ij")
    lev = np.zeros((ni, nj, nk), dtype=np.int64)
    mask = in_field < threshold  # field[0,0,0] < 4
    while np.any(mask):
        lev[mask] += 1
        if np.any((kk + lev)[mask] >= nk):
            raise UndefinedBehaviour("while_in_function read below the last level")
        probe = in_field[ii, jj, np.minimum(kk + lev, nk - 1)]
        mask = mask & (probe < threshold)
    out_field[:, :, :] = lev  # int -> Float cast on assignment


def while_in_function_scan(in_field: np.ndarray, out_field: np.ndarray, threshold: float = 4.0) -> int:
    """Single BACKWARD-scan equivalent of :func:`while_in_function`.

    ``nxt = k if in[k] >= 4 else nxt; out[k] = nxt - k``.  Identical results whenever the
    literal form is defined; where it is not (no level >= 4 at or below k) this writes
    ``nk - k`` and counts the point.  Returns the number of undefined points.
    """
    ni, nj, nk = in_field.shape
    nxt = np.full((ni, nj), nk, dtype=np.int64)
    undefined = 0
    for k in range(nk - 1, -1, -1):
        hit = ~(in_field[:, :, k] < threshold)  # NaN counts as a hit, as `while NaN < 4` stops
        nxt = np.where(hit, k, nxt)
        undefined += int(np.count_nonzero(nxt == nk))
        out_field[:, :, k] = nxt - k
    return undefined


# --------------------------------------------------------------------------------------
# S3  dsl_patterns/WIP__hybrid_index_2dout.py:34-42
# --------------------------------------------------------------------------------------


def hybrid_index_2dout(
    data_field: np.ndarray, k_mask: np.ndarray, k_index_desired: np.ndarray, out_field: np.ndarray
) -> None:
    """``stencil`` of WIP__hybrid_index_2dout.py:34-42.

    with computation(FORWARD), interval(...):
        if k_mask == k_index_desired: out_field = data_field
    Last match wins; columns with no match keep their previous ``out_field``.
    """
    nk = data_field.shape[2]
    for k in range(nk):
        m = k_mask[:, :, k] == k_index_desired
        out_field[m] = data_field[:, :, k][m]


# --------------------------------------------------------------------------------------
# S4  GEOS moist-physics-style column stencils.  No reference source: this is the spec
#     (names from geos_documentation/moist/GFDL_1M.drawio:76,111 and the
#     Fortran quoted at WIP__hybrid_index_2dout.py:10-15, Do__get_top_of_the_column.py:9-14).
#     k = nk-1 is the surface ("LM"), pressure grows with k.
# --------------------------------------------------------------------------------------


def find_klcl(PLmb: np.ndarray, PLCL: np.ndarray, KLCL: np.ndarray, PLmb_at_KLCL: np.ndarray) -> None:
    """S4a: bottom-up search for the first level whose pressure is <= the LCL pressure.

    with computation(BACKWARD), interval(...):
        if found == 0 and PLmb <= PLCL: KLCL = k; PLmb_at_KLCL = PLmb; found = 1
    KLCL = -1 and PLmb_at_KLCL untouched when no level qualifies (Fortran's KLCL = 0).
    """
    nk = PLmb.shape[2]
    found = np.zeros(PLCL.shape, dtype=bool)
    KLCL[:, :] = -1
    for k in range(nk - 1, -1, -1):
        m = (~found) & (PLmb[:, :, k] <= PLCL)
        KLCL[m] = k
        PLmb_at_KLCL[m] = PLmb[:, :, k][m]
        found |= m


SAT_EPS = 0.622  # Rd / Rv
SAT_LV = 2.5e6  # J / kg
SAT_CP = 1004.0  # J / kg / K
SAT_NEWTON_STEPS = 2


def saturation_adjust(T: np.ndarray, q: np.ndarray, ql: np.ndarray, p: np.ndarray) -> None:
    """S4b: pointwise saturation adjustment, PARALLEL, in place on T, q, ql.

    es   = 611.2 * exp(17.67 * (T - 273.15) / (T - 29.65))
    qs   = eps * es / (p - (1 - eps) * es)
    dqs  = eps * p * des / (p - (1 - eps) * es)**2,  des = es * 17.67 * 243.5 / (T - 29.65)**2
    dq   = max((q - qs) / (1 + (Lv/cp) * dqs), -ql)
    T += (Lv/cp) * dq;  q -= dq;  ql += dq               (two fixed Newton steps)
    """
    dt = T.dtype.type
    eps, lcp = dt(SAT_EPS), dt(SAT_LV / SAT_CP)
    one = dt(1.0)
    for _ in range(SAT_NEWTON_STEPS):
        tm = T - dt(29.65)
        es = dt(611.2) * np.exp(dt(17.67) * (T - dt(273.15)) / tm)
        den = p - (one - eps) * es
        qs = eps * es / den
        des = es * dt(17.67 * 243.5) / (tm * tm)
        dqs = eps * p * des / (den * den)
        dq = (q - qs) / (one + lcp * dqs)
        dq = np.maximum(dq, -ql)
        T += lcp * dq
        q -= dq
        ql += dq


CLOUD_QL_MIN = 1.0e-8


def cloud_top(ql: np.ndarray, ktop: np.ndarray, ql_min: float = CLOUD_QL_MIN) -> None:
    """S4c: smallest k with ql[k] > ql_min (k = 0 is the model top); -1 if the column is clear.

    with computation(FORWARD), interval(...):
        if found == 0 and ql > ql_min: ktop = k; found = 1
    """
    nk = ql.shape[2]
    found = np.zeros(ktop.shape, dtype=bool)
    ktop[:, :] = -1
    for k in range(nk):
        m = (~found) & (ql[:, :, k] > ql.dtype.type(ql_min))
        ktop[m] = k
        found |= m


# --------------------------------------------------------------------------------------
# S5  FV3-style horizontal finite-volume transport (unlimited PPM, 3-cell halo).
#     No reference source: this is the spec (argument names from
#     src/tcn/py_ftn_interface/example_def_dycore.yaml:38,53,66-69; halo width 3 from
#     src/tcn/validation/serialbox/serialbox_dat_to_netcdf.py:161).
# --------------------------------------------------------------------------------------

FV_HALO = 3


def _ppm_flux(qm3, qm2, qm1, q0, qp1, qp2, c):
    """Flux through the interface between cells -1 and 0 given q at cells -3..+2 and Courant c."""
    dt = q0.dtype.type
    c7, c1, one = dt(7.0 / 12.0), dt(1.0 / 12.0), dt(1.0)
    al_m1 = c7 * (qm2 + qm1) - c1 * (qm3 + q0)  # west interface of cell -1
    al_0 = c7 * (qm1 + q0) - c1 * (qm2 + qp1)  # west interface of cell 0
    al_p1 = c7 * (q0 + qp1) - c1 * (qm1 + qp2)  # west interface of cell +1
    # upwind cell -1 (c > 0)
    bl_m = al_m1 - qm1
    br_m = al_0 - qm1
    b0_m = bl_m + br_m
    f_pos = qm1 + (one - c) * (br_m - c * b0_m)
    # upwind cell 0 (c <= 0)
    bl_0 = al_0 - q0
    br_0 = al_p1 - q0
    b0_0 = bl_0 + br_0
    f_neg = q0 + (one + c) * (bl_0 + c * b0_0)
    return np.where(c > 0, f_pos, f_neg)


def fv_tp2d(
    q: np.ndarray,
    crx: np.ndarray,
    xfx: np.ndarray,
    cry: np.ndarray,
    yfx: np.ndarray,
    rarea: np.ndarray,
    q_out: np.ndarray,
) -> None:
    """S5: one flux-form PPM transport step.

    q      (ni+6, nj+6, nk)  halo 3 on every side, halo already filled
    crx,xfx (ni+1, nj, nk)   Courant number / area flux at x-interfaces
    cry,yfx (ni, nj+1, nk)   same at y-interfaces
    rarea  (ni, nj)          reciprocal cell area
    q_out  (ni, nj, nk)      q - rarea * (fx[i+1]*xfx[i+1] - fx[i]*xfx[i] + fy[j+1]*yfx[j+1] - fy[j]*yfx[j])

    al = 7/12 (q[-1] + q) - 1/12 (q[-2] + q[+1]);  bl = al - q;  br = al[+1] - q;  b0 = bl + br
    flux(c > 0)  = q[-1] + (1 - c) (br[-1] - c b0[-1]);  flux(c <= 0) = q + (1 + c) (bl + c b0)
    """
    h = FV_HALO
    ni, nj, nk = q_out.shape
    assert q.shape == (ni + 2 * h, nj + 2 * h, nk)
    assert crx.shape == xfx.shape == (ni + 1, nj, nk)
    assert cry.shape == yfx.shape == (ni, nj + 1, nk)

    def xs(off):  # q at x-interface I + off, I = 0..ni, compute rows
        return q[h + off : h + off + ni + 1, h : h + nj, :]

    def ys(off):
        return q[h : h + ni, h + off : h + off + nj + 1, :]

    fx = _ppm_flux(xs(-3), xs(-2), xs(-1), xs(0), xs(1), xs(2), crx)
    fy = _ppm_flux(ys(-3), ys(-2), ys(-1), ys(0), ys(1), ys(2), cry)
    fxx = fx * xfx
    fyy = fy * yfx
    div = (fxx[1:, :, :] - fxx[:-1, :, :]) + (fyy[:, 1:, :] - fyy[:, :-1, :])
    q_out[:, :, :] = q[h : h + ni, h : h + nj, :] - rarea[:, :, None] * div


# --------------------------------------------------------------------------------------
# S6  vertical column scans of the dycore.  No reference source: this is the spec
#     (field names delp/pe from example_def_dycore.yaml:52-58).
# --------------------------------------------------------------------------------------


def pe_prefix(delp: np.ndarray, ptop: float, pe: np.ndarray) -> None:
    """S6a: pe[0] = ptop; pe[k+1] = pe[k] + delp[k]  (FORWARD; pe has nk+1 levels)."""
    nk = delp.shape[2]
    pe[:, :, 0] = pe.dtype.type(ptop)
    for k in range(nk):
        pe[:, :, k + 1] = pe[:, :, k] + delp[:, :, k]


def remap_column(pe1, q1, pe2):
    """S6b for ONE column, plain Python: the definition the vectorised forms must match."""
    nk1, nk2 = len(q1), len(pe2) - 1
    q2 = np.zeros(nk2, dtype=q1.dtype)
    k1 = 0
    for k2 in range(nk2):
        lo, hi = pe2[k2], pe2[k2 + 1]
        while k1 < nk1 - 1 and pe1[k1 + 1] <= lo:
            k1 += 1
        acc = q1.dtype.type(0.0)
        kk = k1
        while True:
            a = max(lo, pe1[kk])
            b = min(hi, pe1[kk + 1])
            if b > a:
                acc = acc + (b - a) * q1[kk]
            if pe1[kk + 1] >= hi or kk == nk1 - 1:
                break
            kk += 1
        k1 = kk
        q2[k2] = acc / (hi - lo)
    return q2


def remap(pe1: np.ndarray, q1: np.ndarray, pe2: np.ndarray, q2: np.ndarray) -> None:
    """S6b: conservative piecewise-constant remap of q1 (layers between edges pe1) onto pe2.

    FORWARD over target layers with a source pointer marching monotonically (a ``while`` with
    variable-K reads, i.e. all three dsl_patterns features).  Vectorised over columns with
    masked updates, same operation order per column as :func:`remap_column`.
    """
    ni, nj, nk1 = q1.shape
    nk2 = q2.shape[2]
    ii, jj = np.meshgrid(np.arange(ni), np.arange(nj), indexing="ij")
    k1 = np.zeros((ni, nj), dtype=np.int64)
    for k2 in range(nk2):
        lo, hi = pe2[:, :, k2], pe2[:, :, k2 + 1]
        adv = (k1 < nk1 - 1) & (pe1[ii, jj, k1 + 1] <= lo)
        while np.any(adv):
            k1 = k1 + adv
            adv = (k1 < nk1 - 1) & (pe1[ii, jj, k1 + 1] <= lo)
        acc = np.zeros((ni, nj), dtype=q1.dtype)
        kk = k1.copy()
        active = np.ones((ni, nj), dtype=bool)
        while np.any(active):
            top, bot = pe1[ii, jj, kk], pe1[ii, jj, kk + 1]
            a = np.maximum(lo, top)
            b = np.minimum(hi, bot)
            add = active & (b > a)
            acc = np.where(add, acc + (b - a) * q1[ii, jj, kk], acc)
            done = (bot >= hi) | (kk == nk1 - 1)
            active = active & ~done
            kk = kk + active
        k1 = kk
        q2[:, :, k2] = acc / (hi - lo)


def tridiag(a: np.ndarray, b: np.ndarray, c: np.ndarray, d: np.ndarray, x: np.ndarray) -> None:
    """S6c: Thomas algorithm, a[k] x[k-1] + b[k] x[k] + c[k] x[k+1] = d[k] per column.

    FORWARD eliminate:  m = b - a cp[-1];  cp = c / m;  dp = (d - a dp[-1]) / m
    BACKWARD substitute: x = dp - cp x[+1]
    (named at geos_documentation/moist/GF.drawio:502)
    """
    nk = b.shape[2]
    cp = np.empty_like(b)
    dp = np.empty_like(b)
    cp[:, :, 0] = c[:, :, 0] / b[:, :, 0]
    dp[:, :, 0] = d[:, :, 0] / b[:, :, 0]
    for k in range(1, nk):
        m = b[:, :, k] - a[:, :, k] * cp[:, :, k - 1]
        cp[:, :, k] = c[:, :, k] / m
        dp[:, :, k] = (d[:, :, k] - a[:, :, k] * dp[:, :, k - 1]) / m
    x[:, :, nk - 1] = dp[:, :, nk - 1]
    for k in range(nk - 2, -1, -1):
        x[:, :, k] = dp[:, :, k] - cp[:, :, k] * x[:, :, k + 1]
