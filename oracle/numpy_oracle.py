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


def _xppm(q, c):
    """Unlimited-PPM x-fluxes: q [ni+6, nj, nk] (3-cell halo in i), c [ni+1, nj, nk] -> [ni+1, nj, nk]."""
    n = c.shape[0]
    return _ppm_flux(q[0:n], q[1 : n + 1], q[2 : n + 2], q[3 : n + 3], q[4 : n + 4], q[5 : n + 5], c)


def _yppm(q, c):
    """Unlimited-PPM y-fluxes: q [ni, nj+6, nk] (3-cell halo in j), c [ni, nj+1, nk] -> [ni, nj+1, nk]."""
    n = c.shape[1]
    return _ppm_flux(q[:, 0:n], q[:, 1 : n + 1], q[:, 2 : n + 2], q[:, 3 : n + 3], q[:, 4 : n + 4], q[:, 5 : n + 5], c)


def corner_fill_source_local(li: int, lj: int, ni: int, nj: int, direction: int):
    """[recalled] FV3 copy_corners, in the 0-based compute coordinates of a sub-domain of ni x nj cells whose halo
    corner (li, lj) sits at a cube corner: the halo cell whose value the corner cell takes for sweeps in x
    (direction 1) or y (direction 2).  1-based FV3 forms: SW q(i,j) = q(j,1-i) | q(1-j,i); SE q(npy-j,i-npx+1) |
    q(npy+j-1,npx-i); NE q(j,2npx-1-i) | q(2npy-1-j,i); NW q(npy-j,i-1+npx) | q(j+1-npx,npy-i)."""
    w, s = li < 0, lj < 0
    if direction == 1:
        if w and s:
            return lj, -li - 1
        if not w and s:
            return ni - 1 - lj, li - ni
        if not w and not s:
            return lj - nj + ni, ni + nj - 1 - li
        return nj - 1 - lj, li + nj
    if w and s:
        return -lj - 1, li
    if not w and s:
        return ni + lj, ni - 1 - li
    if not w and not s:
        return ni + nj - 1 - lj, li - ni + nj
    return lj - nj, nj - 1 - li


def copy_corners(q: np.ndarray, direction: int, flags: int = 15, h: int = FV_HALO) -> None:
    """Fill, in place, the halo-corner blocks of q [ni+2h, nj+2h(, nk)] marked in ``flags`` (1 SW, 2 SE, 4 NW, 8 NE)
    from the edge halos, FV3 copy_corners rule for ``direction``."""
    ni, nj = q.shape[0] - 2 * h, q.shape[1] - 2 * h
    for bit, irange, jrange in ((1, range(-h, 0), range(-h, 0)), (2, range(ni, ni + h), range(-h, 0)),
                                (4, range(-h, 0), range(nj, nj + h)), (8, range(ni, ni + h), range(nj, nj + h))):  # fmt: skip
        if not flags & bit:
            continue
        for li in irange:
            for lj in jrange:
                si, sj = corner_fill_source_local(li, lj, ni, nj, direction)
                q[li + h, lj + h] = q[si + h, sj + h]


def fv_tp2d_split(
    q: np.ndarray,
    crx: np.ndarray,
    xfx: np.ndarray,
    cry: np.ndarray,
    yfx: np.ndarray,
    area: np.ndarray,
    rarea: np.ndarray,
    q_out: np.ndarray,
    fx_out: np.ndarray = None,
    fy_out: np.ndarray = None,
    corner_flags: int = 0,
) -> None:
    """S5b: FV3's fv_tp_2d -- the inner/outer operator splitting of Lin & Rood that
    removes the directional-splitting error of S5's plain sum of 1-D fluxes.  [recalled] from FV3
    tp_core.F90 fv_tp_2d with the unlimited PPM of S5 as xppm/yppm; no source in /root/reference.

    q        (ni+6, nj+6, nk)  3-cell halo on every side INCLUDING the corners
    crx, xfx (ni+1, nj+6, nk)  Courant number / area flux at x-interfaces, rows -3 .. nj+2
    cry, yfx (ni+6, nj+1, nk)  same at y-interfaces, columns -3 .. ni+2
    area     (ni+6, nj+6)      cell area with halo;  rarea (ni, nj) its reciprocal on the compute domain
    q_out    (ni, nj, nk);  fx_out (ni+1, nj, nk), fy_out (ni, nj+1, nk) optional: the averaged fluxes

      fy2 = yppm(q, cry)                         inner y-sweep, every column incl. the i-halo
      q_i = (q area + yfx fy2 [j] - yfx fy2 [j+1]) / (area + yfx[j] - yfx[j+1])      advected in y
      fx  = xppm(q_i, crx)                       outer x-sweep on the y-advected field
      fx2 = xppm(q, crx)                         inner x-sweep, every row incl. the j-halo
      q_j = (q area + xfx fx2 [i] - xfx fx2 [i+1]) / (area + xfx[i] - xfx[i+1])      advected in x
      fy  = yppm(q_j, cry)                       outer y-sweep on the x-advected field
      fx <- 0.5 (fx + fx2) xfx ;  fy <- 0.5 (fy + fy2) yfx
      q_out = q + rarea (fx[i] - fx[i+1] + fy[j] - fy[j+1])

    ``corner_flags`` (1 SW | 2 SE | 4 NW | 8 NE): halo corners of this sub-domain that are cube corners.  FV3 calls
    copy_corners(q, dir=2) before the inner y-sweep and copy_corners(q, dir=1) before the inner x-sweep; here the
    corner cells of ``q`` already hold the direction-1 values (the halo update writes them) and the y-sweep runs on a
    copy whose flagged corners are re-filled with the direction-2 rule.
    """
    h = FV_HALO
    ni, nj, nk = q_out.shape
    dt = q.dtype.type
    assert q.shape == (ni + 2 * h, nj + 2 * h, nk) and area.shape == (ni + 2 * h, nj + 2 * h)
    assert crx.shape == xfx.shape == (ni + 1, nj + 2 * h, nk)
    assert cry.shape == yfx.shape == (ni + 2 * h, nj + 1, nk)
    ci, cj = slice(h, h + ni), slice(h, h + nj)
    a3 = area[:, :, None]
    qy = q
    if corner_flags:
        qy = q.copy()
        copy_corners(qy, 2, corner_flags)
    fy2 = _yppm(qy, cry)  # (ni+6, nj+1, nk)
    fyy = yfx * fy2
    ra_y = a3[:, cj] + (yfx[:, :-1] - yfx[:, 1:])
    q_i = (q[:, cj] * a3[:, cj] + (fyy[:, :-1] - fyy[:, 1:])) / ra_y  # (ni+6, nj, nk)
    fx = _xppm(q_i, crx[:, cj])  # (ni+1, nj, nk)
    fx2 = _xppm(q, crx)  # (ni+1, nj+6, nk)
    fxx = xfx * fx2
    ra_x = a3[ci] + (xfx[:-1] - xfx[1:])
    q_j = (q[ci] * a3[ci] + (fxx[:-1] - fxx[1:])) / ra_x  # (ni, nj+6, nk)
    fy = _yppm(q_j, cry[ci])  # (ni, nj+1, nk)
    fx = dt(0.5) * (fx + fx2[:, cj]) * xfx[:, cj]
    fy = dt(0.5) * (fy + fy2[ci]) * yfx[ci]
    q_out[:, :, :] = q[ci, cj] + rarea[:, :, None] * ((fx[:-1] - fx[1:]) + (fy[:, :-1] - fy[:, 1:]))
    if fx_out is not None:
        fx_out[:, :, :] = fx
    if fy_out is not None:
        fy_out[:, :, :] = fy


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


# --------------------------------------------------------------------------------------
# S6d  PPM vertical remap ("PPM-limited map_single remap (kord)").
#      No source in /root/reference.  [recalled] from FV3's fv_mapz.F90: ppm_profile (4th-order
#      interface values from monotonised slopes, area-preserving cubics at the top and the
#      surface, ppm_limiters lmt = 0 | 1 | 2) and map1_ppm (integration of the piecewise parabola
#      f(s) = AL + s [(AR - AL) + A6 (1 - s)] over the target layers).  This restatement is the
#      specification: parity unpinned.
# --------------------------------------------------------------------------------------

PPM_R3, PPM_R23, PPM_R12 = 1.0 / 3.0, 2.0 / 3.0, 1.0 / 12.0


def _sign(a, b):
    """Fortran SIGN(a, b): |a| with the sign of b (b = +0 counts as positive)."""
    return np.where(b >= 0, np.abs(a), -np.abs(a))


def _ppm_limiters(dm, q, al, ar, a6, lmt):
    """ppm_limiters of fv_mapz.F90 on one layer of many columns; returns (al, ar, a6)."""
    dt = q.dtype.type
    if lmt == 3:
        return al, ar, a6
    if lmt == 0:  # standard PPM constraint
        flat = dm == 0
        da1 = ar - al
        da2 = da1 * da1
        a6da = a6 * da1
        lo = a6da < -da2
        hi = (~lo) & (a6da > da2)
        a6n = np.where(lo, dt(3.0) * (al - q), np.where(hi, dt(3.0) * (ar - q), a6))
        arn = np.where(lo, al - a6n, ar)
        aln = np.where(hi, ar - a6n, al)
        return np.where(flat, q, aln), np.where(flat, q, arn), np.where(flat, dt(0.0), a6n)
    if lmt == 1:  # improved full monotonicity constraint (Lin 2004)
        qmp = dt(2.0) * dm
        aln = q - _sign(np.minimum(np.abs(qmp), np.abs(al - q)), qmp)
        arn = q + _sign(np.minimum(np.abs(qmp), np.abs(ar - q)), qmp)
        return aln, arn, dt(3.0) * (dt(2.0) * q - (aln + arn))
    # lmt == 2: positive definite constraint
    with np.errstate(divide="ignore", invalid="ignore"):
        fmin = q + dt(0.25) * (ar - al) ** 2 / a6 + a6 * dt(PPM_R12)
    act = (np.abs(ar - al) < -a6) & (fmin < 0)
    c1 = act & (q < ar) & (q < al)
    c2 = act & ~c1 & (ar > al)
    c3 = act & ~c1 & ~c2
    a6n = np.where(c1, dt(0.0), np.where(c2, dt(3.0) * (al - q), np.where(c3, dt(3.0) * (ar - q), a6)))
    arn = np.where(c1, q, np.where(c2, al - a6n, ar))
    aln = np.where(c1, q, np.where(c3, ar - a6n, al))
    return aln, arn, a6n


def ppm_profile(q: np.ndarray, delp: np.ndarray, kord: int = 4, iv: int = 1):
    """[recalled] FV3 fv_mapz.F90 ppm_profile: (AL, AR, A6) of every layer, arrays [i, j, k], km >= 4.

    kord 4 | 5 | 6 -> interior limiter lmt = kord - 3 (1 monotone, 2 positive definite, 3 none); iv = 0
    (positive definite scalar) caps lmt at 2 and clips the boundary edge values at 0; iv = 1 otherwise.
    The two top and two bottom layers always use the standard constraint (lmt = 0).
    """
    dt = q.dtype.type
    km = q.shape[2]
    assert km >= 4 and kord in (4, 5, 6) and iv in (0, 1)
    delq = q[:, :, 1:] - q[:, :, :-1]  # delq[k] = q[k+1] - q[k]
    d4 = np.zeros_like(q)
    d4[:, :, 1:] = delp[:, :, :-1] + delp[:, :, 1:]  # d4[k] = delp[k-1] + delp[k], k >= 1
    dc = np.zeros_like(q)
    for k in range(1, km - 1):
        c1 = (delp[:, :, k - 1] + dt(0.5) * delp[:, :, k]) / d4[:, :, k + 1]
        c2 = (delp[:, :, k + 1] + dt(0.5) * delp[:, :, k]) / d4[:, :, k]
        df2 = delp[:, :, k] * (c1 * delq[:, :, k] + c2 * delq[:, :, k - 1]) / (d4[:, :, k] + delp[:, :, k + 1])
        qmax = np.maximum(np.maximum(q[:, :, k - 1], q[:, :, k]), q[:, :, k + 1]) - q[:, :, k]
        qmin = q[:, :, k] - np.minimum(np.minimum(q[:, :, k - 1], q[:, :, k]), q[:, :, k + 1])
        dc[:, :, k] = _sign(np.minimum(np.minimum(np.abs(df2), qmax), qmin), df2)
    al = np.zeros_like(q)
    ar = np.zeros_like(q)
    for k in range(2, km - 1):  # 4th-order interpolation of the provisional edge value
        c1 = delq[:, :, k - 1] * delp[:, :, k - 1] / d4[:, :, k]
        a1 = d4[:, :, k - 1] / (d4[:, :, k] + delp[:, :, k - 1])
        a2 = d4[:, :, k + 1] / (d4[:, :, k] + delp[:, :, k])
        al[:, :, k] = q[:, :, k - 1] + c1 + dt(2.0) / (d4[:, :, k - 1] + d4[:, :, k + 1]) * (
            delp[:, :, k] * (c1 * (a1 - a2) + a2 * dc[:, :, k - 1]) - delp[:, :, k - 1] * a1 * dc[:, :, k])
    # top: area-preserving cubic with zero second derivative at the boundary
    d1, d2 = delp[:, :, 0], delp[:, :, 1]
    qm = (d2 * q[:, :, 0] + d1 * q[:, :, 1]) / (d1 + d2)
    dq = dt(2.0) * (q[:, :, 1] - q[:, :, 0]) / (d1 + d2)
    c1 = dt(4.0) * (al[:, :, 2] - qm - d2 * dq) / (d2 * (dt(2.0) * d2 * d2 + d1 * (d2 + dt(3.0) * d1)))
    c3 = dq - dt(0.5) * c1 * (d2 * (dt(5.0) * d1 + d2) - dt(3.0) * d1 * d1)
    al1 = qm - dt(0.25) * c1 * d1 * d2 * (d2 + dt(3.0) * d1)
    al0 = d1 * (dt(2.0) * c1 * d1 * d1 - c3) + al1
    al1 = np.maximum(al1, np.minimum(q[:, :, 0], q[:, :, 1]))  # no over- and undershoot
    al1 = np.minimum(al1, np.maximum(q[:, :, 0], q[:, :, 1]))
    dc[:, :, 0] = dt(0.5) * (al1 - q[:, :, 0])
    if iv == 0:
        al0, al1 = np.maximum(dt(0.0), al0), np.maximum(dt(0.0), al1)
    al[:, :, 0], al[:, :, 1] = al0, al1
    # bottom: the same cubic at the surface
    d1, d2 = delp[:, :, km - 1], delp[:, :, km - 2]
    qm = (d2 * q[:, :, km - 1] + d1 * q[:, :, km - 2]) / (d1 + d2)
    dq = dt(2.0) * (q[:, :, km - 2] - q[:, :, km - 1]) / (d1 + d2)
    c1 = (al[:, :, km - 2] - qm - d2 * dq) / (d2 * (dt(2.0) * d2 * d2 + d1 * (d2 + dt(3.0) * d1)))
    c3 = dq - dt(2.0) * c1 * (d2 * (dt(5.0) * d1 + d2) - dt(3.0) * d1 * d1)
    alb = qm - c1 * d1 * d2 * (d2 + dt(3.0) * d1)
    arb = d1 * (dt(8.0) * c1 * d1 * d1 - c3) + alb
    alb = np.maximum(alb, np.minimum(q[:, :, km - 1], q[:, :, km - 2]))
    alb = np.minimum(alb, np.maximum(q[:, :, km - 1], q[:, :, km - 2]))
    dc[:, :, km - 1] = dt(0.5) * (q[:, :, km - 1] - alb)
    if iv == 0:
        alb, arb = np.maximum(dt(0.0), alb), np.maximum(dt(0.0), arb)
    al[:, :, km - 1] = alb
    ar[:, :, : km - 1] = al[:, :, 1:]
    ar[:, :, km - 1] = arb
    a6 = np.zeros_like(q)
    lmt = max(0, kord - 3)
    if iv == 0:
        lmt = min(2, lmt)
    for k in range(km):
        edge = k < 2 or k >= km - 2
        a6k = dt(3.0) * (dt(2.0) * q[:, :, k] - (al[:, :, k] + ar[:, :, k]))
        al[:, :, k], ar[:, :, k], a6[:, :, k] = _ppm_limiters(dc[:, :, k], q[:, :, k], al[:, :, k], ar[:, :, k], a6k, 0 if edge else lmt)
    return al, ar, a6


def remap_ppm_column(pe1, q, al, ar, a6, pe2):
    """[recalled] FV3 map1_ppm for ONE column, plain Python: the definition of the integration.

    Precondition (FV3's): both edge arrays strictly increasing and pe1[0] <= pe2[0], pe2[-1] <= pe1[-1].
    """
    dt = q.dtype.type
    km, kn = len(q), len(pe2) - 1
    r3, r23 = dt(PPM_R3), dt(PPM_R23)
    q2 = np.zeros(kn, dtype=q.dtype)
    k0 = 0
    for k in range(kn):
        top, bot = pe2[k], pe2[k + 1]
        for l in range(k0, km):
            if top >= pe1[l] and top <= pe1[l + 1]:
                dp1 = pe1[l + 1] - pe1[l]
                pl = (top - pe1[l]) / dp1
                if bot <= pe1[l + 1]:  # the target layer lies inside source layer l
                    pr = (bot - pe1[l]) / dp1
                    q2[k] = al[l] + dt(0.5) * (a6[l] + ar[l] - al[l]) * (pr + pl) - a6[l] * r3 * (pr * (pr + pl) + pl * pl)
                    k0 = l
                else:
                    qsum = (pe1[l + 1] - top) * (al[l] + dt(0.5) * (a6[l] + ar[l] - al[l]) * (dt(1.0) + pl)
                                                 - a6[l] * (r3 * (dt(1.0) + pl * (dt(1.0) + pl))))
                    for m in range(l + 1, km):
                        if bot > pe1[m + 1]:  # whole layer
                            qsum = qsum + (pe1[m + 1] - pe1[m]) * q[m]
                        else:
                            dp = bot - pe1[m]
                            esl = dp / (pe1[m + 1] - pe1[m])
                            qsum = qsum + dp * (al[m] + dt(0.5) * esl * (ar[m] - al[m] + a6[m] * (dt(1.0) - r23 * esl)))
                            k0 = m
                            break
                    q2[k] = qsum / (bot - top)
                break
    return q2


def remap_ppm(pe1: np.ndarray, q1: np.ndarray, pe2: np.ndarray, q2: np.ndarray, kord: int = 4, iv: int = 1) -> None:
    """S6d: PPM remap of q1 (layers between edges pe1) onto the layers between pe2 (see the section header)."""
    ni, nj, _ = q1.shape
    delp = pe1[:, :, 1:] - pe1[:, :, :-1]
    al, ar, a6 = ppm_profile(q1, delp, kord, iv)
    for i in range(ni):
        for j in range(nj):
            q2[i, j, :] = remap_ppm_column(pe1[i, j], q1[i, j], al[i, j], ar[i, j], a6[i, j], pe2[i, j])


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
