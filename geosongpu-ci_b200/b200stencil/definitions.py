"""gtscript definitions of the stencils that have no source in the reference.

The three dsl_patterns files ARE gtscript definitions and dispatch to their kernels by AST hash (registry.py).
The moist-physics scans, the FV transport and the vertical solvers exist in the reference only as names
(BASELINE.json configs[2..4]); their specification is oracle/numpy_oracle.py.  This module states the same
specifications in the reference's own DSL, the way a GEOS developer would write them against NDSL:

  * each definition is tagged with the hand-written sm_100a kernel that implements it, so it can be handed to
    ``StencilFactory.from_dims_halo(func=definitions.find_klcl, ...)`` exactly like a pattern's ``stencil``;
  * the bodies are never executed by the product (there is no DSL compiler here).  They ARE executed by the test
    infrastructure: ``tests/golden/gtscript_interp.py`` runs them with gt4py numpy-backend semantics and
    ``tests/test_definitions_golden.py`` checks the oracle against the result -- an independent executable statement
    of every specification, with the arithmetic written in the oracle's operation order.

``THIS_K`` (the level index) is a gt4py >= 1.0.4 builtin; with the NDSL 2024.04 stack of the reference one passes a
k-index field instead, as WIP__hybrid_index_2dout.py:79-82 does.  ``remap`` / ``remap_ppm`` (a source pointer
marching under a ``while`` with variable-K reads) are left to the oracle: gtscript has no per-column scalar state
outside IJ fields, and the kernels' signatures carry none.
"""
from __future__ import annotations

from .gtscript import BACKWARD, FORWARD, PARALLEL, THIS_K, computation, function, interval  # noqa: F401
from .registry import kernel
from .typing import Float, FloatField, FloatFieldIJ, IntFieldIJ  # noqa: F401

# constants of the saturation adjustment (oracle/numpy_oracle.py SAT_*)
EPS = 0.622
LCP = 2.5e6 / 1004.0


@kernel("find_klcl")
def find_klcl(PLmb: FloatField, PLCL: FloatFieldIJ, KLCL: IntFieldIJ, PLmb_at_KLCL: FloatFieldIJ):
    """S4a: bottom-up search for the first level whose pressure is <= the LCL pressure (Fortran PLmb(i,j,KLCL(i,j)))."""
    with computation(BACKWARD):
        with interval(-1, None):
            KLCL = -1
            if PLmb <= PLCL:
                KLCL = THIS_K
                PLmb_at_KLCL = PLmb
        with interval(0, -1):
            if KLCL < 0 and PLmb <= PLCL:
                KLCL = THIS_K
                PLmb_at_KLCL = PLmb


@kernel("cloud_top")
def cloud_top(ql: FloatField, ktop: IntFieldIJ, ql_min: Float):
    """S4c: smallest k with ql > ql_min, -1 for a clear column."""
    with computation(FORWARD):
        with interval(0, 1):
            ktop = -1
            if ql > ql_min:
                ktop = THIS_K
        with interval(1, None):
            if ktop < 0 and ql > ql_min:
                ktop = THIS_K


@kernel("saturation_adjust")
def saturation_adjust(T: FloatField, q: FloatField, ql: FloatField, p: FloatField):
    """S4b: two fixed Newton steps towards saturation, condensate cannot go negative."""
    with computation(PARALLEL), interval(...):
        tm = T - 29.65
        es = 611.2 * exp(17.67 * (T - 273.15) / tm)
        den = p - (1.0 - 0.622) * es
        qs = 0.622 * es / den
        des = es * (17.67 * 243.5) / (tm * tm)
        dqs = 0.622 * p * des / (den * den)
        dq = max((q - qs) / (1.0 + (2.5e6 / 1004.0) * dqs), -ql)
        T = T + (2.5e6 / 1004.0) * dq
        q = q - dq
        ql = ql + dq
        tm = T - 29.65
        es = 611.2 * exp(17.67 * (T - 273.15) / tm)
        den = p - (1.0 - 0.622) * es
        qs = 0.622 * es / den
        des = es * (17.67 * 243.5) / (tm * tm)
        dqs = 0.622 * p * des / (den * den)
        dq = max((q - qs) / (1.0 + (2.5e6 / 1004.0) * dqs), -ql)
        T = T + (2.5e6 / 1004.0) * dq
        q = q - dq
        ql = ql + dq


@function
def ppm_flux(qm3, qm2, qm1, q0, qp1, qp2, c):
    """Unlimited-PPM flux through the interface between cells -1 and 0 (oracle/numpy_oracle.py _ppm_flux)."""
    al_m1 = (7.0 / 12.0) * (qm2 + qm1) - (1.0 / 12.0) * (qm3 + q0)
    al_0 = (7.0 / 12.0) * (qm1 + q0) - (1.0 / 12.0) * (qm2 + qp1)
    al_p1 = (7.0 / 12.0) * (q0 + qp1) - (1.0 / 12.0) * (qm1 + qp2)
    bl_m = al_m1 - qm1
    br_m = al_0 - qm1
    b0_m = bl_m + br_m
    f_pos = qm1 + (1.0 - c) * (br_m - c * b0_m)
    bl_0 = al_0 - q0
    br_0 = al_p1 - q0
    b0_0 = bl_0 + br_0
    f_neg = q0 + (1.0 + c) * (bl_0 + c * b0_0)
    return f_pos if c > 0 else f_neg


@kernel("fv_tp2d")
def fv_tp2d(q: FloatField, crx: FloatField, xfx: FloatField, cry: FloatField, yfx: FloatField, rarea: FloatFieldIJ,
            q_out: FloatField):
    """S5: flux-form PPM transport; q carries a 3-cell halo, crx / xfx sit on x-interfaces, cry / yfx on y-interfaces."""
    with computation(PARALLEL), interval(...):
        fx_lo = ppm_flux(q[-3, 0, 0], q[-2, 0, 0], q[-1, 0, 0], q, q[1, 0, 0], q[2, 0, 0], crx) * xfx
        fx_hi = ppm_flux(q[-2, 0, 0], q[-1, 0, 0], q, q[1, 0, 0], q[2, 0, 0], q[3, 0, 0], crx[1, 0, 0]) * xfx[1, 0, 0]
        fy_lo = ppm_flux(q[0, -3, 0], q[0, -2, 0], q[0, -1, 0], q, q[0, 1, 0], q[0, 2, 0], cry) * yfx
        fy_hi = ppm_flux(q[0, -2, 0], q[0, -1, 0], q, q[0, 1, 0], q[0, 2, 0], q[0, 3, 0], cry[0, 1, 0]) * yfx[0, 1, 0]
        q_out = q - rarea * ((fx_hi - fx_lo) + (fy_hi - fy_lo))


@kernel("pe_prefix")
def pe_prefix(delp: FloatField, ptop: Float, pe: FloatField):
    """S6a: interface pressures, pe[0] = ptop, pe[k] = pe[k-1] + delp[k-1]; the compute domain has nk + 1 levels."""
    with computation(FORWARD):
        with interval(0, 1):
            pe = ptop
        with interval(1, None):
            pe = pe[0, 0, -1] + delp[0, 0, -1]


@kernel("tridiag")
def tridiag(a: FloatField, b: FloatField, c: FloatField, d: FloatField, x: FloatField):
    """S6c: Thomas algorithm for a x[k-1] + b x[k] + c x[k+1] = d."""
    with computation(FORWARD):
        with interval(0, 1):
            cp = c / b
            dp = d / b
        with interval(1, None):
            m = b - a * cp[0, 0, -1]
            cp = c / m
            dp = (d - a * dp[0, 0, -1]) / m
    with computation(BACKWARD):
        with interval(-1, None):
            x = dp
        with interval(0, -1):
            x = dp - cp * x[0, 0, 1]
