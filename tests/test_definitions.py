"""b200stencil.definitions: the gtscript statements of S4 - S6 (stencils without source in the reference).

Each definition is (1) resolvable to its hand-written kernel through the NDSL-shaped factory, like a pattern's
`stencil`, and (2) an executable specification: run through tests/golden/gtscript_interp.py (gt4py numpy-backend
semantics) it must reproduce the oracle -- bit for bit where the arithmetic has one possible order (index outputs,
moves, sums in sequence, the fp64 transport and Thomas solves written in the oracle's operation order), to a few ulp
where a library `exp` is involved or fp32 constants are folded differently."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from gtscript_interp import GtscriptProgram, UndefinedRead  # noqa: E402

from b200stencil import definitions, registry  # noqa: E402
from oracle import inputs as gen  # noqa: E402
from oracle import numpy_oracle as orc  # noqa: E402

NAMES = ["find_klcl", "cloud_top", "saturation_adjust", "fv_tp2d", "pe_prefix", "tridiag"]


def prog(name):
    return GtscriptProgram.from_function(getattr(definitions, name))


@pytest.mark.parametrize("name", NAMES)
def test_definitions_resolve_to_their_kernels(name):
    assert registry.resolve(getattr(definitions, name)) == name
    from b200stencil import get_factories_single_tile_numpy
    from b200stencil.constants import X_DIM, Y_DIM, Z_DIM

    sf, _ = get_factories_single_tile_numpy(4, 3, 5, 0)
    st = sf.from_dims_halo(func=getattr(definitions, name), compute_dims=[X_DIM, Y_DIM, Z_DIM])
    assert st.kernel_name == name


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_moist_definitions_match_the_oracle(dtype):
    ni, nj, nk = 5, 4, 9
    m = gen.moist_inputs(ni, nj, nk, dtype)
    it = np.int64 if dtype == np.float64 else np.int32
    # find_klcl, incl. columns where no level qualifies (KLCL = -1, PLmb_at_KLCL untouched)
    plcl = m["PLCL"].copy()
    plcl[0, 0] = 1.0
    k_ref, p_ref = np.zeros((ni, nj), it), np.full((ni, nj), 7, dtype)
    orc.find_klcl(m["p"], plcl, k_ref, p_ref)
    k_got, p_got = np.zeros((ni, nj), it), np.full((ni, nj), 7, dtype)
    prog("find_klcl")(m["p"], plcl, k_got, p_got)
    assert np.array_equal(k_got, k_ref) and np.array_equal(p_got, p_ref) and k_ref[0, 0] == -1
    # cloud_top, incl. a clear column
    ql = m["ql"].copy()
    ql[1, 2, :] = 0
    c_ref, c_got = np.zeros((ni, nj), it), np.zeros((ni, nj), it)
    orc.cloud_top(ql, c_ref)
    prog("cloud_top")(ql, c_got, dtype(orc.CLOUD_QL_MIN))
    assert np.array_equal(c_got, c_ref) and c_ref[1, 2] == -1
    # saturation_adjust
    ref = {k: m[k].copy() for k in ("T", "q", "ql")}
    orc.saturation_adjust(ref["T"], ref["q"], ref["ql"], m["p"])
    got = {k: m[k].copy() for k in ("T", "q", "ql")}
    prog("saturation_adjust")(got["T"], got["q"], got["ql"], m["p"])
    tol = 1e-14 if dtype == np.float64 else 2e-6
    for k in ref:
        scale = np.abs(ref[k]).max()
        assert np.abs(got[k].astype(np.float64) - ref[k]).max() <= tol * scale, k


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fv_tp2d_definition_matches_the_oracle(dtype):
    ni, nj, nk = 6, 5, 2
    f = gen.fv_inputs(ni, nj, nk, dtype)
    ref = np.zeros((ni, nj, nk), dtype)
    orc.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], ref)
    got = np.zeros((ni, nj, nk), dtype)
    p = prog("fv_tp2d")
    p(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], got, domain=(ni, nj, nk), origins={"q": (3, 3, 0)})
    if dtype == np.float64:
        assert np.array_equal(got, ref)  # same operations in the same order
    else:
        assert np.abs(got.astype(np.float64) - ref).max() <= 2e-6 * np.abs(ref).max()
    # the reach is exactly the 3-cell halo: with a 2-cell halo the definition reads outside its storage
    with pytest.raises(UndefinedRead):
        p(np.ascontiguousarray(f["q"][1:-1, 1:-1]), f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], got, domain=(ni, nj, nk),
          origins={"q": (2, 2, 0)})


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_vertical_definitions_match_the_oracle(dtype):
    ni, nj, nk = 4, 3, 10
    v = gen.vertical_inputs(ni, nj, nk, dtype)
    ref = np.zeros((ni, nj, nk + 1), dtype)
    orc.pe_prefix(v["delp"], float(v["ptop"]), ref)
    got = np.zeros((ni, nj, nk + 1), dtype)
    prog("pe_prefix")(v["delp"], dtype(v["ptop"]), got, domain=(ni, nj, nk + 1))
    assert np.array_equal(got, ref)  # one possible order of additions
    t = gen.tridiag_inputs(ni, nj, nk, dtype)
    ref = np.zeros((ni, nj, nk), dtype)
    orc.tridiag(t["a"], t["b"], t["c"], t["d"], ref)
    got = np.zeros((ni, nj, nk), dtype)
    prog("tridiag")(t["a"], t["b"], t["c"], t["d"], got)
    assert np.array_equal(got, ref)
    resid = t["b"] * got
    resid[:, :, 1:] += t["a"][:, :, 1:] * got[:, :, :-1]
    resid[:, :, :-1] += t["c"][:, :, :-1] * got[:, :, 1:]
    assert np.abs(resid - t["d"]).max() <= (1e-13 if dtype == np.float64 else 1e-5)
