"""The oracle against every golden vector the reference holds for this path, and the two
oracle implementations (NumPy restatement, C restatement) against each other."""
import numpy as np
import pytest

from oracle import inputs as gen
from oracle import numpy_oracle as orc
from oracle.c_oracle import COracle


@pytest.fixture(scope="module")
def corc():
    return COracle()


def zeros_ifirst(shape, dtype=np.float64):
    a = gen.ifirst_empty(shape, dtype)
    a[...] = 0
    return a


# ---- reference asserts -------------------------------------------------------------------


def test_golden_top_of_column():
    """Do__get_top_of_the_column.py:59-68: ones with 42 at the last level -> all 42."""
    I = gen.golden_column_input()
    O = np.zeros((3, 3, 4))
    tmp = np.zeros((3, 3))
    orc.top_of_column(I, tmp, O)
    assert np.all(O == 42)
    assert np.all(tmp == 42)


def test_golden_while_in_function():
    """Do__while_in_gt_functions.py:52-62: same input -> O[0,0,:] == [3,2,1,0]."""
    I = gen.golden_column_input()
    O = np.zeros((3, 3, 4))
    orc.while_in_function(I, O)
    assert (O[0, 0, :] == [3.0, 2.0, 1.0, 0.0]).all()
    assert (O == np.array([3.0, 2.0, 1.0, 0.0])[None, None, :]).all()
    O2 = np.zeros((3, 3, 4))
    assert orc.while_in_function_scan(I, O2) == 0
    assert np.array_equal(O, O2)


def test_golden_c_oracle(corc):
    I = gen.as_ifirst(gen.golden_column_input())
    O, tmp = zeros_ifirst((3, 3, 4)), zeros_ifirst((3, 3))
    corc.top_of_column(I, tmp, O)
    assert np.all(O == 42) and np.all(tmp == 42)
    O = zeros_ifirst((3, 3, 4))
    assert corc.while_in_function(I, O) == 0
    assert (O[0, 0, :] == [3.0, 2.0, 1.0, 0.0]).all()


def test_hybrid_by_inspection():
    """WIP__hybrid_index_2dout.py has no assert; by inspection O[i,j] == data[i,j,K[i,j]] (:79-82)."""
    data, kmask, kidx = gen.hybrid_inputs(3, 3, 4)
    O = np.zeros((3, 3))
    orc.hybrid_index_2dout(data, kmask, kidx, O)
    ii, jj = np.meshgrid(np.arange(3), np.arange(3), indexing="ij")
    assert np.array_equal(O, data[ii, jj, kidx.astype(int)])


def test_while_undefined_read_is_flagged():
    I = np.ones((2, 2, 5))
    with pytest.raises(orc.UndefinedBehaviour):
        orc.while_in_function(I, np.zeros_like(I))
    O = np.zeros_like(I)
    assert orc.while_in_function_scan(I, O) == I.size
    assert (O[0, 0] == [5, 4, 3, 2, 1]).all()


# ---- NumPy restatement vs C restatement ------------------------------------------------------

SHAPES = [(3, 3, 4), (24, 24, 72), (17, 5, 9), (1, 1, 1), (33, 2, 137)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_patterns_numpy_vs_c(corc, shape, dtype):
    ni, nj, nk = shape
    P = gen.top_of_column_inputs(ni, nj, nk, dtype)
    o1, t1 = zeros_ifirst(shape, dtype), zeros_ifirst(shape[:2], dtype)
    o2, t2 = zeros_ifirst(shape, dtype), zeros_ifirst(shape[:2], dtype)
    orc.top_of_column(P, t1, o1)
    corc.top_of_column(P, t2, o2)
    assert np.array_equal(o1, o2) and np.array_equal(t1, t2)

    W = gen.while_inputs(ni, nj, nk, dtype)
    o1, o2, o3 = (zeros_ifirst(shape, dtype) for _ in range(3))
    orc.while_in_function(W, o1)
    assert orc.while_in_function_scan(W, o3) == 0
    assert corc.while_in_function(W, o2) == 0
    assert np.array_equal(o1, o2) and np.array_equal(o1, o3)

    for miss in (0.0, 0.3):
        data, kmask, kidx = gen.hybrid_inputs(ni, nj, nk, dtype, miss_fraction=miss)
        o1, o2 = zeros_ifirst(shape[:2], dtype), zeros_ifirst(shape[:2], dtype)
        orc.hybrid_index_2dout(data, kmask, kidx, o1)
        corc.hybrid_index_2dout(data, kmask, kidx, o2)
        assert np.array_equal(o1, o2)
        assert np.all(o1[kidx < 0] == 0)


def test_hybrid_last_match_wins(corc):
    data, kmask, kidx = gen.hybrid_inputs(4, 3, 6)
    kmask[...] = 2.0  # every level matches columns that ask for 2
    kidx[...] = 2.0
    o1, o2 = zeros_ifirst((4, 3)), zeros_ifirst((4, 3))
    orc.hybrid_index_2dout(data, kmask, kidx, o1)
    corc.hybrid_index_2dout(data, kmask, kidx, o2)
    assert np.array_equal(o1, data[:, :, -1]) and np.array_equal(o1, o2)


@pytest.mark.parametrize("shape", [(3, 3, 4), (24, 24, 72), (7, 11, 13)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_moist_numpy_vs_c(corc, shape, dtype):
    ni, nj, nk = shape
    m = gen.moist_inputs(ni, nj, nk, dtype)
    idt = np.int64 if dtype == np.float64 else np.int32
    k1, k2 = zeros_ifirst(shape[:2], idt), zeros_ifirst(shape[:2], idt)
    p1, p2 = zeros_ifirst(shape[:2], dtype), zeros_ifirst(shape[:2], dtype)
    orc.find_klcl(m["p"], m["PLCL"], k1, p1)
    corc.find_klcl(m["p"], m["PLCL"], k2, p2)
    assert np.array_equal(k1, k2) and np.array_equal(p1, p2)
    # PLCL above the model top: nothing found, outputs untouched
    hi = gen.as_ifirst(np.full(shape[:2], 1.0), dtype)
    orc.find_klcl(m["p"], hi, k1, p1)
    corc.find_klcl(m["p"], hi, k2, p2)
    assert np.all(k1 == -1) and np.array_equal(k1, k2) and np.array_equal(p1, p2)

    c1, c2 = zeros_ifirst(shape[:2], idt), zeros_ifirst(shape[:2], idt)
    orc.cloud_top(m["ql"], c1)
    corc.cloud_top(m["ql"], c2)
    assert np.array_equal(c1, c2)
    clear = gen.as_ifirst(np.zeros(shape), dtype)
    orc.cloud_top(clear, c1)
    corc.cloud_top(clear, c2)
    assert np.all(c1 == -1) and np.array_equal(c1, c2)

    a = {k: gen.as_ifirst(m[k]) for k in ("T", "q", "ql")}
    b = {k: gen.as_ifirst(m[k]) for k in ("T", "q", "ql")}
    orc.saturation_adjust(a["T"], a["q"], a["ql"], m["p"])
    corc.saturation_adjust(b["T"], b["q"], b["ql"], m["p"])
    rtol = 1e-12 if dtype == np.float64 else 1e-5
    for k in a:
        scale = np.abs(a[k]).max()
        assert np.all(np.abs(a[k] - b[k]) <= rtol * np.maximum(np.abs(a[k]), scale)), k
    assert np.all(a["ql"] >= 0)


@pytest.mark.parametrize("shape", [(3, 3, 4), (24, 24, 8), (13, 7, 5), (1, 1, 2)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fv_numpy_vs_c(corc, shape, dtype):
    ni, nj, nk = shape
    f = gen.fv_inputs(ni, nj, nk, dtype)
    o1, o2 = zeros_ifirst(shape, dtype), zeros_ifirst(shape, dtype)
    orc.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], o1)
    corc.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], o2)
    assert np.array_equal(o1, o2)  # both built without FP contraction, same operation order


def test_fv_constant_field_and_zero_wind():
    ni, nj, nk = 8, 6, 3
    f = gen.fv_inputs(ni, nj, nk)
    out = zeros_ifirst((ni, nj, nk))
    # uniform q and non-divergent unit fluxes: nothing changes
    q = gen.as_ifirst(np.full((ni + 6, nj + 6, nk), 2.5))
    one_x = gen.as_ifirst(np.ones((ni + 1, nj, nk)))
    one_y = gen.as_ifirst(np.ones((ni, nj + 1, nk)))
    orc.fv_tp2d(q, f["crx"], one_x, f["cry"], one_y, f["rarea"], out)
    assert np.allclose(out, 2.5, rtol=0, atol=1e-14)
    # zero area fluxes: q_out == q
    zx, zy = one_x * 0, one_y * 0
    orc.fv_tp2d(f["q"], f["crx"], zx, f["cry"], zy, f["rarea"], out)
    assert np.array_equal(out, f["q"][3:-3, 3:-3, :])


@pytest.mark.parametrize("shape", [(3, 3, 4), (12, 9, 72), (5, 4, 137)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_vertical_numpy_vs_c(corc, shape, dtype):
    ni, nj, nk = shape
    nk2 = nk + 3 if nk > 4 else nk
    v = gen.vertical_inputs(ni, nj, nk, dtype, nk2=nk2)
    pe_a, pe_b = zeros_ifirst((ni, nj, nk + 1), dtype), zeros_ifirst((ni, nj, nk + 1), dtype)
    orc.pe_prefix(v["delp"], v["ptop"], pe_a)
    corc.pe_prefix(v["delp"], v["ptop"], pe_b)
    assert np.array_equal(pe_a, pe_b) and np.array_equal(pe_a, v["pe1"])

    q2a, q2b = zeros_ifirst((ni, nj, nk2), dtype), zeros_ifirst((ni, nj, nk2), dtype)
    orc.remap(v["pe1"], v["q1"], v["pe2"], q2a)
    corc.remap(v["pe1"], v["q1"], v["pe2"], q2b)
    assert np.array_equal(q2a, q2b)
    col = orc.remap_column(v["pe1"][1, 2], v["q1"][1, 2], v["pe2"][1, 2])
    assert np.array_equal(col, q2a[1, 2])
    # conservation of the column integral
    m1 = (v["q1"].astype(np.float64) * np.diff(v["pe1"].astype(np.float64), axis=2)).sum(axis=2)
    m2 = (q2a.astype(np.float64) * np.diff(v["pe2"].astype(np.float64), axis=2)).sum(axis=2)
    assert np.allclose(m1, m2, rtol=1e-12 if dtype == np.float64 else 2e-5)
    # identity remap
    q2c = zeros_ifirst((ni, nj, nk), dtype)
    orc.remap(v["pe1"], v["q1"], v["pe1"], q2c)
    assert np.allclose(q2c, v["q1"], rtol=1e-12 if dtype == np.float64 else 1e-5)

    t = gen.tridiag_inputs(ni, nj, nk, dtype)
    xa, xb = zeros_ifirst(shape, dtype), zeros_ifirst(shape, dtype)
    orc.tridiag(t["a"], t["b"], t["c"], t["d"], xa)
    corc.tridiag(t["a"], t["b"], t["c"], t["d"], xb)
    assert np.array_equal(xa, xb)
    # residual of the solve
    r = t["b"] * xa
    r[:, :, 1:] += t["a"][:, :, 1:] * xa[:, :, :-1]
    r[:, :, :-1] += t["c"][:, :, :-1] * xa[:, :, 1:]
    assert np.allclose(r, t["d"], atol=1e-12 if dtype == np.float64 else 1e-4)


# ---- S6d PPM remap -----------------------------------------------------------------------------


@pytest.mark.parametrize("shape", [(3, 3, 4, 4), (6, 5, 20, 23), (5, 3, 72, 60), (4, 2, 137, 137), (4, 2, 5, 9)])
@pytest.mark.parametrize("kord,iv", [(4, 1), (4, 0), (5, 1), (5, 0), (6, 1), (6, 0)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_remap_ppm_numpy_vs_c(corc, shape, kord, iv, dtype):
    """The two restatements of the PPM remap agree bit for bit (both built without FP contraction)."""
    ni, nj, nk, nk2 = shape
    for smooth in (True, False):
        v = gen.ppm_inputs(ni, nj, nk, dtype, nk2=nk2, smooth=smooth, positive=(iv == 0))
        a, b = zeros_ifirst((ni, nj, nk2), dtype), zeros_ifirst((ni, nj, nk2), dtype)
        orc.remap_ppm(v["pe1"], v["q1"], v["pe2"], a, kord, iv)
        corc.remap_ppm(v["pe1"], v["q1"], v["pe2"], b, kord, iv)
        assert np.isfinite(a).all() and np.array_equal(a, b)


@pytest.mark.parametrize("kord,iv", [(4, 1), (5, 0), (6, 1)])
def test_remap_ppm_properties(corc, kord, iv):
    ni, nj, nk, nk2 = 8, 6, 72, 80
    v = gen.ppm_inputs(ni, nj, nk, np.float64, nk2=nk2, positive=True)
    q2 = zeros_ifirst((ni, nj, nk2), np.float64)
    corc.remap_ppm(v["pe1"], v["q1"], v["pe2"], q2, kord, iv)
    dp1, dp2 = np.diff(v["pe1"], axis=2), np.diff(v["pe2"], axis=2)
    # conservative: the integral of the parabolas over the column is the integral of the means
    assert np.allclose((q2 * dp2).sum(2), (v["q1"] * dp1).sum(2), rtol=1e-13)
    # identity: remapping onto the source grid returns the means (each target layer = one whole parabola)
    same = zeros_ifirst((ni, nj, nk), np.float64)
    corc.remap_ppm(v["pe1"], v["q1"], v["pe1"], same, kord, iv)
    assert np.allclose(same, v["q1"], rtol=1e-12, atol=1e-13)
    # every parabola of the monotone scheme stays between its neighbours' means in the interior
    al, ar, a6 = orc.ppm_profile(v["q1"], dp1, kord, iv)
    assert np.allclose(a6, 3.0 * (2.0 * v["q1"] - (al + ar)), rtol=1e-9, atol=1e-12)  # what the CUDA kernel relies on
    if kord == 4:
        lo = np.minimum(np.minimum(v["q1"][:, :, :-2], v["q1"][:, :, 1:-1]), v["q1"][:, :, 2:])
        hi = np.maximum(np.maximum(v["q1"][:, :, :-2], v["q1"][:, :, 1:-1]), v["q1"][:, :, 2:])
        for edge in (al[:, :, 1:-1], ar[:, :, 1:-1]):
            assert (edge[:, :, 1:-1] >= lo[:, :, 1:-1] - 1e-12).all() and (edge[:, :, 1:-1] <= hi[:, :, 1:-1] + 1e-12).all()
    if iv == 0:
        assert (q2 >= -1e-12).all()  # positive definite in, positive definite out


def test_remap_ppm_beats_piecewise_constant_on_a_smooth_profile(corc):
    """Remap a smooth function of pressure there and back: the PPM round trip must lose far less than the
    piecewise-constant remap (S6b) does -- the reason the dycore uses it."""
    ni, nj, nk = 4, 3, 72
    v = gen.vertical_inputs(ni, nj, nk, np.float64, nk2=nk)
    pe1, pe2 = v["pe1"], v["pe2"]
    # f(p) = 1 + 0.5 sin(3 pi p / ps), F = its integral
    F = lambda p: p - 0.5 * pe1[:, :, -1:] / (3.0 * np.pi) * np.cos(3.0 * np.pi * p / pe1[:, :, -1:])  # noqa: E731
    q1 = gen.as_ifirst(np.diff(F(pe1), axis=2) / np.diff(pe1, axis=2))  # exact layer means
    exact2 = np.diff(F(pe2), axis=2) / np.diff(pe2, axis=2)
    ppm, pcm = zeros_ifirst((ni, nj, nk), np.float64), zeros_ifirst((ni, nj, nk), np.float64)
    corc.remap_ppm(pe1, q1, pe2, ppm, 4, 1)
    corc.remap(pe1, q1, pe2, pcm)
    e_ppm, e_pcm = np.abs(ppm - exact2)[:, :, 3:-3].max(), np.abs(pcm - exact2)[:, :, 3:-3].max()
    assert e_ppm < 0.2 * e_pcm, (e_ppm, e_pcm)  # measured: 1.4e-3 vs 1.5e-2 on layers of random thickness


# ---- S5b fv_tp2d_split: FV3 fv_tp_2d inner/outer splitting ---------------------------------------------


@pytest.mark.parametrize("shape", [(12, 9, 3), (3, 3, 4), (33, 17, 2), (1, 1, 1)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fv_tp2d_split_numpy_vs_c(corc, shape, dtype):
    ni, nj, nk = shape
    f = gen.fv_split_inputs(ni, nj, nk, dtype)
    a, fa, ga = zeros_ifirst(shape, dtype), zeros_ifirst((ni + 1, nj, nk), dtype), zeros_ifirst((ni, nj + 1, nk), dtype)
    b, fb, gb = zeros_ifirst(shape, dtype), zeros_ifirst((ni + 1, nj, nk), dtype), zeros_ifirst((ni, nj + 1, nk), dtype)
    orc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], a, fa, ga)
    corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], b, fb, gb)
    assert np.array_equal(a, b) and np.array_equal(fa, fb) and np.array_equal(ga, gb)
    c = zeros_ifirst(shape, dtype)
    corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], c)  # fluxes not requested
    assert np.array_equal(b, c)
    # flux form: the update is exactly the divergence of the returned fluxes
    upd = f["q"][3:-3, 3:-3] + f["rarea"][:, :, None] * ((fa[:-1] - fa[1:]) + (ga[:, :-1] - ga[:, 1:]))
    assert np.array_equal(upd.astype(dtype), a)


def test_fv_tp2d_split_properties(corc):
    ni, nj, nk = 20, 14, 3
    f = gen.fv_split_inputs(ni, nj, nk)
    z = np.zeros_like
    # a constant field stays constant under non-divergent (here: zero) flow, and takes exactly the flow divergence otherwise
    out = zeros_ifirst((ni, nj, nk), np.float64)
    corc.fv_tp2d_split(np.full_like(f["q"], 2.5), z(f["crx"]), z(f["xfx"]), z(f["cry"]), z(f["yfx"]), f["area"], f["rarea"], out)
    assert np.array_equal(out, np.full_like(out, 2.5))
    corc.fv_tp2d_split(np.full_like(f["q"], 2.5), f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], out)
    xf, yf = f["xfx"][:, 3:-3], f["yfx"][3:-3]
    want = 2.5 + f["rarea"][:, :, None] * 2.5 * ((xf[:-1] - xf[1:]) + (yf[:, :-1] - yf[:, 1:]))
    assert np.allclose(out, want, rtol=1e-13)
    # with no flow in y the splitting collapses onto the 1-D operator of S5 (q_i = q, fx = fx2): same update as fv_tp2d
    s5 = zeros_ifirst((ni, nj, nk), np.float64)
    corc.fv_tp2d(f["q"], gen.as_ifirst(f["crx"][:, 3:-3]), gen.as_ifirst(f["xfx"][:, 3:-3]), gen.as_ifirst(z(f["cry"])[3:-3]),
                 gen.as_ifirst(z(f["yfx"])[3:-3]), f["rarea"], s5)
    corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], z(f["cry"]), z(f["yfx"]), f["area"], f["rarea"], out)
    assert np.allclose(out, s5, rtol=1e-12, atol=1e-13)
    # the cross terms are what the splitting adds: with flow in both directions the two operators must differ
    corc.fv_tp2d(f["q"], gen.as_ifirst(f["crx"][:, 3:-3]), gen.as_ifirst(f["xfx"][:, 3:-3]), gen.as_ifirst(f["cry"][3:-3]),
                 gen.as_ifirst(f["yfx"][3:-3]), f["rarea"], s5)
    corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], out)
    assert np.abs(out - s5).max() > 1e-3


@pytest.mark.parametrize("flags", [1, 2, 4, 8, 15, 6])
def test_fv_tp2d_split_cube_corners(corc, flags):
    """copy_corners: the y-sweep of a sub-domain at a cube corner sees direction-2 corner values."""
    ni, nj, nk = 10, 7, 2
    f = gen.fv_split_inputs(ni, nj, nk)
    q1 = f["q"].copy()
    orc.copy_corners(q1, 1, flags)  # what the halo update leaves in q
    a, b = zeros_ifirst((ni, nj, nk), np.float64), zeros_ifirst((ni, nj, nk), np.float64)
    q1 = gen.as_ifirst(q1)
    orc.fv_tp2d_split(q1, f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], a, corner_flags=flags)
    corc.fv_tp2d_split(q1, f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], b, corner_flags=flags)
    assert np.array_equal(a, b)
    plain = zeros_ifirst((ni, nj, nk), np.float64)
    orc.fv_tp2d_split(q1, f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], plain)
    diff = np.argwhere(np.abs(plain - a).max(axis=2) > 0)
    assert len(diff) > 0  # the corner rule matters ...
    for i, j in diff:  # ... and only within reach (3 cells in i, via q_i) of a flagged corner
        near = [(i < 3 and j < 6, 1), (i >= ni - 3 and j < 6, 2), (i < 3 and j >= nj - 6, 4), (i >= ni - 3 and j >= nj - 6, 8)]
        assert any(c and (flags & bit) for c, bit in near), (i, j)
    # direction-1 and direction-2 fills are each other's transposes about the corner diagonal
    qx, qy = np.array(f["q"]), np.array(f["q"])
    orc.copy_corners(qx, 1, 1)
    orc.copy_corners(qy, 2, 1)
    assert np.array_equal(qy[2, 2], f["q"][3, 2]) and np.array_equal(qx[2, 2], f["q"][2, 3])  # cell (-1,-1) <- (0,-1) | (-1,0)
