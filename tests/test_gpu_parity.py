"""-m gpu parity tests: every CUDA stencil, called through the C-ABI, against the oracle on the
same seeded inputs.  Bit-exact for integer / index outputs and pure-move stencils, 1e-12 (fp64) /
1e-5 (fp32) relative for floating-point fields (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import inputs as gen
from oracle import numpy_oracle as orc
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

from gpu_util import assert_close, down, up, up_batch, zeros_like_np  # noqa: E402

RTOL = {np.float64: 1e-12, np.float32: 1e-5}
DTYPES = [np.float64, np.float32]
# (ni, nj, nk): golden size, cfg1 size, ragged sizes (scalar path), vector-friendly, tall
SHAPES = [(3, 3, 4), (24, 24, 72), (17, 5, 9), (1, 1, 1), (33, 2, 137), (96, 96, 72), (64, 3, 5)]


@pytest.fixture(scope="module")
def st():
    from b200stencil import stencils

    return stencils


@pytest.fixture(scope="module")
def corc():
    return COracle()


def tdt(dtype):
    return torch.float64 if dtype == np.float64 else torch.float32


def idt(dtype):
    return np.int64 if dtype == np.float64 else np.int32


# ---- the reference's own golden vectors, on the GPU ------------------------------------------------


def test_golden_top_of_column(st):
    I = up(gen.golden_column_input())
    O = up(np.zeros((3, 3, 4)))
    tmp = up(np.zeros((3, 3)))
    st.top_of_column(I, tmp, O)
    assert np.all(down(O) == 42) and np.all(down(tmp) == 42)


def test_golden_while_in_function(st):
    I = up(gen.golden_column_input())
    O = up(np.zeros((3, 3, 4)))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    st.while_in_function(I, O, undefined_count=cnt)
    assert (down(O)[0, 0, :] == [3.0, 2.0, 1.0, 0.0]).all()
    assert int(cnt.item()) == 0


# ---- S1-S3 ------------------------------------------------------------------------------------------


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("align", [True, False])
def test_patterns(st, shape, dtype, align):
    ni, nj, nk = shape
    P = gen.top_of_column_inputs(ni, nj, nk, dtype)
    o, t = zeros_like_np(shape, dtype), zeros_like_np(shape[:2], dtype)
    orc.top_of_column(P, t, o)
    dO, dT = up(np.zeros(shape, dtype), align), up(np.zeros(shape[:2], dtype), align)
    st.top_of_column(up(P, align), dT, dO)
    assert np.array_equal(down(dO), o) and np.array_equal(down(dT), t)

    W = gen.while_inputs(ni, nj, nk, dtype)
    o = zeros_like_np(shape, dtype)
    assert orc.while_in_function_scan(W, o) == 0
    dO = up(np.zeros(shape, dtype), align)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    st.while_in_function(up(W, align), dO, undefined_count=cnt)
    assert np.array_equal(down(dO), o) and int(cnt.item()) == 0

    for miss in (0.0, 0.3):
        data, kmask, kidx = gen.hybrid_inputs(ni, nj, nk, dtype, miss_fraction=miss)
        o = zeros_like_np(shape[:2], dtype)
        orc.hybrid_index_2dout(data, kmask, kidx, o)
        dO = up(np.zeros(shape[:2], dtype), align)
        st.hybrid_index_2dout(up(data, align), up(kmask, align), up(kidx, align), dO)
        assert np.array_equal(down(dO), o)


def test_while_literal_form_matches(st):
    """The literal per-point while (not the scan restatement) is what the GPU must reproduce."""
    ni, nj, nk = 12, 7, 19
    W = gen.while_inputs(ni, nj, nk)
    o = zeros_like_np((ni, nj, nk), np.float64)
    orc.while_in_function(W, o)
    dO = up(np.zeros((ni, nj, nk)))
    st.while_in_function(up(W), dO)
    assert np.array_equal(down(dO), o)


def test_while_undefined_is_counted(st):
    I = np.ones((4, 2, 5))
    I[0, 0, 4] = 7.0  # one defined column
    O = up(np.zeros_like(I))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    st.while_in_function(up(I), O, undefined_count=cnt)
    ref = np.zeros_like(I)
    n = orc.while_in_function_scan(I, ref)
    assert int(cnt.item()) == n == 7 * 5
    assert np.array_equal(down(O), ref)
    st.while_in_function(up(I), O)  # NULL counter is allowed


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("shape", [(3, 3, 4), (33, 5, 72), (64, 4, 137), (17, 3, 9), (8, 2, 10), (6, 2, 300)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_while_variants(st, variant, shape, dtype):
    """Column scan and k-split kernels of while_in_function (k_patterns.cu) against the oracle, with sparse
    hits (whole segments without one, so the carry crosses several segments) and undefined points."""
    from b200stencil import _abi

    ni, nj, nk = shape
    rng = np.random.default_rng(7)
    I = rng.uniform(0.0, 3.999, shape).astype(dtype)
    hits = rng.integers(0, nk, size=(ni, nj, 2))
    for a in range(ni):
        for b in range(nj):
            if (a + b) % 5:  # every fifth column has no hit at all: undefined everywhere
                I[a, b, hits[a, b]] = 4.0 + rng.random(2) * 30
    I = gen.as_ifirst(I)
    ref = zeros_like_np(shape, dtype)
    n = orc.while_in_function_scan(I, ref)
    assert n > 0
    _abi.set_option("while_variant", variant)
    try:
        for align in (True, False):
            O = up(np.zeros(shape, dtype), align)
            cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            st.while_in_function(up(I, align), O, undefined_count=cnt)
            assert np.array_equal(down(O), ref) and int(cnt.item()) == n
    finally:
        _abi.set_option("while_variant", 0)


def test_hybrid_arbitrary_mask_last_match_wins(st):
    data, kmask, kidx = gen.hybrid_inputs(8, 6, 10)
    rng = np.random.default_rng(5)
    kmask = gen.as_ifirst(rng.integers(0, 4, size=kmask.shape).astype(np.float64))
    kidx = gen.as_ifirst(rng.integers(0, 5, size=kidx.shape).astype(np.float64))
    o = zeros_like_np((8, 6), np.float64)
    o[...] = -5.0  # prior contents survive where nothing matches
    dO = up(o.copy())
    orc.hybrid_index_2dout(data, kmask, kidx, o)
    st.hybrid_index_2dout(up(data), up(kmask), up(kidx), dO)
    assert np.array_equal(down(dO), o)
    assert (o == -5.0).any()


@pytest.mark.parametrize("dtype", DTYPES)
def test_patterns_batched_tiles(st, dtype):
    """6 tiles in one launch (cfg2 shape class), including halo-padded storage via slicing."""
    ni, nj, nk, nb = 12, 10, 9, 6
    Ws = [gen.while_inputs(ni, nj, nk, dtype, cfg=20 + b) for b in range(nb)]
    dW = up_batch(Ws)
    store = torch.zeros((nb, nk, nj + 6, ni + 6), dtype=tdt(dtype), device="cuda")
    dO = store.permute(0, 3, 2, 1)[:, 3:-3, 3:-3, :]  # compute window of a halo-3 field
    st.while_in_function(dW, dO)
    for b in range(nb):
        o = zeros_like_np((ni, nj, nk), dtype)
        orc.while_in_function_scan(Ws[b], o)
        assert np.array_equal(down(dO[b]), o)
    assert float(store.sum()) == float(dO.sum())  # halo cells untouched (writes stay in the domain)


# ---- S4 ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("shape", [(3, 3, 4), (24, 24, 72), (7, 11, 13), (64, 8, 72)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_moist(st, shape, dtype):
    ni, nj, nk = shape
    m = gen.moist_inputs(ni, nj, nk, dtype)
    k1, p1 = zeros_like_np(shape[:2], idt(dtype)), zeros_like_np(shape[:2], dtype)
    p1[...] = -1.0
    dK = up(np.zeros(shape[:2], idt(dtype)))
    dP = up(p1.copy())
    orc.find_klcl(m["p"], m["PLCL"], k1, p1)
    st.find_klcl(up(m["p"]), up(m["PLCL"]), dK, dP)
    assert np.array_equal(down(dK), k1) and np.array_equal(down(dP), p1)
    # nothing found: KLCL = -1, PLmb_at_KLCL untouched
    hi = np.full(shape[:2], 1.0, dtype)
    st.find_klcl(up(m["p"]), up(hi), dK, dP)
    assert np.all(down(dK) == -1) and np.array_equal(down(dP), p1)

    c1 = zeros_like_np(shape[:2], idt(dtype))
    dC = up(np.zeros(shape[:2], idt(dtype)))
    orc.cloud_top(m["ql"], c1)
    st.cloud_top(up(m["ql"]), dC)
    assert np.array_equal(down(dC), c1)
    st.cloud_top(up(np.zeros(shape, dtype)), dC)
    assert np.all(down(dC) == -1)

    a = {k: gen.as_ifirst(m[k]) for k in ("T", "q", "ql")}
    d = {k: up(m[k]) for k in ("T", "q", "ql")}
    orc.saturation_adjust(a["T"], a["q"], a["ql"], m["p"])
    st.saturation_adjust(d["T"], d["q"], d["ql"], up(m["p"]))
    for k in a:
        assert_close(down(d[k]), a[k], RTOL[dtype], k)


# ---- S5 ------------------------------------------------------------------------------------------------


def _fv_case(st, corc, ni, nj, nk, dtype, variant=0, region=None, align=True):
    from b200stencil import _abi

    f = gen.fv_inputs(ni, nj, nk, dtype)
    ref = zeros_like_np((ni, nj, nk), dtype)
    corc.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], ref)
    d = {k: up(v, align) for k, v in f.items()}
    out = up(np.full((ni, nj, nk), -7.0, dtype), align)
    _abi.set_option("fv_variant", variant)
    try:
        st.fv_tp2d(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out, region=region)
    finally:
        _abi.set_option("fv_variant", 0)
    got = down(out)
    if region is None:
        assert_close(got, ref, RTOL[dtype], "q_out")
    else:
        i0, i1, j0, j1 = region
        assert_close(got[i0:i1, j0:j1], ref[i0:i1, j0:j1], RTOL[dtype], "q_out[region]")
        mask = np.ones((ni, nj), bool)
        mask[i0:i1, j0:j1] = False
        assert np.all(got[mask] == -7.0)  # nothing written outside the rectangle


@pytest.mark.parametrize("shape", [(3, 3, 4), (24, 24, 8), (13, 7, 5), (1, 1, 2), (130, 20, 3), (128, 64, 2), (300, 9, 2)])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("variant", [0, 1, 3])
def test_fv_tp2d(st, corc, shape, dtype, variant):
    _fv_case(st, corc, *shape, dtype, variant=variant)


@pytest.mark.parametrize("dtype", DTYPES)
def test_fv_tp2d_variants_bit_identical(st, dtype):
    """Direct kernel, every TMA tile geometry: same bits (explicit-rounding arithmetic, csrc/fv_math.cuh)."""
    from b200stencil import _abi

    ni, nj, nk = 150, 37, 3
    f = gen.fv_inputs(ni, nj, nk, dtype)
    d = {k: up(v) for k, v in f.items()}
    outs = []
    combos = [(1, 0, 0, 0)] + [(2, ti, r, st) for ti in (32, 64, 96, 128, 192) for r, st in ((4, 2), (8, 3))]
    # streaming kernel (variant 3): strip widths, ring depths, rows per item (5 and 16 put item boundaries inside the domain)
    combos += [(3, ti, jb, st) for ti in (32, 64, 96, 128) for jb, st in ((0, 2), (5, 3), (16, 4))]
    for variant, ti, rows, stages in combos:
        for name, v in (("fv_variant", variant), ("fv_ti", ti), ("fv_rows", rows if variant == 2 else 0),
                        ("fv_jb", rows if variant == 3 else 0), ("fv_stages", stages)):
            _abi.set_option(name, v)
        try:
            out = up(np.zeros((ni, nj, nk), dtype))
            st.fv_tp2d(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
            outs.append(down(out))
        finally:
            for name in ("fv_variant", "fv_ti", "fv_rows", "fv_jb", "fv_stages"):
                _abi.set_option(name, 0)
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])


def test_fv_tp2d_streaming_regions(st, corc):
    _fv_case(st, corc, 40, 30, 3, np.float64, 3, region=(3, 37, 3, 27))
    _fv_case(st, corc, 40, 30, 3, np.float64, 3, region=(0, 40, 0, 3))
    _fv_case(st, corc, 40, 30, 3, np.float32, 3, region=(37, 40, 3, 27))
    _fv_case(st, corc, 200, 150, 2, np.float64, 3)  # several strips, two row blocks


@pytest.mark.parametrize("variant", [0, 1])
def test_fv_tp2d_regions_and_unaligned(st, corc, variant):
    _fv_case(st, corc, 40, 30, 3, np.float64, variant, region=(3, 37, 3, 27))
    _fv_case(st, corc, 40, 30, 3, np.float64, variant, region=(0, 40, 0, 3))
    _fv_case(st, corc, 40, 30, 3, np.float64, variant, region=(37, 40, 3, 27))
    _fv_case(st, corc, 40, 30, 3, np.float32, variant, region=(5, 5, 0, 30))  # empty
    _fv_case(st, corc, 37, 11, 3, np.float64, variant, align=False)  # odd strides: general path


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shift", [1, 2, 3])
def test_fv_tp2d_misaligned_base(st, corc, dtype, shift):
    """Fields whose first element is not 16-byte aligned (views into wider storage) but whose strides
    are: the TMA path must shift its boxes, not fault."""
    ni, nj, nk = 70, 9, 3
    f = gen.fv_inputs(ni, nj, nk, dtype)
    ref = zeros_like_np((ni, nj, nk), dtype)
    corc.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], ref)

    def shifted(a):
        wide = up(np.zeros((a.shape[0] + 8,) + a.shape[1:], dtype))
        view = wide[shift : shift + a.shape[0]]
        view.copy_(torch.from_numpy(np.ascontiguousarray(a)))
        return view

    d = {k: shifted(v) for k, v in f.items()}
    out = shifted(np.zeros((ni, nj, nk), dtype))
    st.fv_tp2d(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
    assert_close(down(out), ref, RTOL[dtype], "q_out")


@pytest.mark.parametrize("dtype", DTYPES)
def test_fv_tp2d_batched(st, corc, dtype):
    ni, nj, nk, nb = 20, 12, 4, 3
    fs = [gen.fv_inputs(ni, nj, nk, dtype, cfg=40 + b) for b in range(nb)]
    d = {k: up_batch([f[k] for f in fs]) for k in fs[0]}
    out = up_batch([np.zeros((ni, nj, nk), dtype)] * nb)
    st.fv_tp2d(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
    for b in range(nb):
        ref = zeros_like_np((ni, nj, nk), dtype)
        f = fs[b]
        corc.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], ref)
        assert_close(down(out[b]), ref, RTOL[dtype], f"q_out[{b}]")


# ---- S5b fv_tp2d_split -------------------------------------------------------------------


def _split_call(st, f, nk, dtype, fluxes=True, ti=0, batch=None, variant=0, jb=0):
    from b200stencil import _abi

    ni, nj = (f["rarea"][0] if batch else f["rarea"]).shape[-2:]
    d = {k: (up_batch(v) if batch else up(v)) for k, v in f.items()}
    shp = lambda s: up_batch([np.zeros(s, dtype)] * batch) if batch else up(np.zeros(s, dtype))  # noqa: E731
    out = shp((ni, nj, nk))
    fx = shp((ni + 1, nj, nk)) if fluxes else None
    fy = shp((ni, nj + 1, nk)) if fluxes else None
    for name, v in (("fv_split_ti", ti), ("fv_split_variant", variant), ("fv_split_jb", jb)):
        _abi.set_option(name, v)
    try:
        st.fv_tp2d_split(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["area"], d["rarea"], out, fx, fy)
    finally:
        for name in ("fv_split_ti", "fv_split_variant", "fv_split_jb"):
            _abi.set_option(name, 0)
    return out, fx, fy


@pytest.mark.parametrize("shape", [(3, 3, 4), (12, 9, 3), (33, 17, 2), (130, 9, 2), (64, 20, 3), (257, 5, 1), (1, 1, 1)])
@pytest.mark.parametrize("ti", [0, 56, 120])
@pytest.mark.parametrize("variant,jb", [(1, 0), (2, 0), (2, 5)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_fv_tp2d_split(st, corc, shape, ti, variant, jb, dtype):
    """Tile kernel (variant 1) and streaming kernel (variant 2; jb = rows per CTA, 5 puts block boundaries inside
    the small test domains) against the oracle; the two kernels use the same arithmetic -> identical bits."""
    ni, nj, nk = shape
    f = gen.fv_split_inputs(ni, nj, nk, dtype)
    ref, rfx, rfy = zeros_like_np(shape, dtype), zeros_like_np((ni + 1, nj, nk), dtype), zeros_like_np((ni, nj + 1, nk), dtype)
    corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], ref, rfx, rfy)
    out, fx, fy = _split_call(st, f, nk, dtype, True, ti, variant=variant, jb=jb)
    assert_close(down(out), ref, RTOL[dtype], "q_out")
    assert_close(down(fx), rfx, RTOL[dtype], "fx")
    assert_close(down(fy), rfy, RTOL[dtype], "fy")
    out2, _, _ = _split_call(st, f, nk, dtype, False, ti, variant=variant, jb=jb)
    assert torch.equal(out, out2)  # the flux outputs are optional and do not change the update
    if variant == 2:
        tile, tfx, tfy = _split_call(st, f, nk, dtype, True, ti, variant=1)
        assert torch.equal(out, tile) and torch.equal(fx, tfx) and torch.equal(fy, tfy)


@pytest.mark.parametrize("dtype", DTYPES)
def test_fv_tp2d_split_rows_per_barrier_pair(st, dtype):
    """Streaming kernel, one or two rows between barriers (fv_split_rp): same bits."""
    from b200stencil import _abi

    f = gen.fv_split_inputs(70, 23, 2, dtype)
    outs = []
    for rp in (1, 2):
        _abi.set_option("fv_split_rp", rp)
        try:
            outs.append(_split_call(st, f, 2, dtype, False, 0, variant=2, jb=7)[0])
        finally:
            _abi.set_option("fv_split_rp", 0)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("shape", [(12, 9, 3), (70, 11, 2), (130, 20, 1)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_fv_tp2d_split_cube_corners(st, corc, shape, dtype):
    """Sub-domains at cube corners: q holds the copy_corners direction-1 values (as after a halo update); with the
    flags set, the kernel's inner y-sweep must use the direction-2 values like the oracle."""
    ni, nj, nk = shape
    nb = 4
    all_flags = [1, 10, 15, 0]
    fs, refs = [], []
    for b, flags in enumerate(all_flags):
        f = gen.fv_split_inputs(ni, nj, nk, dtype, cfg=8 + b)
        q1 = np.array(f["q"])
        orc.copy_corners(q1, 1, flags)
        f["q"] = gen.as_ifirst(q1)
        r = zeros_like_np((ni, nj, nk), dtype)
        corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], r, corner_flags=flags)
        fs.append(f)
        refs.append(r)
    d = {k: up_batch([f[k] for f in fs]) for k in fs[0]}
    out = up_batch([np.zeros((ni, nj, nk), dtype)] * nb)
    from b200stencil import _abi

    cf = torch.tensor(all_flags, dtype=torch.int32, device="cuda")
    for variant, jb in ((1, 0), (2, 0), (2, 7)):
        _abi.set_option("fv_split_variant", variant)
        _abi.set_option("fv_split_jb", jb)
        try:
            out.zero_()
            st.fv_tp2d_split(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["area"], d["rarea"], out, corner_flags=cf)
        finally:
            _abi.set_option("fv_split_variant", 0)
            _abi.set_option("fv_split_jb", 0)
        for b in range(nb):
            assert_close(down(out[b]), refs[b], RTOL[dtype], f"variant {variant} sub-domain {b} flags {all_flags[b]}")
    plain = up_batch([np.zeros((ni, nj, nk), dtype)] * nb)
    st.fv_tp2d_split(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["area"], d["rarea"], plain)
    assert not torch.equal(plain[0], out[0]) and torch.equal(plain[3], out[3])  # the flags matter, and only where set


@pytest.mark.parametrize("dtype", DTYPES)
def test_fv_tp2d_split_batched_and_unaligned(st, corc, dtype):
    """A batch of sub-domains in one launch, and fields that are interior windows of larger storage (their halo
    origins start off a 16-byte boundary: every TMA box starts at the aligned column before)."""
    from b200stencil import fields

    ni, nj, nk, nb = 70, 11, 2, 3
    fs = [gen.fv_split_inputs(ni, nj, nk, dtype, cfg=8 + b) for b in range(nb)]
    refs = []
    for f in fs:
        r = zeros_like_np((ni, nj, nk), dtype)
        corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], r)
        refs.append(r)
    out, _, _ = _split_call(st, {k: [f[k] for f in fs] for k in fs[0]}, nk, dtype, False, 0, batch=nb)
    for b in range(nb):
        assert_close(down(out[b]), refs[b], RTOL[dtype], f"batch {b}")

    def window(a, off):
        big = fields.zeros(tuple(n + 2 * off for n in a.shape[:2]) + a.shape[2:], dtype=tdt(dtype))
        win = big[off:off + a.shape[0], off:off + a.shape[1]]
        win.copy_(torch.from_numpy(np.ascontiguousarray(a)))
        return win

    f = fs[0]
    d = {k: window(f[k], 1 if k in ("q", "cry") else 3) for k in f}
    assert d["q"].data_ptr() % 16 != 0
    o = fields.zeros((ni, nj, nk), dtype=tdt(dtype))
    st.fv_tp2d_split(d["q"], d["crx"], d["xfx"], d["cry"], d["yfx"], d["area"], d["rarea"], o)
    assert_close(down(o), refs[0], RTOL[dtype], "unaligned windows")


# ---- S6 ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("shape", [(3, 3, 4), (12, 9, 72), (5, 4, 137), (64, 40, 20)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_vertical(st, shape, dtype):
    ni, nj, nk = shape
    nk2 = nk + 3 if nk > 4 else nk
    v = gen.vertical_inputs(ni, nj, nk, dtype, nk2=nk2)
    pe = up(np.zeros((ni, nj, nk + 1), dtype))
    st.pe_prefix(up(v["delp"]), v["ptop"], pe)
    assert np.array_equal(down(pe), v["pe1"])  # sequential-in-k adds, no contraction: bit-exact

    ref = zeros_like_np((ni, nj, nk2), dtype)
    orc.remap(v["pe1"], v["q1"], v["pe2"], ref)
    q2 = up(np.zeros((ni, nj, nk2), dtype))
    st.remap(up(v["pe1"]), up(v["q1"]), up(v["pe2"]), q2)
    assert_close(down(q2), ref, RTOL[dtype], "q2")
    assert np.array_equal(down(q2), ref), "remap is compiled without FMA contraction: expected bit-exact"

    # pe_prefix fused into the remap: same bits as the two kernels (and as the oracle)
    q2f = up(np.zeros((ni, nj, nk2), dtype))
    st.remap_delp(up(v["delp"]), v["ptop"], up(v["q1"]), up(v["pe2"]), q2f)
    assert np.array_equal(down(q2f), ref)

    t = gen.tridiag_inputs(ni, nj, nk, dtype)
    ref = zeros_like_np(shape, dtype)
    orc.tridiag(t["a"], t["b"], t["c"], t["d"], ref)
    x = up(np.zeros(shape, dtype))
    st.tridiag(up(t["a"]), up(t["b"]), up(t["c"]), up(t["d"]), x)
    assert_close(down(x), ref, RTOL[dtype], "x")


# (remap_variant, remap_nw, remap_cg): kernel, warps per column group, 32-column groups per CTA
REMAP_VARIANTS = {
    (1, 0, 0): "nested (thread per column)", (2, 0, 0): "slab + cp.async", (3, 0, 0): "slab + TMA, automatic geometry",
    (3, 8, 1): "slab + TMA 8x1", (3, 16, 1): "slab + TMA 16x1", (3, 8, 2): "slab + TMA 8x2",
}  # fmt: skip


@pytest.fixture
def remap_variant():
    """Force one remap kernel for a test and restore the automatic choice afterwards."""
    from b200stencil import _abi

    def force(v):
        v = (v, 0, 0) if isinstance(v, int) else v
        for name, x in zip(("remap_variant", "remap_nw", "remap_cg"), v):
            _abi.set_option(name, x)

    yield force
    force(0)


def _degenerate_vertical(ni, nj, nk, nk2, dtype):
    """Edge cases of the marching pointer: zero-thickness source layers, target edges that coincide with
    source edges, and a target column that starts above and ends below the source column."""
    v = gen.vertical_inputs(ni, nj, nk, dtype, nk2=nk2)
    delp = np.array(v["delp"])
    delp[:, :, 2::5] = 0  # zero-thickness layers
    pe1 = np.empty((ni, nj, nk + 1), dtype)
    pe1[:, :, 0] = v["ptop"]
    for k in range(nk):
        pe1[:, :, k + 1] = pe1[:, :, k] + delp[:, :, k]
    span = pe1[:, :, -1:] - pe1[:, :, :1]
    sig = (np.arange(nk2 + 1, dtype=np.float64) / nk2).astype(dtype)
    pe2 = (pe1[:, :, :1] - dtype(0.05) * span + dtype(1.1) * span * sig[None, None, :]).astype(dtype)
    m = min(nk, nk2)
    pe2[::2, :, 1:m:3] = pe1[::2, :, 1:m:3]  # coinciding edges (kept monotone below)
    for k in range(1, nk2 + 1):  # strictly increasing target edges (a zero-thickness target layer divides 0/0)
        prev = pe2[:, :, k - 1]
        pe2[:, :, k] = np.where(pe2[:, :, k] > prev, pe2[:, :, k], np.nextafter(prev, dtype(np.inf)))
    return {"delp": gen.as_ifirst(delp), "pe1": gen.as_ifirst(pe1), "pe2": gen.as_ifirst(pe2), "q1": v["q1"], "ptop": v["ptop"]}


@pytest.mark.parametrize("variant", sorted(REMAP_VARIANTS))
@pytest.mark.parametrize("shape", [(3, 3, 4, 4), (40, 9, 72, 75), (70, 5, 137, 150), (33, 4, 20, 90), (64, 6, 100, 30), (96, 2, 137, 137), (65, 3, 72, 72)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_remap_variants_bit_identical(st, remap_variant, variant, shape, dtype):
    """Every remap kernel (thread-per-column, shared-memory slab with cp.async or TMA loads) must give the
    oracle's bits: same overlaps, same order of accumulation (k_remap_slab.cu replaces the marching start
    by a binary search, which is the same layer for monotone edges)."""
    ni, nj, nk, nk2 = shape
    for v in (gen.vertical_inputs(ni, nj, nk, dtype, nk2=nk2), _degenerate_vertical(ni, nj, nk, nk2, dtype)):
        ref = zeros_like_np((ni, nj, nk2), dtype)
        orc.remap(v["pe1"], v["q1"], v["pe2"], ref)
        remap_variant(variant)
        q2 = up(np.zeros((ni, nj, nk2), dtype))
        st.remap(up(v["pe1"]), up(v["q1"]), up(v["pe2"]), q2)
        assert np.array_equal(down(q2), ref), f"remap, {REMAP_VARIANTS[variant]}"
        q2f = up(np.zeros((ni, nj, nk2), dtype))
        st.remap_delp(up(v["delp"]), v["ptop"], up(v["q1"]), up(v["pe2"]), q2f)
        assert np.array_equal(down(q2f), ref), f"remap_delp, {REMAP_VARIANTS[variant]}"


@pytest.mark.parametrize("dtype", DTYPES)
def test_remap_slab_unaligned_and_batched(st, remap_variant, dtype):
    """Fields that start off a 16-byte boundary (interior windows of halo-padded storage, the usual case
    for NDSL-style Quantities): the TMA loader starts its boxes at the aligned column before the window
    (`LOADER == 2` in k_remap_slab.cu); every kernel must give the same bits, and a batch of sub-domains
    in one launch must equal the per-sub-domain results."""
    from b200stencil import fields

    ni, nj, nk, nk2, nb, h = 45, 7, 72, 72, 3, 3
    vs = [gen.vertical_inputs(ni, nj, nk, dtype, cfg=5 + b, nk2=nk2) for b in range(nb)]
    refs = []
    for v in vs:
        r = zeros_like_np((ni, nj, nk2), dtype)
        orc.remap(v["pe1"], v["q1"], v["pe2"], r)
        refs.append(r)

    def padded(name, levels, hh=h):
        big = fields.zeros((ni + 2 * hh, nj + 2 * hh, levels), dtype=tdt(dtype), batch=nb)
        win = big[:, hh:hh + ni, hh:hh + nj, :]
        for b, v in enumerate(vs):
            win[b].copy_(torch.from_numpy(np.ascontiguousarray(v[name])))
        return win

    # q1 with a different halo than pe1/delp: the two slabs are shifted by different amounts
    pe1, q1, pe2, delp = padded("pe1", nk + 1), padded("q1", nk, 1), padded("pe2", nk2 + 1), padded("delp", nk)
    assert pe1.data_ptr() % 16 != 0 and q1.data_ptr() % 16 != 0  # 3 cells = 24 / 12 bytes, 1 cell = 8 / 4 bytes
    for variant in (0, 1, 2, 3):
        remap_variant(variant)
        q2 = fields.zeros((ni, nj, nk2), dtype=tdt(dtype), batch=nb)
        st.remap(pe1, q1, pe2, q2)
        q2f = fields.zeros((ni, nj, nk2), dtype=tdt(dtype), batch=nb)
        st.remap_delp(delp, vs[0]["ptop"], q1, pe2, q2f)
        for b in range(nb):
            assert np.array_equal(down(q2[b]), refs[b]), (variant, b)
            assert np.array_equal(down(q2f[b]), refs[b]), (variant, b)


def test_remap_tma_rules(st, remap_variant):
    """Row strides that are not multiples of 16 bytes cannot be described by a tensor map: the automatic
    choice uses the cp.async loader, forcing the TMA loader fails loudly (no silent change of kernel)."""
    from b200stencil import _abi

    ni, nj, nk = 33, 5, 20
    v = gen.vertical_inputs(ni, nj, nk, np.float64, nk2=nk)
    ref = zeros_like_np((ni, nj, nk), np.float64)
    orc.remap(v["pe1"], v["q1"], v["pe2"], ref)
    pe1, q1, pe2 = up(v["pe1"], align_rows=False), up(v["q1"], align_rows=False), up(v["pe2"], align_rows=False)
    assert pe1.stride(1) % 2 == 1
    remap_variant(0)
    q2 = up(np.zeros((ni, nj, nk)))
    st.remap(pe1, q1, pe2, q2)
    assert np.array_equal(down(q2), ref)
    remap_variant(3)
    with pytest.raises(_abi.B200StencilError):
        st.remap(pe1, q1, pe2, q2)


def test_remap_tall_columns_fall_back(st, remap_variant):
    """More source levels than two resident slabs allow: the automatic choice is the nested kernel."""
    ni, nj, nk = 8, 3, 600
    v = gen.vertical_inputs(ni, nj, nk, np.float64, nk2=nk)
    ref = zeros_like_np((ni, nj, nk), np.float64)
    orc.remap(v["pe1"], v["q1"], v["pe2"], ref)
    remap_variant(0)
    q2 = up(np.zeros((ni, nj, nk)))
    st.remap(up(v["pe1"]), up(v["q1"]), up(v["pe2"]), q2)
    assert np.array_equal(down(q2), ref)



# ---- S6d PPM remap ---------------------------------------------------------------------


@pytest.fixture
def ppm_options():
    from b200stencil import _abi

    def force(cols=0, loader=0):
        _abi.set_option("remap_ppm_cols", cols)
        _abi.set_option("remap_ppm_loader", loader)

    yield force
    force()


@pytest.mark.parametrize("shape", [(3, 3, 4, 4), (40, 9, 72, 75), (70, 5, 137, 150), (33, 4, 20, 90), (17, 6, 100, 30), (96, 2, 137, 137), (5, 3, 5, 9)])
@pytest.mark.parametrize("kord,iv", [(4, 1), (4, 0), (5, 0), (5, 1), (6, 1)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_remap_ppm(st, corc, ppm_options, shape, kord, iv, dtype):
    """FV3-style PPM remap against the oracle: every limiter, 16- and 32-column CTAs, TMA and cp.async loaders.
    The kernel forms reciprocals by Newton iteration, so parity is to the stated tolerance, not bit for bit."""
    ni, nj, nk, nk2 = shape
    for smooth in (True, False):
        v = gen.ppm_inputs(ni, nj, nk, dtype, nk2=nk2, smooth=smooth, positive=(iv == 0))
        ref = zeros_like_np((ni, nj, nk2), dtype)
        corc.remap_ppm(v["pe1"], v["q1"], v["pe2"], ref, kord, iv)
        for cols, loader in ((0, 0), (16, 2), (32, 2), (16, 1), (32, 1)):
            ppm_options(cols, loader)
            q2 = up(np.full((ni, nj, nk2), -7.0, dtype))
            st.remap_ppm(up(v["pe1"]), up(v["q1"]), up(v["pe2"]), q2, kord=kord, iv=iv)
            assert_close(down(q2), ref, RTOL[dtype], f"q2 cols={cols} loader={loader} smooth={smooth}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_remap_ppm_unaligned_and_batched(st, corc, ppm_options, dtype):
    """Interior windows of halo-padded storage (fields start off a 16-byte boundary: shifted TMA boxes), a batch
    of sub-domains in one launch, and row strides TMA cannot describe (cp.async loader; forcing TMA fails loudly)."""
    from b200stencil import _abi, fields

    ni, nj, nk, nk2, nb = 45, 7, 72, 72, 3
    vs = [gen.ppm_inputs(ni, nj, nk, dtype, cfg=7 + b, nk2=nk2) for b in range(nb)]
    refs = []
    for v in vs:
        r = zeros_like_np((ni, nj, nk2), dtype)
        corc.remap_ppm(v["pe1"], v["q1"], v["pe2"], r, 4, 1)
        refs.append(r)

    def padded(name, levels, hh):
        big = fields.zeros((ni + 2 * hh, nj + 2 * hh, levels), dtype=tdt(dtype), batch=nb)
        win = big[:, hh:hh + ni, hh:hh + nj, :]
        for b, v in enumerate(vs):
            win[b].copy_(torch.from_numpy(np.ascontiguousarray(v[name])))
        return win

    pe1, q1, pe2 = padded("pe1", nk + 1, 3), padded("q1", nk, 1), padded("pe2", nk2 + 1, 3)
    assert pe1.data_ptr() % 16 != 0 and q1.data_ptr() % 16 != 0
    for cols, loader in ((0, 0), (16, 2), (32, 2), (32, 1)):
        ppm_options(cols, loader)
        q2 = fields.zeros((ni, nj, nk2), dtype=tdt(dtype), batch=nb)
        st.remap_ppm(pe1, q1, pe2, q2)
        for b in range(nb):
            assert_close(down(q2[b]), refs[b], RTOL[dtype], f"batch {b} cols={cols} loader={loader}")
    v = vs[0]
    odd = [up(v[n], align_rows=False) for n in ("pe1", "q1", "pe2")]
    assert odd[0].stride(1) % 2 == 1
    ppm_options(0, 0)
    q2 = up(np.zeros((ni, nj, nk2), dtype))
    st.remap_ppm(*odd, q2)
    assert_close(down(q2), refs[0], RTOL[dtype], "odd row stride")
    ppm_options(0, 2)
    with pytest.raises(_abi.B200StencilError):
        st.remap_ppm(*odd, q2)


def test_remap_ppm_argument_errors(st):
    from b200stencil import _abi

    v = gen.ppm_inputs(4, 3, 3, np.float64)  # three source layers: the profile needs four
    with pytest.raises(_abi.B200StencilError) as e:
        st.remap_ppm(up(v["pe1"]), up(v["q1"]), up(v["pe2"]), up(np.zeros((4, 3, 3))))
    assert "4 source layers" in str(e.value) or "fewer than 4" in str(e.value)
    v = gen.ppm_inputs(4, 3, 8, np.float64)
    with pytest.raises(_abi.B200StencilError):
        st.remap_ppm(up(v["pe1"]), up(v["q1"]), up(v["pe2"]), up(np.zeros((4, 3, 8))), kord=9)
    tall = gen.ppm_inputs(2, 2, 900, np.float64)
    with pytest.raises(_abi.B200StencilError) as e:
        st.remap_ppm(up(tall["pe1"]), up(tall["q1"]), up(tall["pe2"]), up(np.zeros((2, 2, 900))))
    assert "shared-memory" in str(e.value)


# ---- error channel ---------------------------------------------------------------------------------------


def test_error_channel(st):
    from b200stencil import _abi

    I = up(np.zeros((4, 4, 4)))
    with pytest.raises(TypeError):
        st.top_of_column(I.cpu(), up(np.zeros((4, 4))), I)  # host tensor: no CPU path
    with pytest.raises(TypeError):
        st.top_of_column(I, up(np.zeros((4, 4), np.float32)), I)  # dtype mismatch: no casting
    q = up(np.zeros((10, 10, 2)))
    cx, cy = up(np.zeros((5, 4, 2))), up(np.zeros((4, 5, 2)))
    with pytest.raises(_abi.B200StencilError) as e:  # the library's own check: the C side refuses the rectangle
        st.fv_tp2d(q, cx, cx, cy, cy, up(np.zeros((4, 4))), up(np.zeros((4, 4, 2))), region=(0, 9, 0, 4))
    assert "rectangle" in str(e.value)
    # the C-ABI carries no extents with a pointer: a wrongly shaped field must be refused on the host side
    with pytest.raises(ValueError, match="crx has shape"):
        st.fv_tp2d(q, q, cx, cy, cy, up(np.zeros((4, 4))), up(np.zeros((4, 4, 2))))
    with pytest.raises(ValueError, match="q_out has shape"):
        st.fv_tp2d(q, cx, cx, cy, cy, up(np.zeros((4, 4))), up(np.zeros((4, 5, 2))))
    with pytest.raises(ValueError, match="PLEmb_top has shape"):
        st.top_of_column(I, up(np.zeros((4, 5))), I)
    with pytest.raises(ValueError, match="pe has shape"):
        st.pe_prefix(I, 1.0, up(np.zeros((4, 4, 4))))  # nk + 1 interface levels wanted
    with pytest.raises(ValueError, match="pe2 has shape"):
        st.remap(up(np.zeros((4, 4, 5))), I, up(np.zeros((4, 4, 4))), up(np.zeros((4, 4, 2))))  # q2 has 2 layers: pe2 needs 3 edges
    with pytest.raises(ValueError, match="ktop has shape"):
        st.cloud_top(I, torch.zeros((4, 3), dtype=torch.int64, device="cuda"))
    bad = torch.zeros((4, 4, 4), dtype=torch.float64, device="cuda")  # k-fastest: rejected
    with pytest.raises(ValueError):
        st.top_of_column(bad, up(np.zeros((4, 4))), I)
