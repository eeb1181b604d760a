"""-m gpu tests at BASELINE.json's FULL sizes (cfg2 C96x72, cfg3 C180x72, cfg4 C384x72, cfg5 C720x137).

The oracle cannot sweep 4e8 points in seconds, so each stencil is checked at full size twice:
  * through a size-independent PROPERTY of the stencil evaluated on the device with elementwise torch
    ops as the checker (recurrences that hold level by level, conservation, linearity, index
    predicates) over EVERY point of the output, and
  * against the oracle on windows of columns / cells cut out of the full-size fields at seeded
    positions (including the first and last rows of a sub-domain), bit-exact where the small-size parity
    tests are bit-exact.
Inputs are the device-generated synthetic fields of the benchmark (b200stencil/bench/workloads.py).
"""
import numpy as np
import pytest
import torch

from oracle import inputs as gen
from oracle import numpy_oracle as orc
from oracle.c_oracle import COracle

pytestmark = pytest.mark.gpu

from gpu_util import assert_close, zeros_like_np  # noqa: E402

F64, F32 = torch.float64, torch.float32
RTOL = {F64: 1e-12, F32: 1e-5}


@pytest.fixture(scope="module")
def env():
    from b200stencil import stencils
    from b200stencil.bench import workloads

    return stencils, workloads


def windows(tiles, ni, nj, w=16, h=8, n=6, seed=11):
    """Seeded (b, i0, j0) window origins, always including the corners of the first and last sub-domain."""
    rng = np.random.default_rng(seed)
    out = [(0, 0, 0), (tiles - 1, ni - w, nj - h)]
    for _ in range(n):
        out.append((int(rng.integers(tiles)), int(rng.integers(0, ni - w + 1)), int(rng.integers(0, nj - h + 1))))
    return [(b, i0, j0, w, h) for b, i0, j0 in out]


def cut(t, b, i0, j0, w, h, halo=0, extra=(0, 0)):
    """Window of a device field [b,i,j(,k)] as an i-fastest NumPy array."""
    sl = t[b, i0 : i0 + w + 2 * halo + extra[0], j0 : j0 + h + 2 * halo + extra[1]]
    return gen.as_ifirst(sl.cpu().numpy())


def make(workloads, stencil, cfg, dtype):
    tiles, n, nk = workloads.CONFIGS[cfg]
    wl = workloads.make(stencil, tiles, n, nk, dtype, slots=1)
    return wl, wl.keep[0], tiles, n, nk


def np_dtype(dtype):
    return np.float64 if dtype == F64 else np.float32


# ---- cfg2: the reference's own patterns on C96 x 72 ------------------------------------------------------


@pytest.mark.parametrize("dtype", [F64, F32])
def test_patterns_full_c96(env, dtype):
    st, workloads = env
    wl, (I, top, O), tiles, n, nk = make(workloads, "top_of_column", "C96x72", dtype)
    wl.run(0)
    assert torch.equal(top, I[..., nk - 1]) and torch.equal(O, I[..., nk - 1 :].expand_as(O))  # pure moves

    wl, (I, O), tiles, n, nk = make(workloads, "while_in_function", "C96x72", dtype)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    st.while_in_function(I, O, undefined_count=cnt)
    assert int(cnt.item()) == 0
    hit = ~(I < 4.0)
    assert bool((O[hit] == 0).all())  # distance 0 where the level itself stops the loop
    nxt = torch.cat([O[..., 1:] + 1, torch.zeros_like(O[..., :1])], dim=-1)
    assert torch.equal(torch.where(hit, torch.zeros_like(O), nxt), O)  # BACKWARD recurrence on every point
    for b, i0, j0, w, h in windows(tiles, n, n):
        ref = zeros_like_np((w, h, nk), np_dtype(dtype))
        orc.while_in_function_scan(cut(I, b, i0, j0, w, h), ref)
        assert np.array_equal(cut(O, b, i0, j0, w, h), ref)

    wl, (data, kmask, kidx, out), tiles, n, nk = make(workloads, "hybrid_index_2dout", "C96x72", dtype)
    wl.run(0)
    want = torch.gather(data, -1, kidx.long().unsqueeze(-1)).squeeze(-1)  # k_mask[..., k] == k in this workload
    assert torch.equal(out, want)


# ---- cfg3: moist column stencils on C180 x 72 ------------------------------------------------------------


@pytest.mark.parametrize("dtype", [F64, F32])
def test_moist_full_c180(env, dtype):
    st, workloads = env
    wl, (p, plcl, klcl, pat), tiles, n, nk = make(workloads, "find_klcl", "C180x72", dtype)
    wl.run(0)
    k = klcl.long()
    assert bool((k >= 0).all())
    at = torch.gather(p, -1, k.unsqueeze(-1)).squeeze(-1)
    assert torch.equal(at, pat) and bool((at <= plcl).all())  # the level found satisfies the predicate ...
    below = torch.gather(p, -1, (k + 1).clamp(max=nk - 1).unsqueeze(-1)).squeeze(-1)
    assert bool(((below > plcl) | (k == nk - 1)).all())  # ... and it is the first one from the surface up

    wl, (ql, ktop), tiles, n, nk = make(workloads, "cloud_top", "C180x72", dtype)
    wl.run(0)
    cloudy = ql > 1.0e-8
    first = torch.where(cloudy.any(-1), cloudy.to(torch.int8).argmax(-1), torch.full_like(ktop.long(), -1))
    assert torch.equal(ktop.long(), first)

    wl, (T, q, l, pp), tiles, n, nk = make(workloads, "saturation_adjust", "C180x72", dtype)
    T0, q0, l0 = T.clone(), q.clone(), l.clone()
    wl.run(0)
    tol = 64 * torch.finfo(dtype).eps
    # water is conserved, the latent heat of what condensed went into T, ql never goes negative
    assert float(((q + l) - (q0 + l0)).abs().max() / (q0 + l0).abs().max()) < tol
    assert float((T - T0 - (2.5e6 / 1004.0) * (l - l0)).abs().max() / T0.abs().max()) < tol
    assert float(l.min()) >= -float(l0.abs().max()) * tol
    for b, i0, j0, w, h in windows(tiles, n, n, n=3):
        a = [cut(x, b, i0, j0, w, h) for x in (T0, q0, l0)]
        orc.saturation_adjust(a[0], a[1], a[2], cut(pp, b, i0, j0, w, h))
        for got, want, name in zip((T, q, l), a, ("T", "q", "ql")):
            assert_close(cut(got, b, i0, j0, w, h), want, RTOL[dtype], name)


# ---- cfg4: fv_tp2d on C384 x 72 ---------------------------------------------------------------------------


@pytest.mark.parametrize("dtype", [F64, F32])
def test_fv_tp2d_full_c384(env, dtype):
    st, workloads = env
    corc = COracle()
    wl, (q, crx, xfx, cry, yfx, rarea, out), tiles, n, nk = make(workloads, "fv_tp2d", "C384x72", dtype)
    wl.run(0)
    # windows against the C oracle (q carries its 3-cell halo: storage index = compute index + 3)
    for b, i0, j0, w, h in windows(tiles, n, n, w=32, h=16, n=4):
        ref = zeros_like_np((w, h, nk), np_dtype(dtype))
        corc.fv_tp2d(cut(q, b, i0, j0, w, h, halo=3), cut(crx, b, i0, j0, w, h, extra=(1, 0)), cut(xfx, b, i0, j0, w, h, extra=(1, 0)),
                     cut(cry, b, i0, j0, w, h, extra=(0, 1)), cut(yfx, b, i0, j0, w, h, extra=(0, 1)), cut(rarea, b, i0, j0, w, h), ref)
        assert_close(cut(out, b, i0, j0, w, h), ref, RTOL[dtype], f"q_out window {(b, i0, j0)}")
    # the unlimited PPM operator is linear in q: L(a q + c r) = a L(q) + c L(r) on every point
    r = torch.empty_like(q).uniform_(0.5, 1.5)
    out_r, out_mix = torch.empty_like(out), torch.empty_like(out)
    st.fv_tp2d(r, crx, xfx, cry, yfx, rarea, out_r)
    mix = 0.75 * q + 0.5 * r
    st.fv_tp2d(mix, crx, xfx, cry, yfx, rarea, out_mix)
    lin = 0.75 * out + 0.5 * out_r
    scale = float(lin.abs().max())
    assert float((out_mix - lin).abs().max()) <= 50 * RTOL[dtype] * scale


# ---- cfg5: vertical scans on C720 x 137 -------------------------------------------------------------------


@pytest.mark.parametrize("dtype", [F64, F32])
def test_vertical_full_c720(env, dtype):
    st, workloads = env
    wl, (delp, ptop, pe), tiles, n, nk = make(workloads, "pe_prefix", "C720x137", dtype)
    wl.run(0)
    assert bool((pe[..., 0] == ptop).all())
    assert torch.equal(pe[..., 1:], pe[..., :-1] + delp)  # the FORWARD recurrence, one IEEE add per level: bit-exact
    del wl, delp, pe
    torch.cuda.empty_cache()

    wl, (delp, ptop, q1, pe2, q2), tiles, n, nk = make(workloads, "remap_delp", "C720x137", dtype)
    wl.run(0)
    pe1 = torch.empty_like(pe2)
    st.pe_prefix(delp, ptop, pe1)
    q2b = torch.empty_like(q2)
    st.remap(pe1, q1, pe2, q2b)
    assert torch.equal(q2, q2b)  # pe_prefix fused into the remap == the two kernels, bit for bit
    # conservative: the column mass of q is unchanged (pe2 spans the same pressure range as pe1)
    m1 = (q1 * delp).sum(-1, dtype=torch.float64)
    m2 = (q2 * (pe2[..., 1:] - pe2[..., :-1])).sum(-1, dtype=torch.float64)
    assert float(((m2 - m1).abs() / m1.abs()).max()) < (1e-12 if dtype == F64 else 2e-5)
    # bounded: a piecewise-constant remap cannot leave the range of its source column
    slack = 1e-12 if dtype == F64 else 1e-5
    assert bool((q2 <= q1.amax(-1, keepdim=True) * (1 + slack)).all()) and bool((q2 >= q1.amin(-1, keepdim=True) * (1 - slack)).all())
    for b, i0, j0, w, h in windows(tiles, n, n, w=16, h=4, n=4):
        ref = zeros_like_np((w, h, nk), np_dtype(dtype))
        orc.remap(cut(pe1, b, i0, j0, w, h), cut(q1, b, i0, j0, w, h), cut(pe2, b, i0, j0, w, h), ref)
        assert np.array_equal(cut(q2, b, i0, j0, w, h), ref), (b, i0, j0)
    del wl, delp, q1, pe2, q2, q2b, pe1, m1, m2
    torch.cuda.empty_cache()

    wl, (a, bb, c, d, x, wk), tiles, n, nk = make(workloads, "tridiag", "C720x137", dtype)
    wl.run(0)
    # residual of the solve on every point: a x[k-1] + b x[k] + c x[k+1] = d
    res = bb * x - d
    res[..., 1:] += a[..., 1:] * x[..., :-1]
    res[..., :-1] += c[..., :-1] * x[..., 1:]
    assert float(res.abs().max()) < (1e-13 if dtype == F64 else 1e-5)
