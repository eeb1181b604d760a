"""-m gpu: the CUDA path, through the C-ABI, against the committed golden fixtures (tests/golden/*.npz).

patterns_golden.npz was produced from the reference's own stencil definitions (tests/golden/make_golden.py);
S1-S3 must reproduce it bit for bit.  oracle_golden.npz holds frozen oracle outputs of the stencils that have
no reference source; integer / index outputs and the sequential vertical scans bit-exact, floating-point
fields within 1e-12 relative (fp64), the tolerance BASELINE.json's north_star states.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import assert_close, down, up  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL64 = 1e-12


@pytest.fixture(scope="module")
def st():
    from b200stencil import stencils

    return stencils


@pytest.fixture(scope="module")
def pat():
    return np.load(os.path.join(GOLDEN, "patterns_golden.npz"))


@pytest.fixture(scope="module")
def frozen():
    return np.load(os.path.join(GOLDEN, "oracle_golden.npz"))


@pytest.mark.parametrize("align", [True, False])
def test_patterns_match_reference_definitions(st, pat, align):
    for c in (str(x) for x in pat["cases"]):
        d_top, d_out = up(np.full_like(pat[f"top/{c}/top"], -7), align), up(np.full_like(pat[f"top/{c}/out"], -7), align)
        st.top_of_column(up(pat[f"top/{c}/in"], align), d_top, d_out)
        assert np.array_equal(down(d_top), pat[f"top/{c}/top"]) and np.array_equal(down(d_out), pat[f"top/{c}/out"]), c

        d_out = up(np.full_like(pat[f"while/{c}/out"], -7), align)
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        st.while_in_function(up(pat[f"while/{c}/in"], align), d_out, undefined_count=cnt)
        assert np.array_equal(down(d_out), pat[f"while/{c}/out"]) and int(cnt.item()) == 0, c

        for variant in ("demo", "miss", "repeats"):
            p = f"hybrid_{variant}/{c}"
            d_out = up(np.full_like(pat[f"{p}/out"], 5), align)
            st.hybrid_index_2dout(up(pat[f"{p}/data"], align), up(pat[f"{p}/k_mask"], align), up(pat[f"{p}/k_index"], align), d_out)
            assert np.array_equal(down(d_out), pat[f"{p}/out"]), p


def _ins(frozen, group):
    pre = f"{group}/in/"
    return {k[len(pre):]: frozen[k] for k in frozen.files if k.startswith(pre)}


def test_frozen_moist(st, frozen):
    m = {k: up(v) for k, v in _ins(frozen, "moist").items()}
    klcl, pat_ = up(np.zeros_like(frozen["moist/KLCL"])), up(np.zeros_like(frozen["moist/PLmb_at_KLCL"]))
    st.find_klcl(m["p"], m["PLCL"], klcl, pat_)
    assert np.array_equal(down(klcl), frozen["moist/KLCL"]) and np.array_equal(down(pat_), frozen["moist/PLmb_at_KLCL"])
    ktop = up(np.zeros_like(frozen["moist/cloud_top"]))
    st.cloud_top(m["ql"], ktop)
    assert np.array_equal(down(ktop), frozen["moist/cloud_top"])
    st.saturation_adjust(m["T"], m["q"], m["ql"], m["p"])
    for n in ("T", "q", "ql"):
        assert_close(down(m[n]), frozen[f"moist/{n}"], RTOL64, n)


def test_frozen_horizontal(st, frozen):
    f = {k: up(v) for k, v in _ins(frozen, "fv").items()}
    out = up(np.zeros_like(frozen["fv/q_out"]))
    st.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], out)
    assert_close(down(out), frozen["fv/q_out"], RTOL64, "fv_tp2d")
    s = {k: up(v) for k, v in _ins(frozen, "fv_split").items()}
    out = up(np.zeros_like(frozen["fv_split/q_out"]))
    st.fv_tp2d_split(s["q"], s["crx"], s["xfx"], s["cry"], s["yfx"], s["area"], s["rarea"], out)
    assert_close(down(out), frozen["fv_split/q_out"], RTOL64, "fv_tp2d_split")


def test_frozen_vertical(st, frozen):
    v = _ins(frozen, "vertical")
    pe = up(np.zeros_like(frozen["vertical/pe"]))
    st.pe_prefix(up(v["delp"]), float(v["ptop"]), pe)
    assert np.array_equal(down(pe), frozen["vertical/pe"])
    q2 = up(np.zeros_like(frozen["vertical/q2"]))
    st.remap(up(v["pe1"]), up(v["q1"]), up(v["pe2"]), q2)
    assert np.array_equal(down(q2), frozen["vertical/q2"])
    q2 = up(np.zeros_like(frozen["vertical/q2"]))
    st.remap_delp(up(v["delp"]), float(v["ptop"]), up(v["q1"]), up(v["pe2"]), q2)
    assert np.array_equal(down(q2), frozen["vertical/q2"])
    p = _ins(frozen, "ppm")
    for kord, iv in ((4, 1), (5, 0), (6, 1)):
        ref = frozen[f"ppm/q2_kord{kord}_iv{iv}"]
        q2 = up(np.zeros_like(ref))
        st.remap_ppm(up(p["pe1"]), up(p["q1"]), up(p["pe2"]), q2, kord=kord, iv=iv)
        assert_close(down(q2), ref, RTOL64, f"remap_ppm kord={kord} iv={iv}")
    t = {k: up(x) for k, x in _ins(frozen, "tridiag").items()}
    x = up(np.zeros_like(frozen["tridiag/x"]))
    st.tridiag(t["a"], t["b"], t["c"], t["d"], x)
    assert_close(down(x), frozen["tridiag/x"], RTOL64, "tridiag")
