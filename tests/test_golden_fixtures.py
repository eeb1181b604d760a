"""The CPU oracle against the committed golden fixtures (tests/golden/*.npz).

patterns_golden.npz  outputs of the REFERENCE's own stencil definitions (dsl_patterns/*.py, taken from source
                     and run through tests/golden/make_golden.py's gtscript interpreter): pins S1, S2, S3.
oracle_golden.npz    frozen oracle outputs for the stencils without reference source: drift protection.
"""
import os

import numpy as np
import pytest

from oracle import inputs as gen
from oracle import numpy_oracle as orc
from oracle.c_oracle import COracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pat():
    return np.load(os.path.join(GOLDEN, "patterns_golden.npz"))


@pytest.fixture(scope="module")
def frozen():
    return np.load(os.path.join(GOLDEN, "oracle_golden.npz"))


@pytest.fixture(scope="module")
def corc():
    return COracle()


def _cases(pat):
    return [str(c) for c in pat["cases"]]


def test_fixture_inventory(pat):
    cases = _cases(pat)
    assert len(cases) == 10 and "float64_3x3x4" in cases and "float32_1x1x1" in cases
    for c in cases:
        for key in ("top/{}/out", "while/{}/out", "hybrid_demo/{}/out", "hybrid_miss/{}/out", "hybrid_repeats/{}/out"):
            assert key.format(c) in pat.files
    assert "reference's own stencil definitions" in str(pat["note"])


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_patterns_match_reference_definitions(pat, corc, impl):
    """Bit-exact: S1-S3 are moves and small-integer counts."""
    o = orc if impl == "numpy" else corc
    prep = (lambda a: a.copy()) if impl == "numpy" else gen.as_ifirst
    for c in _cases(pat):
        x = prep(pat[f"top/{c}/in"])
        tmp, out = prep(np.full_like(pat[f"top/{c}/top"], -7)), prep(np.full_like(pat[f"top/{c}/out"], -7))
        o.top_of_column(x, tmp, out)
        assert np.array_equal(tmp, pat[f"top/{c}/top"]) and np.array_equal(out, pat[f"top/{c}/out"]), c

        x = prep(pat[f"while/{c}/in"])
        out = prep(np.full_like(x, -7))
        o.while_in_function(x, out)
        assert np.array_equal(out, pat[f"while/{c}/out"]), c
        if impl == "numpy":  # the single-scan formulation the CUDA kernel uses
            out2 = np.full_like(x, -7)
            assert orc.while_in_function_scan(x, out2) == 0
            assert np.array_equal(out2, pat[f"while/{c}/out"]), c

        for variant in ("demo", "miss", "repeats"):
            p = f"hybrid_{variant}/{c}"
            out = prep(np.full_like(pat[f"{p}/out"], 5))
            o.hybrid_index_2dout(prep(pat[f"{p}/data"]), prep(pat[f"{p}/k_mask"]), prep(pat[f"{p}/k_index"]), out)
            assert np.array_equal(out, pat[f"{p}/out"]), p


def test_hybrid_fixture_exercises_the_edge_cases(pat):
    c = "float64_5x4x7"
    miss = pat[f"hybrid_miss/{c}/k_index"] < 0
    assert miss.any() and (pat[f"hybrid_miss/{c}/out"][miss] == 5).all()  # no match: previous value kept
    km, ki, data = (pat[f"hybrid_repeats/{c}/{n}"] for n in ("k_mask", "k_index", "data"))
    nmatch = (km == ki[:, :, None]).sum(axis=2)
    assert (nmatch > 1).any()  # repeated mask values: the LAST matching level wins
    i, j = np.argwhere(nmatch > 1)[0]
    last = np.nonzero(km[i, j] == ki[i, j])[0][-1]
    assert pat[f"hybrid_repeats/{c}/out"][i, j] == data[i, j, last]


@pytest.mark.skipif(not os.path.isdir("/root/reference/dsl_patterns"), reason="reference tree only exists in the dev container")
def test_fixtures_regenerate_from_the_reference_source(pat):
    """The committed file is what the generator produces from the reference source today."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.pattern_fixtures()
    assert sorted(fresh) == sorted(pat.files)
    for k in pat.files:
        assert np.array_equal(fresh[k], pat[k]), k


# ---- frozen oracle outputs (no reference source: parity unpinned, this only guards against drift) ----


def _ins(frozen, group):
    pre = f"{group}/in/"
    return {k[len(pre):]: frozen[k] for k in frozen.files if k.startswith(pre)}


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_has_not_drifted(frozen, corc, impl):
    o = orc if impl == "numpy" else corc
    prep = (lambda a: a.copy()) if impl == "numpy" else gen.as_ifirst
    z = lambda ref: prep(np.zeros_like(ref))  # noqa: E731
    tol = dict(rtol=0, atol=0) if impl == "numpy" else dict(rtol=1e-13, atol=0)

    m = {k: prep(v) for k, v in _ins(frozen, "moist").items()}
    klcl, pat_ = z(frozen["moist/KLCL"]), z(frozen["moist/PLmb_at_KLCL"])
    o.find_klcl(m["p"], m["PLCL"], klcl, pat_)
    assert np.array_equal(klcl, frozen["moist/KLCL"]) and np.array_equal(pat_, frozen["moist/PLmb_at_KLCL"])
    ktop = z(frozen["moist/cloud_top"])
    o.cloud_top(m["ql"], ktop)
    assert np.array_equal(ktop, frozen["moist/cloud_top"])
    o.saturation_adjust(m["T"], m["q"], m["ql"], m["p"])
    for n in ("T", "q", "ql"):
        np.testing.assert_allclose(m[n], frozen[f"moist/{n}"], **tol)

    f = {k: prep(v) for k, v in _ins(frozen, "fv").items()}
    out = z(frozen["fv/q_out"])
    o.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], out)
    np.testing.assert_allclose(out, frozen["fv/q_out"], **tol)
    s = {k: prep(v) for k, v in _ins(frozen, "fv_split").items()}
    out = z(frozen["fv_split/q_out"])
    o.fv_tp2d_split(s["q"], s["crx"], s["xfx"], s["cry"], s["yfx"], s["area"], s["rarea"], out)
    np.testing.assert_allclose(out, frozen["fv_split/q_out"], **tol)

    v = {k: prep(x) if x.ndim else x for k, x in _ins(frozen, "vertical").items()}
    pe = z(frozen["vertical/pe"])
    o.pe_prefix(v["delp"], float(v["ptop"]), pe)
    assert np.array_equal(pe, frozen["vertical/pe"])
    q2 = z(frozen["vertical/q2"])
    o.remap(v["pe1"], v["q1"], v["pe2"], q2)
    assert np.array_equal(q2, frozen["vertical/q2"])
    p = {k: prep(x) if x.ndim else x for k, x in _ins(frozen, "ppm").items()}
    for kord, iv in ((4, 1), (5, 0), (6, 1)):
        ref = frozen[f"ppm/q2_kord{kord}_iv{iv}"]
        q2 = z(ref)
        o.remap_ppm(p["pe1"], p["q1"], p["pe2"], q2, kord=kord, iv=iv)
        np.testing.assert_allclose(q2, ref, **tol)
    t = {k: prep(x) for k, x in _ins(frozen, "tridiag").items()}
    x = z(frozen["tridiag/x"])
    o.tridiag(t["a"], t["b"], t["c"], t["d"], x)
    np.testing.assert_allclose(x, frozen["tridiag/x"], **tol)
