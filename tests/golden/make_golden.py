"""Golden vectors for the dsl_patterns stencils, made from the REFERENCE'S OWN SOURCE (test infrastructure).

gt4py / NDSL are not installable in this image, so the three pattern files cannot be executed as they are.
What can be done is to take their stencil definitions verbatim -- this script reads
`/root/reference/dsl_patterns/*.py`, pulls the `stencil` function and the `@function` helpers out of the
file's AST -- and run them through `gtscript_interp.py` (beside this file), a small interpreter of the gtscript
subset they use with the semantics of gt4py's numpy backend (listed in its docstring).

The interpreter loops over points in pure Python: small cases only.  The outputs, together with the seeded
inputs that produced them, are committed as `patterns_golden.npz`; `tests/test_golden_fixtures.py` checks the
CPU oracle (NumPy and C) against them and `tests/test_gpu_golden.py` checks the CUDA path -- bit for bit.
`/root/reference` exists only in the development container: this script is run there, once, by hand
(`python tests/golden/make_golden.py`); nothing at test time reads the reference.

The second file, `oracle_golden.npz`, freezes outputs of the oracle itself for the stencils that have no source
in the reference (S4-S6, see SURVEY.md 8c: "parity unpinned"): it pins the oracle against drift, NOT against
the reference, and says so in its `note` entry.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference/dsl_patterns"


sys.path.insert(0, HERE)
from gtscript_interp import GtscriptProgram  # noqa: E402

# ----------------------------------------------------------------------------------------------------------
# fixtures
# ----------------------------------------------------------------------------------------------------------
SHAPES = [(3, 3, 4), (5, 4, 7), (2, 9, 13), (6, 1, 24), (1, 1, 1)]


def _column_input(rng, shape, dtype):
    """SURVEY.md 8d cfg2 recipe: values < 4 everywhere, 1-3 hit levels per column, 42 at the last level."""
    ni, nj, nk = shape
    a = rng.uniform(0.0, 3.999, size=shape)
    for i in range(ni):
        for j in range(nj):
            for h in rng.integers(0, nk, size=rng.integers(1, 4)):
                a[i, j, h] = 4.0 + rng.uniform(0.0, 38.0)
    a[:, :, nk - 1] = 42.0
    return a.astype(dtype)


def pattern_fixtures() -> dict:
    top = GtscriptProgram.from_file(os.path.join(REFERENCE, "Do__get_top_of_the_column.py"))
    whl = GtscriptProgram.from_file(os.path.join(REFERENCE, "Do__while_in_gt_functions.py"))
    hyb = GtscriptProgram.from_file(os.path.join(REFERENCE, "WIP__hybrid_index_2dout.py"))
    assert top.params == ["PLEmb", "PLEmb_top", "out_field"]
    assert whl.params == ["in_field", "out_field"] and "while_in_function" in whl.functions
    assert hyb.params == ["data_field", "k_mask", "k_index_desired", "out_field"]
    out = {}
    cases = []

    # the reference's own demo inputs and asserts first (Do__get_top_of_the_column.py:59-68,
    # Do__while_in_gt_functions.py:52-62): the interpreter itself must reproduce them
    I = np.ones((3, 3, 4))
    I[:, :, 3] = 42
    O, tmp = np.zeros((3, 3, 4)), np.zeros((3, 3))
    top(I, tmp, O)
    assert np.all(O == 42)
    O = np.zeros((3, 3, 4))
    whl(I, O)
    assert (O[0, 0, :] == [3.0, 2.0, 1.0, 0.0]).all()

    rng = np.random.default_rng(20240724)
    for dtype in (np.float64, np.float32):
        for shape in SHAPES:
            tag = f"{np.dtype(dtype).name}_{shape[0]}x{shape[1]}x{shape[2]}"
            cases.append(tag)
            # S1 top_of_column
            x = (1000.0 * (np.arange(shape[2]) + 1) / shape[2] + rng.uniform(0, 1, size=shape)).astype(dtype)
            tmp = np.full(shape[:2], -7, dtype)
            o = np.full(shape, -7, dtype)
            top(x, tmp, o)
            out[f"top/{tag}/in"], out[f"top/{tag}/top"], out[f"top/{tag}/out"] = x, tmp, o
            # S2 while_in_function
            x = _column_input(rng, shape, dtype)
            o = np.full(shape, -7, dtype)
            whl(x, o)
            out[f"while/{tag}/in"], out[f"while/{tag}/out"] = x, o
            # S3 hybrid_index_2dout: (a) the demo's k_mask[...,k] = k; (b) 1 column in 4 without a match (k_index = -1):
            # the output keeps its previous value; (c) arbitrary mask contents with repeats: the LAST match wins
            data = rng.integers(800, 900, size=shape).astype(dtype)
            kmask = np.broadcast_to(np.arange(shape[2], dtype=dtype), shape).copy()
            kidx = rng.integers(0, shape[2], size=shape[:2]).astype(dtype)
            for variant in ("demo", "miss", "repeats"):
                km, ki = kmask.copy(), kidx.copy()
                if variant == "miss":
                    ki[rng.uniform(size=shape[:2]) < 0.25] = -1
                if variant == "repeats":
                    km = rng.integers(0, max(2, shape[2] // 2), size=shape).astype(dtype)
                o = np.full(shape[:2], 5, dtype)  # a visible previous value
                hyb(data, km, ki, o)
                p = f"hybrid_{variant}/{tag}"
                out[f"{p}/data"], out[f"{p}/k_mask"], out[f"{p}/k_index"], out[f"{p}/out"] = data, km, ki, o
    out["cases"] = np.array(cases)
    out["note"] = np.array(
        "outputs of the reference's own stencil definitions (dsl_patterns/*.py, taken from source) run through "
        "tests/golden/make_golden.py's gtscript interpreter with gt4py numpy-backend semantics"
    )
    return out


def oracle_fixtures() -> dict:
    """Frozen oracle outputs for S4-S6 (no reference source): drift protection only."""
    sys.path.insert(0, ROOT)
    from oracle import inputs as gen
    from oracle import numpy_oracle as orc

    out = {}
    ni, nj, nk = 7, 5, 12
    m = gen.moist_inputs(ni, nj, nk)
    klcl, pat = np.zeros((ni, nj), np.int64), np.zeros((ni, nj))
    orc.find_klcl(m["p"], m["PLCL"], klcl, pat)
    T, q, ql = m["T"].copy(), m["q"].copy(), m["ql"].copy()
    orc.saturation_adjust(T, q, ql, m["p"])
    ktop = np.zeros((ni, nj), np.int64)
    orc.cloud_top(m["ql"], ktop)
    for k, v in m.items():
        out[f"moist/in/{k}"] = v
    out["moist/KLCL"], out["moist/PLmb_at_KLCL"], out["moist/cloud_top"] = klcl, pat, ktop
    out["moist/T"], out["moist/q"], out["moist/ql"] = T, q, ql

    ni, nj, nk = 12, 9, 3
    f = gen.fv_inputs(ni, nj, nk)
    qo = np.zeros((ni, nj, nk))
    orc.fv_tp2d(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["rarea"], qo)
    for k, v in f.items():
        out[f"fv/in/{k}"] = v
    out["fv/q_out"] = qo

    s = gen.fv_split_inputs(ni, nj, nk)
    qs = np.zeros((ni, nj, nk))
    orc.fv_tp2d_split(s["q"], s["crx"], s["xfx"], s["cry"], s["yfx"], s["area"], s["rarea"], qs)
    for k, v in s.items():
        out[f"fv_split/in/{k}"] = v
    out["fv_split/q_out"] = qs

    ni, nj, nk = 6, 4, 15
    v = gen.vertical_inputs(ni, nj, nk, nk2=11)
    pe = np.zeros((ni, nj, nk + 1))
    orc.pe_prefix(v["delp"], float(v["ptop"]), pe)
    q2 = np.zeros((ni, nj, 11))
    orc.remap(v["pe1"], v["q1"], v["pe2"], q2)
    for k, x in v.items():
        out[f"vertical/in/{k}"] = np.asarray(x)
    out["vertical/pe"], out["vertical/q2"] = pe, q2
    p = gen.ppm_inputs(ni, nj, nk, nk2=11)
    for kord, iv in ((4, 1), (5, 0), (6, 1)):
        q2p = np.zeros((ni, nj, 11))
        orc.remap_ppm(p["pe1"], p["q1"], p["pe2"], q2p, kord=kord, iv=iv)
        out[f"ppm/q2_kord{kord}_iv{iv}"] = q2p
    for k, x in p.items():
        out[f"ppm/in/{k}"] = np.asarray(x)
    t = gen.tridiag_inputs(ni, nj, nk)
    x = np.zeros((ni, nj, nk))
    orc.tridiag(t["a"], t["b"], t["c"], t["d"], x)
    for k, a in t.items():
        out[f"tridiag/in/{k}"] = a
    out["tridiag/x"] = x
    out["note"] = np.array(
        "frozen outputs of oracle/numpy_oracle.py for the stencils WITHOUT reference source (S4, S5, S5b, S6): "
        "pins the oracle against drift, not against the reference (parity unpinned, SURVEY.md 8c)"
    )
    return out


def main() -> int:
    if not os.path.isdir(REFERENCE):
        print(f"{REFERENCE} not found: the fixtures can only be regenerated in the development container", file=sys.stderr)
        return 1
    pat = pattern_fixtures()
    np.savez_compressed(os.path.join(HERE, "patterns_golden.npz"), **pat)
    orc = oracle_fixtures()
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **orc)
    for name in ("patterns_golden.npz", "oracle_golden.npz"):
        print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
