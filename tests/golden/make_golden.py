"""Golden vectors for the dsl_patterns stencils, made from the REFERENCE'S OWN SOURCE (test infrastructure).

gt4py / NDSL are not installable in this image, so the three pattern files cannot be executed as they are.
What can be done is to take their stencil definitions verbatim -- this script reads
`/root/reference/dsl_patterns/*.py`, pulls the `stencil` function and the `@function` helpers out of the
file's AST -- and run them through a small interpreter of the gtscript subset they use, with the semantics
of gt4py's numpy backend:

  * `with computation(PARALLEL), interval(a, b)`: every statement is applied to the whole k-interval (right-hand
    side evaluated for every point first, then stored) before the next statement starts;
  * `with computation(FORWARD | BACKWARD), interval(a, b)`: k sequential, all statements of the block per level,
    each statement over the whole horizontal plane (evaluate, then store);
  * `interval(a, b)` is a Python slice of the k axis (negative = from the end, None = end), `interval(...)` = all;
  * a 2-D (IJ) field broadcasts over k on reads and is written at (i, j);
  * `field[di, dj, dk]` is a relative read; `dk` may be a run-time expression (variable-K offset), unchecked in
    gt4py -- here an out-of-range read raises, so no fixture depends on undefined behaviour;
  * `if` inside a computation masks the statements of its body per point;
  * a `@function` is inlined: its parameters alias the caller's fields, its locals are per-point scalars, `while`
    loops run per point;
  * the value stored into a field is cast to the field's dtype (an integer `lev` becomes a float).

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

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference/dsl_patterns"


# ----------------------------------------------------------------------------------------------------------
# gtscript-subset interpreter
# ----------------------------------------------------------------------------------------------------------
class UndefinedRead(IndexError):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


class GtscriptProgram:
    """A stencil definition and the @function helpers of one pattern file, taken from its source text."""

    def __init__(self, path: str, stencil_name: str = "stencil"):
        with open(path) as f:
            self.source = f.read()
        tree = ast.parse(self.source)
        self.functions = {}
        self.stencil = None
        for node in tree.body:
            if isinstance(node, ast.FunctionDef):
                decos = {d.id if isinstance(d, ast.Name) else getattr(d, "attr", "") for d in node.decorator_list}
                if node.name == stencil_name:
                    self.stencil = node
                elif "function" in decos:
                    self.functions[node.name] = node
        if self.stencil is None:
            raise ValueError(f"{path}: no `{stencil_name}` definition")
        self.params = [a.arg for a in self.stencil.args.args]

    # -- execution -----------------------------------------------------------------------------------------
    def __call__(self, *arrays: np.ndarray) -> None:
        if len(arrays) != len(self.params):
            raise TypeError(f"stencil takes {self.params}, got {len(arrays)} arrays")
        fields = dict(zip(self.params, arrays))
        dom = next(a.shape for a in arrays if a.ndim == 3)
        for a in arrays:
            assert a.shape == dom or a.shape == dom[:2], (a.shape, dom)
        self.dom = dom
        for block in self.stencil.body:
            if isinstance(block, ast.Expr) and isinstance(block.value, ast.Constant):
                continue  # docstring
            if not isinstance(block, ast.With):
                raise NotImplementedError(ast.dump(block))
            order, (k0, k1) = self._with_items(block)
            ks = list(range(k0, k1))
            if order == "PARALLEL":
                for stmt in block.body:
                    self._apply(stmt, fields, ks, mask=None)
            else:
                for k in (ks if order == "FORWARD" else ks[::-1]):
                    for stmt in block.body:
                        self._apply(stmt, fields, [k], mask=None)

    def _with_items(self, block: ast.With):
        order, interval = None, None
        for item in block.items:
            call = item.context_expr
            assert isinstance(call, ast.Call) and isinstance(call.func, ast.Name), ast.dump(call)
            if call.func.id == "computation":
                order = call.args[0].id
            elif call.func.id == "interval":
                nk = self.dom[2]
                if len(call.args) == 1 and isinstance(call.args[0], ast.Constant) and call.args[0].value is Ellipsis:
                    interval = (0, nk)
                else:
                    a, b = (ast.literal_eval(x) for x in call.args)
                    interval = tuple(range(nk)[slice(a, b)][i] for i in (0, -1))
                    interval = (interval[0], interval[1] + 1)
        assert order in ("PARALLEL", "FORWARD", "BACKWARD") and interval is not None
        return order, interval

    def _points(self, ks, mask):
        ni, nj, _ = self.dom
        for k in ks:
            for i in range(ni):
                for j in range(nj):
                    if mask is None or mask[(i, j, k)]:
                        yield i, j, k

    def _apply(self, stmt, fields, ks, mask):
        """One statement over the horizontal plane x the k-set: evaluate everywhere, then store."""
        if isinstance(stmt, ast.If):
            cond = {p: bool(self._eval(stmt.test, fields, {}, *p)) for p in self._points(ks, mask)}
            full = {p: False for p in self._points(ks, None)}
            for s in stmt.body:
                self._apply(s, fields, ks, {**full, **cond})
            if stmt.orelse:
                neg = {p: not c for p, c in cond.items()}
                for s in stmt.orelse:
                    self._apply(s, fields, ks, {**full, **neg})
            return
        if isinstance(stmt, ast.Assign):
            assert len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name), ast.dump(stmt)
            target = fields[stmt.targets[0].id]
            values = {p: self._eval(stmt.value, fields, {}, *p) for p in self._points(ks, mask)}
            for (i, j, k), v in values.items():
                if target.ndim == 3:
                    target[i, j, k] = v  # NumPy casts to the field's dtype, as gt4py does
                else:
                    target[i, j] = v
            return
        raise NotImplementedError(ast.dump(stmt))

    # -- expressions ---------------------------------------------------------------------------------------
    def _read(self, arr, i, j, k):
        ni, nj, nk = self.dom
        if not (0 <= i < ni and 0 <= j < nj and 0 <= k < nk):
            raise UndefinedRead(f"read at ({i},{j},{k}) outside the {self.dom} domain: undefined in gt4py")
        return arr[i, j, k] if arr.ndim == 3 else arr[i, j]

    def _eval(self, node, fields, local, i, j, k):
        ev = lambda n: self._eval(n, fields, local, i, j, k)  # noqa: E731
        if isinstance(node, ast.Constant):
            return node.value
        if isinstance(node, ast.Name):
            if node.id in local:
                return local[node.id]
            return self._read(fields[node.id], i, j, k)
        if isinstance(node, ast.Subscript):
            arr = fields[node.value.id]
            offs = node.slice.elts if isinstance(node.slice, ast.Tuple) else [node.slice]
            di, dj, dk = (int(ev(o)) for o in offs)
            return self._read(arr, i + di, j + dj, k + dk)
        if isinstance(node, ast.UnaryOp):
            v = ev(node.operand)
            return {ast.USub: lambda: -v, ast.UAdd: lambda: +v, ast.Not: lambda: not v}[type(node.op)]()
        if isinstance(node, ast.BinOp):
            a, b = ev(node.left), ev(node.right)
            return {ast.Add: lambda: a + b, ast.Sub: lambda: a - b, ast.Mult: lambda: a * b, ast.Div: lambda: a / b,
                    ast.Pow: lambda: a ** b}[type(node.op)]()  # fmt: skip
        if isinstance(node, ast.BoolOp):
            vals = [ev(v) for v in node.values]
            return all(vals) if isinstance(node.op, ast.And) else any(vals)
        if isinstance(node, ast.Compare):
            left = ev(node.left)
            for op, right in zip(node.ops, node.comparators):
                r = ev(right)
                ok = {ast.Lt: left < r, ast.LtE: left <= r, ast.Gt: left > r, ast.GtE: left >= r, ast.Eq: left == r,
                      ast.NotEq: left != r}[type(op)]  # fmt: skip
                if not ok:
                    return False
                left = r
            return True
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in self.functions:
            fn = self.functions[node.func.id]
            # field arguments alias the caller's fields (gt4py inlines the function)
            inner = dict(fields)
            for p, a in zip(fn.args.args, node.args):
                assert isinstance(a, ast.Name) and a.id in fields, "only field arguments are used by the patterns"
                inner[p.arg] = fields[a.id]
            try:
                self._run_body(fn.body, inner, {}, i, j, k)
            except _Return as r:
                return r.value
            raise ValueError(f"{fn.name} returned nothing")
        raise NotImplementedError(ast.dump(node))

    def _run_body(self, body, fields, local, i, j, k):
        for s in body:
            if isinstance(s, ast.Expr) and isinstance(s.value, ast.Constant):
                continue
            if isinstance(s, ast.Assign):
                local[s.targets[0].id] = self._eval(s.value, fields, local, i, j, k)
            elif isinstance(s, ast.AugAssign):
                cur = local[s.target.id]
                inc = self._eval(s.value, fields, local, i, j, k)
                local[s.target.id] = {ast.Add: cur + inc, ast.Sub: cur - inc, ast.Mult: cur * inc}[type(s.op)]
            elif isinstance(s, ast.While):
                while self._eval(s.test, fields, local, i, j, k):
                    self._run_body(s.body, fields, local, i, j, k)
            elif isinstance(s, ast.If):
                self._run_body(s.body if self._eval(s.test, fields, local, i, j, k) else s.orelse, fields, local, i, j, k)
            elif isinstance(s, ast.Return):
                raise _Return(self._eval(s.value, fields, local, i, j, k))
            else:
                raise NotImplementedError(ast.dump(s))


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
    top = GtscriptProgram(os.path.join(REFERENCE, "Do__get_top_of_the_column.py"))
    whl = GtscriptProgram(os.path.join(REFERENCE, "Do__while_in_gt_functions.py"))
    hyb = GtscriptProgram(os.path.join(REFERENCE, "WIP__hybrid_index_2dout.py"))
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
