"""A small interpreter of the gtscript subset used by the dsl_patterns files and by b200stencil.definitions
(test infrastructure; gt4py itself cannot be installed in this image).

It runs a stencil DEFINITION -- taken as an AST from a source file or from a Python function object -- point by
point with the semantics of gt4py's numpy backend:

  * `with computation(PARALLEL), interval(a, b)`: every statement is applied to the whole k-interval (right-hand
    side evaluated for every point first, then stored) before the next statement starts;
  * `with computation(FORWARD | BACKWARD)`: k sequential, all statements of an interval block per level, each
    statement over the whole horizontal plane (evaluate, then store); several `with interval(...)` blocks inside
    one computation run in the order they are written (FV3 idiom: BACKWARD starts with `interval(-1, None)`);
  * `interval(a, b)` is a Python slice of the k axis (negative = from the end, None = end), `interval(...)` = all;
  * a 2-D (IJ) field broadcasts over k on reads and is written at (i, j); a scalar parameter is a scalar;
  * `field[di, dj, dk]` is a relative read; `dk` may be a run-time expression (variable-K offset).  A field may be
    larger than the compute domain: its `origin` says which element is compute point (0, 0, 0), so horizontal
    offsets read halo cells.  gt4py leaves out-of-range reads undefined -- here they raise;
  * a name first assigned in the stencil body is a temporary: a full 3-D field of the compute domain;
  * `if` / `else` inside a computation mask the statements of their bodies per point;
  * a `@function` is inlined: plain field-name arguments alias the caller's fields (so the helper can offset
    them), any other argument expression is evaluated at the point and bound as a scalar; locals are per-point
    scalars; `while` loops run per point;
  * `THIS_K` is the level index of the point; `exp, log, sqrt, abs, min, max, floor` are the usual maths;
  * the value stored into a field is cast to the field's dtype (an integer `lev` becomes a float).

Pure-Python loops over points: small cases only.
"""
from __future__ import annotations

import ast
import inspect
import textwrap
from typing import Dict, Optional, Sequence

import numpy as np


class UndefinedRead(IndexError):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


_MATH = {"exp": np.exp, "log": np.log, "sqrt": np.sqrt, "abs": abs, "min": min, "max": max, "floor": np.floor}


def _functions_of(tree: ast.Module) -> Dict[str, ast.FunctionDef]:
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            decos = {d.id if isinstance(d, ast.Name) else getattr(d, "attr", "") for d in node.decorator_list}
            if "function" in decos:
                out[node.name] = node
    return out


class GtscriptProgram:
    """A stencil definition plus the `@function` helpers of its module."""

    def __init__(self, stencil: ast.FunctionDef, functions: Dict[str, ast.FunctionDef]):
        self.stencil = stencil
        self.functions = functions
        self.params = [a.arg for a in stencil.args.args + stencil.args.kwonlyargs]

    @classmethod
    def from_file(cls, path: str, stencil_name: str = "stencil") -> "GtscriptProgram":
        with open(path) as f:
            tree = ast.parse(f.read())
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name == stencil_name:
                return cls(node, _functions_of(tree))
        raise ValueError(f"{path}: no `{stencil_name}` definition")

    @classmethod
    def from_function(cls, fn) -> "GtscriptProgram":
        mod_tree = ast.parse(inspect.getsource(inspect.getmodule(fn)))
        node = ast.parse(textwrap.dedent(inspect.getsource(fn))).body[0]
        assert isinstance(node, ast.FunctionDef)
        return cls(node, _functions_of(mod_tree))

    # -- execution -----------------------------------------------------------------------------------------
    def __call__(self, *args, domain: Optional[Sequence[int]] = None, origins: Optional[Dict[str, Sequence[int]]] = None) -> None:
        if len(args) != len(self.params):
            raise TypeError(f"stencil takes {self.params}, got {len(args)} arguments")
        self.fields: Dict[str, np.ndarray] = {}
        self.scalars: Dict[str, object] = {}
        for name, a in zip(self.params, args):
            if isinstance(a, np.ndarray) and a.ndim >= 2:
                self.fields[name] = a
            else:
                self.scalars[name] = a
        self.dom = tuple(domain) if domain is not None else next(a.shape for a in self.fields.values() if a.ndim == 3)
        self._origin = {id(self.fields[k]): tuple(v) for k, v in (origins or {}).items()}  # keyed by storage, so aliases share it
        self.float_dtype = next(a.dtype for a in self.fields.values() if a.dtype.kind == "f")
        for block in self.stencil.body:
            if isinstance(block, ast.Expr) and isinstance(block.value, ast.Constant):
                continue  # docstring
            if not isinstance(block, ast.With):
                raise NotImplementedError(ast.dump(block))
            order, interval = self._with_items(block)
            if interval is not None:
                self._run_interval(order, interval, block.body)
            else:  # `with computation(X):` holding `with interval(...):` blocks
                for inner in block.body:
                    assert isinstance(inner, ast.With), ast.dump(inner)
                    _, iv = self._with_items(inner)
                    self._run_interval(order, iv, inner.body)

    def _run_interval(self, order, interval, body):
        ks = list(range(*interval))
        if order == "PARALLEL":
            for stmt in body:
                self._apply(stmt, ks, mask=None)
        else:
            for k in ks if order == "FORWARD" else ks[::-1]:
                for stmt in body:
                    self._apply(stmt, [k], mask=None)

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
                    sel = range(nk)[slice(a, b)]
                    interval = (sel[0], sel[-1] + 1) if len(sel) else (0, 0)
        return order, interval

    def _points(self, ks, mask):
        ni, nj, _ = self.dom
        for k in ks:
            for i in range(ni):
                for j in range(nj):
                    if mask is None or mask[(i, j, k)]:
                        yield i, j, k

    def _apply(self, stmt, ks, mask):
        """One statement over the horizontal plane x the k-set: evaluate everywhere, then store."""
        if isinstance(stmt, ast.If):
            cond = {p: bool(self._eval(stmt.test, self.fields, {}, *p)) for p in self._points(ks, mask)}
            full = {p: False for p in self._points(ks, None)}
            for s in stmt.body:
                self._apply(s, ks, {**full, **cond})
            if stmt.orelse:
                neg = {p: not c for p, c in cond.items()}
                for s in stmt.orelse:
                    self._apply(s, ks, {**full, **neg})
            return
        if isinstance(stmt, ast.Assign):
            assert len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name), ast.dump(stmt)
            name = stmt.targets[0].id
            if name not in self.fields:  # a temporary: a 3-D field of the compute domain
                assert name not in self.scalars, f"assignment to the scalar parameter {name}"
                self.fields[name] = np.zeros(self.dom, dtype=self.float_dtype)
            target = self.fields[name]
            oi, oj, ok = self._origin.get(id(target), (0, 0, 0))
            values = {p: self._eval(stmt.value, self.fields, {}, *p) for p in self._points(ks, mask)}
            for (i, j, k), v in values.items():
                if target.ndim == 3:
                    target[i + oi, j + oj, k + ok] = v  # NumPy casts to the field's dtype, as gt4py does
                else:
                    target[i + oi, j + oj] = v
            return
        raise NotImplementedError(ast.dump(stmt))

    # -- expressions ---------------------------------------------------------------------------------------
    def _read(self, name, arr, i, j, k):
        oi, oj, ok = self._origin.get(id(arr), (0, 0, 0))
        i, j, k = i + oi, j + oj, k + ok
        if not (0 <= i < arr.shape[0] and 0 <= j < arr.shape[1] and (arr.ndim == 2 or 0 <= k < arr.shape[2])):
            raise UndefinedRead(f"read of {name} at storage index ({i},{j},{k}) outside {arr.shape}: undefined in gt4py")
        return arr[i, j, k] if arr.ndim == 3 else arr[i, j]

    def _eval(self, node, fields, local, i, j, k):
        ev = lambda n: self._eval(n, fields, local, i, j, k)  # noqa: E731
        if isinstance(node, ast.Constant):
            return node.value
        if isinstance(node, ast.Name):
            if node.id in local:
                return local[node.id]
            if node.id == "THIS_K":
                return k
            if node.id in self.scalars:
                return self.scalars[node.id]
            return self._read(node.id, fields[node.id], i, j, k)
        if isinstance(node, ast.Subscript):
            name = node.value.id
            offs = node.slice.elts if isinstance(node.slice, ast.Tuple) else [node.slice]
            di, dj, dk = (int(ev(o)) for o in offs)
            return self._read(name, fields[name], i + di, j + dj, k + dk)
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
        if isinstance(node, ast.IfExp):
            return ev(node.body) if ev(node.test) else ev(node.orelse)
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
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name):
            if node.func.id in _MATH:
                return _MATH[node.func.id](*[ev(a) for a in node.args])
            if node.func.id in self.functions:
                fn = self.functions[node.func.id]
                inner_fields, inner_local = dict(fields), {}
                for p, a in zip(fn.args.args, node.args):
                    if isinstance(a, ast.Name) and a.id in fields and a.id not in local:
                        inner_fields[p.arg] = fields[a.id]  # gt4py inlines the helper: the parameter IS the field
                    else:
                        inner_local[p.arg] = ev(a)
                try:
                    self._run_body(fn.body, inner_fields, inner_local, i, j, k)
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
