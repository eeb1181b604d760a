"""Stencil definition -> hand-written kernel ("dispatch key that is not the name").

All three reference patterns name their gtscript function ``stencil``
(Do__get_top_of_the_column.py:33, Do__while_in_gt_functions.py:30, WIP__hybrid_index_2dout.py:34), so the
key is a hash of the normalised AST of the definition (annotations, decorators and docstrings dropped)
plus the ``@function`` helpers it calls.  Unknown definitions raise -- there is no CPU fallback and no
DSL compiler.  ``register()`` adds user kernels; ``kernel=`` on the factory call overrides the lookup.
"""
from __future__ import annotations

import ast
import hashlib
import inspect
import textwrap
from typing import Callable, Dict, Optional


class NoKernelError(LookupError):
    pass


def _normalised(fn) -> str:
    tree = ast.parse(textwrap.dedent(inspect.getsource(fn)))
    node = tree.body[0]
    assert isinstance(node, ast.FunctionDef)
    node.decorator_list = []
    node.returns = None
    node.name = "_"
    for a in node.args.args + node.args.kwonlyargs:
        a.annotation = None
    if node.body and isinstance(node.body[0], ast.Expr) and isinstance(getattr(node.body[0], "value", None), ast.Constant) \
            and isinstance(node.body[0].value.value, str):
        node.body = node.body[1:]
    return _canonical(node)


_SKIPPED_FIELDS = {"ctx", "type_comment", "kind", "type_params", "type_ignores"}


def _canonical(node) -> str:
    """A dump of the tree that does not depend on the interpreter version: ``ast.dump`` prints empty and ``None`` fields
    up to Python 3.12 and omits them from 3.13 on, and 3.8 wraps subscripts in ``Index``.  Here a node is its class name
    and its non-empty fields by name; load/store contexts and the fields newer grammars added are left out."""
    if isinstance(node, ast.AST):
        if type(node).__name__ == "Index":  # Python 3.8
            return _canonical(node.value)
        parts = []
        for f in node._fields:
            if f in _SKIPPED_FIELDS:
                continue
            v = getattr(node, f, None)
            if v is None or v == []:
                continue
            parts.append(f"{f}={_canonical(v)}")
        return f"{type(node).__name__}({','.join(parts)})"
    if isinstance(node, list):
        return "[" + ",".join(_canonical(x) for x in node) + "]"
    return repr(node)


def definition_key(fn) -> str:
    """sha256 of the definition and of every module-level ``@function`` helper it calls (by name order)."""
    parts = [_normalised(fn)]
    tree = ast.parse(textwrap.dedent(inspect.getsource(fn)))
    called = sorted({n.func.id for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name)})
    for name in called:
        helper = getattr(fn, "__globals__", {}).get(name)
        if helper is not None and getattr(helper, "__gtscript_function__", False):
            parts.append(_normalised(helper))
    return hashlib.sha256("\n".join(parts).encode()).hexdigest()[:16]


# key -> (kernel name, adapter(*call args) running the kernel on device fields)
_BY_KEY: Dict[str, str] = {}
_ADAPTERS: Dict[str, Callable] = {}


def register(name: str, adapter: Callable, *keys: str) -> None:
    _ADAPTERS[name] = adapter
    for k in keys:
        _BY_KEY[k] = name


def resolve(fn, kernel: Optional[str] = None) -> str:
    if kernel is not None:
        if kernel not in _ADAPTERS:
            raise NoKernelError(f"no hand-written sm_100a kernel named {kernel!r}; known: {sorted(_ADAPTERS)}")
        return kernel
    tagged = getattr(fn, "__b200_kernel__", None)
    if tagged:
        return resolve(fn, tagged)
    key = definition_key(fn)
    if key not in _BY_KEY:
        raise NoKernelError(
            f"no hand-written sm_100a kernel registered for stencil definition {fn.__module__}.{fn.__qualname__} "
            f"(key {key}). b200stencil does not compile DSL code and has no CPU fallback: pass kernel='<name>' "
            f"(known: {sorted(_ADAPTERS)}) or register one with b200stencil.registry.register()."
        )
    return _BY_KEY[key]


def adapter(name: str) -> Callable:
    return _ADAPTERS[name]


def kernel(name: str):
    """Decorator tagging a stencil definition with the kernel that implements it."""

    def deco(fn):
        fn.__b200_kernel__ = name
        return fn

    return deco


def _install_builtin() -> None:
    from . import stencils

    # keys computed from the reference definitions (see tests/test_api_shim.py::test_reference_files_resolve)
    register("top_of_column", lambda PLEmb, PLEmb_top, out_field: stencils.top_of_column(PLEmb, PLEmb_top, out_field),
             KEY_TOP_OF_COLUMN)
    register("while_in_function", lambda in_field, out_field: stencils.while_in_function(in_field, out_field),
             KEY_WHILE_IN_FUNCTION)
    register("hybrid_index_2dout",
             lambda data_field, k_mask, k_index_desired, out_field: stencils.hybrid_index_2dout(
                 data_field, k_mask, k_index_desired, out_field), KEY_HYBRID_INDEX)
    register("find_klcl", stencils.find_klcl)
    register("saturation_adjust", stencils.saturation_adjust)
    register("cloud_top", stencils.cloud_top)
    register("fv_tp2d", stencils.fv_tp2d)
    register("pe_prefix", stencils.pe_prefix)
    register("fv_tp2d_split", stencils.fv_tp2d_split)
    register("remap", stencils.remap)
    register("remap_delp", stencils.remap_delp)
    register("remap_ppm", stencils.remap_ppm)
    register("tridiag", stencils.tridiag)


# normalised-AST keys of the three dsl_patterns definitions
KEY_TOP_OF_COLUMN = "4171b7d0e0c22cb6"
KEY_WHILE_IN_FUNCTION = "32099175f9960450"
KEY_HYBRID_INDEX = "6e87534ceae78890"

_install_builtin()
