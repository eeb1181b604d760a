"""Import-time stand-ins for ``gt4py.cartesian.gtscript``.

The reference patterns import ``computation, interval, PARALLEL, FORWARD, function`` at module import
(/root/reference/dsl_patterns/Do__get_top_of_the_column.py:19, Do__while_in_gt_functions.py:8,
WIP__hybrid_index_2dout.py:20) and decorate helpers with ``@function``; the stencil BODIES are never
executed by Python -- here they are never compiled either: a stencil definition is a dispatch key to a
hand-written sm_100a kernel (registry.py).
"""
from __future__ import annotations

PARALLEL, FORWARD, BACKWARD = "PARALLEL", "FORWARD", "BACKWARD"
THIS_K = "THIS_K"  # level index of the point (gt4py >= 1.0.4 experimental builtin; older stacks pass a k-index field)
I, J, K = "I", "J", "K"
IJ, IK, JK, IJK = "IJ", "IK", "JK", "IJK"


class _Ctx:
    def __init__(self, *a, **k):
        self.args = a

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def computation(order):
    return _Ctx(order)


def interval(*bounds):
    return _Ctx(*bounds)


def horizontal(*regions):
    return _Ctx(*regions)


def function(fn):
    """``@gtscript.function``: marks a helper inlined into stencils; identity here."""
    fn.__gtscript_function__ = True
    return fn


def stencil(backend=None, definition=None, **kwargs):
    raise NotImplementedError(
        "b200stencil has no DSL compiler: build stencils with StencilFactory.from_dims_halo / from_origin_domain, "
        "which dispatch a stencil definition to its hand-written sm_100a kernel"
    )
