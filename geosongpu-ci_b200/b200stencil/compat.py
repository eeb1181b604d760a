"""``install()``: make ``gt4py.cartesian.gtscript`` / ``ndsl.*`` importable as aliases of this package
ONLY when the real packages are absent, so the reference pattern files run byte-for-byte unmodified
(SURVEY.md Appendix B).  If gt4py/ndsl are installed they win and nothing is touched.
"""
from __future__ import annotations

import importlib.util
import sys
import types


def _absent(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install(force: bool = False) -> bool:
    from . import api, constants, gtscript, typing

    if not force and not (_absent("gt4py") and _absent("ndsl")):
        return False

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__b200stencil_alias__ = True
        sys.modules[name] = m
        return m

    gt4py = mod("gt4py")
    cart = mod("gt4py.cartesian")
    sys.modules["gt4py.cartesian.gtscript"] = gtscript
    gt4py.cartesian, cart.gtscript = cart, gtscript

    ndsl = mod("ndsl", StencilFactory=api.StencilFactory, QuantityFactory=api.QuantityFactory,
               Quantity=api.Quantity, orchestrate=api.orchestrate)  # fmt: skip
    ndsl.boilerplate = mod("ndsl.boilerplate", get_factories_single_tile_numpy=api.get_factories_single_tile_numpy,
                           get_factories_single_tile=api.get_factories_single_tile)  # fmt: skip
    sys.modules["ndsl.constants"] = constants
    ndsl.constants = constants
    dsl = mod("ndsl.dsl")
    sys.modules["ndsl.dsl.typing"] = typing
    dsl.typing, ndsl.dsl = typing, dsl
    return True


def uninstall() -> None:
    for name in list(sys.modules):
        m = sys.modules[name]
        if name.split(".")[0] in ("gt4py", "ndsl") and (
            getattr(m, "__b200stencil_alias__", False) or getattr(m, "__name__", "").startswith("b200stencil")
        ):
            del sys.modules[name]
