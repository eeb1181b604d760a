"""b200stencil -- B200-native stencil hot path behind the reference's stencil call surface.

Python host layer (this package) -> C-ABI (include/b200stencil.h, libb200stencil.so) -> hand-written
sm_100a kernels (../csrc).  No DSL compiler, no backend dispatch, no CPU fallback.
"""
from .api import (FrozenStencil, Quantity, QuantityFactory, StencilFactory, get_factories_single_tile,  # noqa: F401
                  get_factories_single_tile_numpy, orchestrate)
from .constants import X_DIM, X_INTERFACE_DIM, Y_DIM, Y_INTERFACE_DIM, Z_DIM, Z_INTERFACE_DIM  # noqa: F401

__version__ = "0.1.0"
