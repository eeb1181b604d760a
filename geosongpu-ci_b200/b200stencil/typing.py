"""Field-type markers (reference: ``ndsl.dsl.typing`` as used at WIP__hybrid_index_2dout.py:23).

``Float``/``Int`` follow the reference's precision switch ``PACE_FLOAT_PRECISION``
(/root/reference/src/tcn/ci/pipeline/gtfv3_config.py:11: 32 in production; the patterns run the 64-bit default).
"""
import os

import numpy as np

_PRECISION = int(os.getenv("PACE_FLOAT_PRECISION", "64"))
if _PRECISION not in (32, 64):
    raise RuntimeError(f"PACE_FLOAT_PRECISION must be 32 or 64, got {_PRECISION}")
Float = np.float64 if _PRECISION == 64 else np.float32
Int = np.int64 if _PRECISION == 64 else np.int32


class _Marker:
    def __init__(self, name, dtype, axes):
        self.name, self.dtype, self.axes = name, dtype, axes

    def __repr__(self):
        return self.name


FloatField = _Marker("FloatField", Float, "IJK")
FloatFieldIJ = _Marker("FloatFieldIJ", Float, "IJ")
FloatFieldK = _Marker("FloatFieldK", Float, "K")
IntField = _Marker("IntField", Int, "IJK")
IntFieldIJ = _Marker("IntFieldIJ", Int, "IJ")
BoolField = _Marker("BoolField", bool, "IJK")
