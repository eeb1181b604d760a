"""ctypes front-end of the C restatement (oracle/c/) -- TEST INFRASTRUCTURE ONLY.

Same call shapes as oracle/numpy_oracle.py, on arrays indexed ``[i, j, k]`` whose memory is
i-fastest (``oracle.inputs.ifirst_empty``).  ``build()`` runs oracle/Makefile; ``load(native=True)``
builds and loads the ``-march=native`` variant used only to time the CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

_I64, _INT = C.c_int64, C.c_int
_P = C.c_void_p


def build(native: bool = False) -> str:
    target = "native" if native else "all"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)
    return os.path.join(_HERE, "_build", "liboracle_native.so" if native else "liboracle.so")


def load(native: bool = False) -> C.CDLL:
    if native not in _LIBS:
        path = os.path.join(_HERE, "_build", "liboracle_native.so" if native else "liboracle.so")
        if native or not os.path.exists(path):
            path = build(native)
        _LIBS[native] = C.CDLL(path)
    return _LIBS[native]


def _suffix(a: np.ndarray) -> str:
    if a.dtype == np.float64:
        return "f64"
    if a.dtype == np.float32:
        return "f32"
    raise TypeError(f"unsupported dtype {a.dtype}")


def _f3(a: np.ndarray, origin=(0, 0, 0)):
    """(ptr, sj, sk) of the element ``origin`` of an i-fastest 3-D array."""
    it = a.itemsize
    assert a.ndim == 3 and (a.shape[0] == 1 or a.strides[0] == it), "field must be i-fastest"
    off = sum(o * s for o, s in zip(origin, a.strides))
    return _P(a.ctypes.data + off), _I64(a.strides[1] // it), _I64(a.strides[2] // it)


def _f2(a: np.ndarray):
    it = a.itemsize
    assert a.ndim == 2 and (a.shape[0] == 1 or a.strides[0] == it), "field must be i-fastest"
    return _P(a.ctypes.data), _I64(a.strides[1] // it)


def _real(a, v):
    return C.c_double(v) if a.dtype == np.float64 else C.c_float(v)


def _int_dtype(a):
    return np.int64 if a.dtype == np.float64 else np.int32


class COracle:
    def __init__(self, native: bool = False):
        self.lib = load(native)
        self.lib.orc_num_threads.restype = C.c_int

    @property
    def threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def set_threads(self, n: int) -> None:
        self.lib.orc_set_num_threads(C.c_int(n))

    def _fn(self, name, a, restype=None):
        f = getattr(self.lib, f"orc_{name}_{_suffix(a)}")
        f.restype = restype
        return f

    def top_of_column(self, PLEmb, PLEmb_top, out_field):
        ni, nj, nk = PLEmb.shape
        self._fn("top_of_column", PLEmb)(_INT(ni), _INT(nj), _INT(nk), *_f3(PLEmb), *_f2(PLEmb_top), *_f3(out_field))

    def while_in_function(self, in_field, out_field, threshold=4.0) -> int:
        ni, nj, nk = in_field.shape
        return int(
            self._fn("while_in_function", in_field, C.c_int64)(
                _INT(ni), _INT(nj), _INT(nk), _real(in_field, threshold), *_f3(in_field), *_f3(out_field)
            )
        )

    def hybrid_index_2dout(self, data_field, k_mask, k_index_desired, out_field):
        ni, nj, nk = data_field.shape
        self._fn("hybrid_index_2dout", data_field)(
            _INT(ni), _INT(nj), _INT(nk), *_f3(data_field), *_f3(k_mask), *_f2(k_index_desired), *_f2(out_field)
        )

    def find_klcl(self, PLmb, PLCL, KLCL, PLmb_at_KLCL):
        ni, nj, nk = PLmb.shape
        assert KLCL.dtype == _int_dtype(PLmb)
        self._fn("find_klcl", PLmb)(_INT(ni), _INT(nj), _INT(nk), *_f3(PLmb), *_f2(PLCL), *_f2(KLCL), *_f2(PLmb_at_KLCL))

    def saturation_adjust(self, T, q, ql, p):
        ni, nj, nk = T.shape
        self._fn("saturation_adjust", T)(_INT(ni), _INT(nj), _INT(nk), *_f3(T), *_f3(q), *_f3(ql), *_f3(p))

    def cloud_top(self, ql, ktop, ql_min=1.0e-8):
        ni, nj, nk = ql.shape
        assert ktop.dtype == _int_dtype(ql)
        self._fn("cloud_top", ql)(_INT(ni), _INT(nj), _INT(nk), _real(ql, ql_min), *_f3(ql), *_f2(ktop))

    def fv_tp2d(self, q, crx, xfx, cry, yfx, rarea, q_out):
        ni, nj, nk = q_out.shape
        self._fn("fv_tp2d", q)(
            _INT(ni), _INT(nj), _INT(nk), *_f3(q, (3, 3, 0)), *_f3(crx), *_f3(xfx), *_f3(cry), *_f3(yfx),
            *_f2(rarea), *_f3(q_out),
        )  # fmt: skip

    def fv_tp2d_split(self, q, crx, xfx, cry, yfx, area, rarea, q_out, fx_out=None, fy_out=None, corner_flags=0):
        from . import numpy_oracle

        ni, nj, nk = q_out.shape
        qy = q
        if corner_flags:  # what the inner y-sweep sees at cube corners (copy_corners direction 2)
            qy = np.empty(tuple(reversed(q.shape)), dtype=q.dtype).transpose()
            qy[...] = q
            numpy_oracle.copy_corners(qy, 2, corner_flags)
        null3 = (_P(None), _I64(0), _I64(0))
        a_it = area.itemsize
        assert area.ndim == 2 and area.strides[0] == a_it
        area_at = (_P(area.ctypes.data + 3 * area.strides[0] + 3 * area.strides[1]), _I64(area.strides[1] // a_it))
        self._fn("fv_tp2d_split", q)(
            _INT(ni), _INT(nj), _INT(nk), *_f3(q, (3, 3, 0)), *_f3(qy, (3, 3, 0)), *_f3(crx, (0, 3, 0)), *_f3(xfx, (0, 3, 0)),
            *_f3(cry, (3, 0, 0)), *_f3(yfx, (3, 0, 0)), *area_at, *_f2(rarea), *_f3(q_out),
            *(_f3(fx_out) if fx_out is not None else null3), *(_f3(fy_out) if fy_out is not None else null3),
        )  # fmt: skip

    def pe_prefix(self, delp, ptop, pe):
        ni, nj, nk = delp.shape
        self._fn("pe_prefix", delp)(_INT(ni), _INT(nj), _INT(nk), _real(delp, ptop), *_f3(delp), *_f3(pe))

    def remap(self, pe1, q1, pe2, q2):
        ni, nj, nk1 = q1.shape
        nk2 = q2.shape[2]
        self._fn("remap", q1)(_INT(ni), _INT(nj), _INT(nk1), _INT(nk2), *_f3(pe1), *_f3(q1), *_f3(pe2), *_f3(q2))

    def remap_ppm(self, pe1, q1, pe2, q2, kord=4, iv=1):
        ni, nj, nk1 = q1.shape
        nk2 = q2.shape[2]
        self._fn("remap_ppm", q1)(_INT(ni), _INT(nj), _INT(nk1), _INT(nk2), _INT(kord), _INT(iv), *_f3(pe1), *_f3(q1), *_f3(pe2), *_f3(q2))

    def tridiag(self, a, b, c, d, x, w=None):
        ni, nj, nk = b.shape
        if w is None:
            w = np.empty(tuple(reversed(b.shape)), dtype=b.dtype).transpose()
        self._fn("tridiag", b)(_INT(ni), _INT(nj), _INT(nk), *_f3(a), *_f3(b), *_f3(c), *_f3(d), *_f3(w), *_f3(x))

    def halo_move(self, links: np.ndarray, nk: int, src: np.ndarray, dst: np.ndarray):
        """Run a 10-word link table (halo/partitioner.py links as laid out by halo/updater.py HaloPlan.tables) on flat
        host storage: the CPU restatement of the CUDA halo_move kernel."""
        links = np.ascontiguousarray(links, dtype=np.int64).reshape(-1, 10)
        assert src.dtype == dst.dtype and src.ndim == 1 and dst.ndim == 1
        self._fn("halo_move", src)(_INT(len(links)), _INT(nk), _P(links.ctypes.data), _P(src.ctypes.data), _P(dst.ctypes.data))
