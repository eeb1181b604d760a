"""Host API shim: the reference pattern files import and resolve to kernels
unmodified; Quantity layout; loud failure without a GPU.  No GPU needed."""
import os
import runpy
import sys

import numpy as np
import pytest
import torch

import b200stencil
from b200stencil import compat, registry
from b200stencil.api import get_factories_single_tile_numpy
from b200stencil.constants import X_DIM, Y_DIM, Z_DIM, Z_INTERFACE_DIM
from b200stencil.gtscript import FORWARD, PARALLEL, computation, function, interval
from b200stencil.typing import FloatField, FloatFieldIJ

REF = "/root/reference/dsl_patterns"


@pytest.fixture
def aliases():
    installed = compat.install()
    yield installed
    if installed:
        compat.uninstall()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize(
    "fname,kernel",
    [("Do__get_top_of_the_column.py", "top_of_column"), ("Do__while_in_gt_functions.py", "while_in_function"),
     ("WIP__hybrid_index_2dout.py", "hybrid_index_2dout")],
)  # fmt: skip
def test_reference_files_resolve(aliases, fname, kernel):
    """The unmodified reference file imports through the aliases, builds its Code object and its
    ``stencil`` definition resolves to the matching hand-written kernel."""
    assert aliases, "gt4py/ndsl unexpectedly installed: the shim must stand aside"
    ns = runpy.run_path(os.path.join(REF, fname), run_name="loaded_by_test")
    assert registry.resolve(ns["stencil"]) == kernel
    code = ns["Code"](ns["stcil_fctry"], ns["ijk_qty_fctry"])
    assert code.stencil.kernel_name == kernel
    assert code.stencil.domain == (3, 3, 4) and code.stencil.origin == (0, 0, 0)


def test_aliases_do_not_shadow_real_packages(monkeypatch):
    fake = type(sys)("gt4py")
    monkeypatch.setitem(sys.modules, "gt4py", fake)
    assert compat.install() is False
    assert sys.modules["gt4py"] is fake


def test_same_definition_different_name_resolves():
    def anything(a: FloatField, b: FloatFieldIJ, c: FloatField):
        """docstrings and annotations do not change the key"""
        with computation(FORWARD), interval(-1, None):
            b = a
        with computation(PARALLEL), interval(...):
            c = b

    # argument NAMES are part of the definition: this one differs from the reference pattern
    with pytest.raises(registry.NoKernelError) as e:
        registry.resolve(anything)
    assert "no hand-written sm_100a kernel" in str(e.value)
    assert registry.resolve(anything, kernel="top_of_column") == "top_of_column"

    def stencil(PLEmb, PLEmb_top, out_field):
        with computation(FORWARD), interval(-1, None):
            PLEmb_top = PLEmb
        with computation(PARALLEL), interval(...):
            out_field = PLEmb_top

    assert registry.resolve(stencil) == "top_of_column"

    @registry.kernel("fv_tp2d")
    def tagged(q):
        pass

    assert registry.resolve(tagged) == "fv_tp2d"


def test_quantity_layout_and_view():
    sf, qf = get_factories_single_tile_numpy(5, 4, 3, 2, device="cpu")
    q = qf.zeros([X_DIM, Y_DIM, Z_DIM], "n/a")
    assert q.shape == (5 + 4 + 1, 4 + 4 + 1, 3 + 1)  # NDSL padding: (nx+2h+1, ny+2h+1, nz+1)
    assert q.data.stride(0) == 1, "i-fastest storage"
    assert q.data.stride(1) % 2 == 0, "rows padded to 16 bytes"
    assert tuple(q.view[:, :, :].shape) == (5, 4, 3)
    q.view[:, :, :] = np.arange(60, dtype=np.float64).reshape(5, 4, 3)
    assert float(q.data[2, 2, 0]) == 0.0 and float(q.data[6, 5, 2]) == 59.0
    assert float(q.data.sum()) == float(np.arange(60).sum())  # halo untouched
    q.view[1, 2, 0] = 7.0
    assert float(q.view[1, 2, 0]) == 7.0
    assert "59." in repr(q.view)
    qi = qf.zeros([X_DIM, Y_DIM, Z_INTERFACE_DIM], "Pa")
    assert tuple(qi.view[:].shape) == (5, 4, 4)
    ij = qf.zeros([X_DIM, Y_DIM], "n/a", dtype=np.float32)
    assert ij.dtype == torch.float32 and tuple(ij.view[:, :].shape) == (5, 4)
    assert sf.config.dace_config is not None and b200stencil.orchestrate(obj=object(), config=sf.config.dace_config) is None


def test_no_cpu_fallback():
    sf, qf = get_factories_single_tile_numpy(3, 3, 4, 0, device="cpu")

    def stencil(PLEmb, PLEmb_top, out_field):
        with computation(FORWARD), interval(-1, None):
            PLEmb_top = PLEmb
        with computation(PARALLEL), interval(...):
            out_field = PLEmb_top

    st = sf.from_dims_halo(func=stencil, compute_dims=[X_DIM, Y_DIM, Z_DIM])
    with pytest.raises(RuntimeError) as e:
        st(np.ones((3, 3, 4)), qf.zeros([X_DIM, Y_DIM], "n/a"), np.zeros((3, 3, 4)))
    assert "no CPU fallback" in str(e.value)
