"""-m gpu: user-level programs written against the reference's call surface (factories, Quantity, Code
classes with the reference signatures), run end to end on the device; the reference's own asserts
(Do__get_top_of_the_column.py:68, Do__while_in_gt_functions.py:62) are the pass criteria."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from b200stencil import QuantityFactory, StencilFactory, get_factories_single_tile_numpy, orchestrate  # noqa: E402
from b200stencil.constants import X_DIM, Y_DIM, Z_DIM  # noqa: E402
from b200stencil.gtscript import FORWARD, PARALLEL, computation, function, interval  # noqa: E402
from b200stencil.typing import FloatField, FloatFieldIJ, IntField, IntFieldIJ  # noqa: E402

domain = (3, 3, 4)


class _Harness:
    """What every pattern file does at import: build its factories, then a Code object around one stencil."""

    def __init__(self, stencil_def, dom=domain, nhalo=0):
        self.sf, self.qf = get_factories_single_tile_numpy(dom[0], dom[1], dom[2], nhalo)
        orchestrate(obj=self, config=self.sf.config.dace_config)
        self.stencil = self.sf.from_dims_halo(func=stencil_def, compute_dims=[X_DIM, Y_DIM, Z_DIM])


def test_top_of_column_program():
    def stencil(PLEmb: FloatField, PLEmb_top: FloatFieldIJ, out_field: FloatField):
        with computation(FORWARD), interval(-1, None):
            PLEmb_top = PLEmb
        with computation(PARALLEL), interval(...):
            out_field = PLEmb_top

    h = _Harness(stencil)
    tmp = h.qf.zeros([X_DIM, Y_DIM], "n/a")
    I = np.ones(domain[0] * domain[1] * domain[2], dtype=np.float64).reshape(domain)
    I[:, :, domain[2] - 1] = 42
    O = np.zeros(domain)
    h.stencil(I, tmp, O)  # raw NumPy arrays of the compute-domain shape, as the reference passes
    assert np.all(O == 42)
    assert np.all(tmp.view[:, :].cpu().numpy() == 42)


@function
def while_in_function(field: FloatField):
    lev = 0
    while field[0, 0, lev] < 4:
        lev += 1
    return lev


def test_while_in_function_program():
    def stencil(in_field: FloatField, out_field: FloatField):
        with computation(PARALLEL), interval(...):
            out_field = while_in_function(in_field)

    h = _Harness(stencil)
    I = np.ones(domain, dtype=np.float64)
    I[:, :, domain[2] - 1] = 42
    O = np.zeros(domain)
    h.stencil(I, O)
    assert (O[0, 0, :] == [3.0, 2.0, 1.0, 0.0]).all()


@pytest.mark.parametrize("dom,nhalo", [((3, 3, 4), 0), ((9, 7, 11), 3)])
def test_hybrid_index_program_with_quantities(dom, nhalo):
    def stencil(data_field: FloatField, k_mask: FloatField, k_index_desired: FloatFieldIJ, out_field: FloatFieldIJ):
        with computation(FORWARD), interval(...):
            if k_mask == k_index_desired:
                out_field = data_field

    h = _Harness(stencil, dom, nhalo)
    rng = np.random.default_rng(3)
    k_mask = h.qf.zeros([X_DIM, Y_DIM, Z_DIM], "n/a")
    k_index = h.qf.zeros([X_DIM, Y_DIM], "n/a")
    data = h.qf.zeros([X_DIM, Y_DIM, Z_DIM], "n/a")
    out = h.qf.zeros([X_DIM, Y_DIM], "n/a")
    want_k = rng.integers(0, dom[2], size=dom[:2])
    k_index.view[:, :] = want_k.astype(np.float64)
    vals = rng.integers(800, 900, size=dom).astype(np.float64)
    data.view[:, :, :] = vals
    for k in range(dom[2]):
        k_mask.view[:, :, k] = float(k)
    h.stencil(data, k_mask, k_index, out)
    ii, jj = np.meshgrid(np.arange(dom[0]), np.arange(dom[1]), indexing="ij")
    assert np.array_equal(out.view[:, :].cpu().numpy(), vals[ii, jj, want_k])
    # halo cells of the output stay untouched (writes land inside the compute domain only)
    assert float(out.data.sum()) == float(vals[ii, jj, want_k].sum())
