"""-m gpu: the reference's three pattern files run UNMODIFIED on the device, ``__main__`` block included
(fixtures: tests/golden/ref_patterns, byte-for-byte /root/reference/dsl_patterns/*.py).

Pass criteria are the files' own asserts (Do__get_top_of_the_column.py:68, Do__while_in_gt_functions.py:62); the WIP file
only prints (WIP__hybrid_index_2dout.py:68-90), so its output is compared with a gather at the desired level."""
import os
import runpy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from b200stencil import compat  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_patterns")


@pytest.fixture
def aliases():
    installed = compat.install()
    yield installed
    if installed:
        compat.uninstall()


def _run_main(name):
    return runpy.run_path(os.path.join(HERE, name), run_name="__main__")


def test_top_of_column_file(aliases, capsys):
    ns = _run_main("Do__get_top_of_the_column.py")  # asserts np.all(O == 42) itself
    assert ns["code"].stencil.kernel_name == "top_of_column"
    assert np.all(ns["O"] == 42) and ns["O"].shape == (3, 3, 4)
    assert np.all(ns["code"]._tmp.view[:, :].cpu().numpy() == 42)
    assert "Output:" in capsys.readouterr().out


def test_while_in_function_file(aliases, capsys):
    ns = _run_main("Do__while_in_gt_functions.py")  # asserts O[0, 0, :] == [3, 2, 1, 0] itself
    assert ns["code"].stencil.kernel_name == "while_in_function"
    assert (ns["O"] == np.array([3.0, 2.0, 1.0, 0.0])).all()  # every column, not only (0, 0)
    assert "Output:" in capsys.readouterr().out


@pytest.mark.parametrize("seed", [0, 1, 20240724])
def test_hybrid_index_file(aliases, capsys, seed):
    np.random.seed(seed)
    ns = _run_main("WIP__hybrid_index_2dout.py")
    assert ns["code"].stencil.kernel_name == "hybrid_index_2dout"
    out = ns["code"].O.view[:, :].cpu().numpy()
    want_k = ns["k_index_desired"].view[:, :].cpu().numpy().astype(np.int64)
    data = ns["input_to_sample_from"].view[:, :, :].cpu().numpy()
    ii, jj = np.meshgrid(np.arange(data.shape[0]), np.arange(data.shape[1]), indexing="ij")
    assert np.array_equal(out, data[ii, jj, want_k])
    assert 800 <= out.min() and out.max() < 900
    assert "K Level Desired for each column" in capsys.readouterr().out
