"""The reference's pattern programs as committed fixtures (tests/golden/ref_patterns): still the reference's bytes, and --
without a GPU -- each file's ``__main__`` runs unmodified up to its stencil call, which fails loudly (no CPU fallback)."""
import hashlib
import os
import runpy

import pytest

from b200stencil import compat, registry

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_patterns")
REF = "/root/reference/dsl_patterns"
FILES = {
    "Do__get_top_of_the_column.py": "top_of_column",
    "Do__while_in_gt_functions.py": "while_in_function",
    "WIP__hybrid_index_2dout.py": "hybrid_index_2dout",
}


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


@pytest.fixture
def aliases():
    installed = compat.install()
    yield installed
    if installed:
        compat.uninstall()


def test_fixtures_match_their_checksums():
    with open(os.path.join(HERE, "SHA256SUMS")) as f:
        sums = dict(reversed(line.split()) for line in f if line.strip())
    assert set(sums) == set(FILES)
    for name, want in sums.items():
        assert _sha(os.path.join(HERE, name)) == want, name


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", sorted(FILES))
def test_fixtures_are_the_reference_bytes(name):
    assert _sha(os.path.join(HERE, name)) == _sha(os.path.join(REF, name))


@pytest.mark.parametrize("name,kernel", sorted(FILES.items()))
def test_fixture_resolves_to_its_kernel(aliases, name, kernel):
    """Same check as tests/test_api_shim.py::test_reference_files_resolve, on the copy that travels to the GPU box."""
    ns = runpy.run_path(os.path.join(HERE, name), run_name="loaded_by_test")
    assert registry.resolve(ns["stencil"]) == kernel
    assert ns["Code"](ns["stcil_fctry"], ns["ijk_qty_fctry"]).stencil.kernel_name == kernel


@pytest.mark.parametrize("name,kernel", sorted(FILES.items()))
def test_main_block_reaches_the_stencil_and_refuses_the_cpu(aliases, name, kernel, capsys):
    import torch

    if torch.cuda.is_available():
        pytest.skip("covered on the device by tests/test_gpu_ref_patterns.py")
    with pytest.raises(RuntimeError) as e:
        runpy.run_path(os.path.join(HERE, name), run_name="__main__")
    assert f"stencil '{kernel}'" in str(e.value) and "no CPU fallback" in str(e.value)
