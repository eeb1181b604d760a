"""pytest configuration: markers and import paths.

The product package lives under ``geosongpu-ci_b200/`` (not an importable name), so that
directory goes on sys.path and the host package is imported as ``b200stencil``.
"""
import os
import sys

import pytest

# The halo tests run several "virtual ranks" (threads) on ONE GPU, whose kernels wait for each other on the device.
# With the default 8 hardware work queues two independent streams can share a queue, and a launch of one rank then sits
# behind a pending wait of another -- a false dependency that the bounded device-side waits turn into a 2 s stall and a
# status flag.  One queue per stream removes it.  (One rank per GPU -- the product configuration -- never shares queues
# between ranks.)  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "geosongpu-ci_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


# Tests in which several ranks share ONE GPU and wait for each other ON THE DEVICE (virtual ranks: threads, or the
# compiled C callers).  They depend on every rank's kernels being co-resident -- a property of the test set-up, not of the
# product (one rank per GPU) -- so they run after everything else: the driver runs the suite with -x, and a stall there
# must not keep the parity tests from being run and counted.
_DEVICE_WAIT_TESTS = ("test_gpu_halo_device.py", "test_c_abi_driver.py::test_halo_lifecycle_from_c",
                      "test_c_abi_driver.py::test_fortran_call_sequence_on_the_device")


def pytest_collection_modifyitems(config, items):
    last = [it for it in items if any(tag in it.nodeid for tag in _DEVICE_WAIT_TESTS)]
    if last:
        keep = [it for it in items if not any(tag in it.nodeid for tag in _DEVICE_WAIT_TESTS)]
        items[:] = keep + last
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """-m gpu runs: the pointwise relative errors every assert_close saw (tests/gpu_util.py), for profiles/."""
    mod = sys.modules.get("gpu_util")
    rows = getattr(mod, "POINTWISE", None) if mod else None
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not rows or not os.path.isdir(out_dir):
        return
    import json

    try:
        with open(os.path.join(out_dir, "parity_pointwise.jsonl"), "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    except OSError:
        pass
