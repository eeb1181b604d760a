"""A compiled C program calls libb200stencil through include/b200stencil.h with no Python in the process
(tests/c_abi/abi_driver.c) -- the counterpart of the reference's acceptance program for its generated bridge
(test/py_ftn_interface/data/fortran_program.f90:24-36).  Without a GPU the program must fail loudly at
b2s_init (no CPU fallback); on the B200 it must reproduce the reference's golden vectors."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_driver.c")
LIBDIR = os.path.join(ROOT, "geosongpu-ci_b200", "b200stencil", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _compile(tmp_path_factory, name, extra=()):
    src = os.path.join(ROOT, "tests", "c_abi", name + ".c")
    exe = str(tmp_path_factory.mktemp("c_abi") / name)
    cmd = ["gcc", "-O1", "-Wall", "-Wextra", "-Werror", src, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", LIBDIR, "-lb200stencil", "-L", os.path.join(CUDA, "lib64"), "-lcudart", *extra, "-o", exe]  # fmt: skip
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.pathsep.join([LIBDIR, os.path.join(CUDA, "lib64"), os.environ.get("LD_LIBRARY_PATH", "")]))
    return exe, env


@pytest.fixture(scope="module")
def halo_driver(tmp_path_factory):
    """tests/c_abi/halo_driver.c: the b2s_halo_* lifecycle with two ranks (threads) and no Python in the process."""
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    return _compile(tmp_path_factory, "halo_driver", ("-lpthread",))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    if not os.path.exists(os.path.join(LIBDIR, "libb200stencil.so")):
        pytest.fail("libb200stencil.so is not built: run python __graft_entry__.py")
    exe = str(tmp_path_factory.mktemp("c_abi") / "abi_driver")
    cmd = ["gcc", "-O1", "-Wall", "-Wextra", "-Werror", SRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", LIBDIR, "-lb200stencil", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-o", exe]  # fmt: skip
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.pathsep.join([LIBDIR, os.path.join(CUDA, "lib64"), os.environ.get("LD_LIBRARY_PATH", "")]))
    return exe, env


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="this is the no-GPU behaviour")
def test_compiled_caller_fails_loudly_without_a_gpu(driver):
    exe, env = driver
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "b2s_init" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_compiled_caller_reproduces_the_golden_vectors(driver):
    exe, env = driver
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi_driver ok" in r.stdout


@pytest.mark.skipif(_has_cuda(), reason="this is the no-GPU behaviour")
def test_halo_driver_fails_loudly_without_a_gpu(halo_driver):
    exe, env = halo_driver
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "b2s_init" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_halo_lifecycle_from_c(halo_driver):
    """Two ranks, symmetric allocation, plain and forked exchanges, gated stencil, finalize -- all through the C-ABI."""
    exe, env = halo_driver
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "halo_driver ok" in r.stdout


@pytest.fixture(scope="module")
def fortran_mirror(tmp_path_factory):
    """tests/c_abi/fortran_mirror.c: the call sequence of include/b2s_example.f90 in C (no Fortran compiler here)."""
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    return _compile(tmp_path_factory, "fortran_mirror")


def test_fortran_mirror_follows_the_fortran_program():
    """The mirror makes the library calls of the generated Fortran program, in the same order."""
    import re

    from b200stencil.bridge import generate

    f90 = generate.Bridge.from_yaml().emit_fortran_example()
    c = open(os.path.join(ROOT, "tests", "c_abi", "fortran_mirror.c")).read()
    body = f90[f90.index("implicit none"):]
    calls_f = re.findall(r"\b(b2s_\w+)\s*\(", body)
    calls_c = [m for m in re.findall(r"\b(b2s_\w+)\s*\(", c[c.index("int main"):]) if m != "b2s_last_error"]
    assert calls_f == calls_c, (calls_f, calls_c)
    for n in re.findall(r"stop (\d+)", f90):
        if n == "1":
            continue  # b2s_init failure: the C drivers all exit with 3 there (the no-GPU test keys on it)
        assert f"STOP({n}," in c or f"return {n};" in c, f"stop {n} of the Fortran program has no counterpart"


@pytest.mark.skipif(_has_cuda(), reason="this is the no-GPU behaviour")
def test_fortran_mirror_fails_loudly_without_a_gpu(fortran_mirror):
    exe, env = fortran_mirror
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "b2s_init" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_fortran_call_sequence_on_the_device(fortran_mirror):
    """init, halo context, symmetric allocations, plan, exchange, status, fv_tp2d, finalize -- as the Fortran program does."""
    exe, env = fortran_mirror
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "b2s_example: ok" in r.stdout
