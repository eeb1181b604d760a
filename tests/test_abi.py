"""The C-ABI boundary without a GPU: the generated header is current, the library loads and exports
every symbol the header declares, the generator follows the reference's schema, errors are loud."""
import ctypes
import os
import re
import subprocess

import pytest

from b200stencil import _abi
from b200stencil.bridge import generate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200stencil.h")


@pytest.fixture(scope="module")
def bridge():
    return generate.Bridge.from_yaml()


def test_header_is_generated_from_the_yaml(bridge):
    with open(HEADER) as f:
        assert f.read() == bridge.emit_header(), "include/b200stencil.h is stale: run python -m b200stencil.bridge.generate"
    with open(generate.DEFAULT_GLUE) as f:
        assert f.read() == bridge.emit_glue()


def test_library_exports_every_declared_symbol(bridge):
    """dlopen works without a device and every prototype of the header resolves (no compute calls)."""
    assert os.path.exists(_abi.LIB_PATH), "build first: python __graft_entry__.py"
    lib = ctypes.CDLL(_abi.LIB_PATH)
    with open(HEADER) as f:
        declared = re.findall(r"^B2S_API\s+[\w\s\*]+?\b(b2s_\w+)\(", f.read(), flags=re.M)
    assert len(declared) >= 30 and set(declared) == set(bridge.symbols())
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/b200stencil.h but not exported"
    # and nothing else leaks out of the library (visibility=hidden by default)
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(declared)


def test_abi_version_and_error_channel_without_gpu():
    ffi, lib = _abi.load()
    assert lib.b2s_abi_version() == 1
    import torch

    if not torch.cuda.is_available():
        rc = lib.b2s_init(0)
        assert rc != 0
        msg = _abi.last_error()
        assert "no CPU fallback" in msg or "CUDA" in msg
        with pytest.raises(_abi.B200StencilError):
            _abi.init(0)
    assert lib.b2s_set_option(b"fv_variant", 1) == 0 and lib.b2s_get_option(b"fv_variant") == 1
    assert lib.b2s_set_option(b"fv_variant", 0) == 0
    assert lib.b2s_set_option(b"no_such_option", 1) == -1
    with pytest.raises(KeyError):
        _abi.set_option("no_such_option", 1)


def test_schema_follows_the_reference_generator(bridge, tmp_path):
    """Same YAML keys / argument order / symbol naming as tcn-fpy
    (/root/reference/src/tcn/py_ftn_interface/argument.py, base.py:35-36, interface.c.jinja2:8)."""
    fn = bridge.functions["hybrid_index_2dout"]
    assert [a.name for a in fn.arguments] == ["ni", "nj", "nk", "nb", "data_field", "k_mask", "k_index_desired", "out_field"]
    assert fn.symbol("b2s", "double") == "b2s_hybrid_index_2dout_c"
    assert fn.symbol("b2s", "float") == "b2s_hybrid_index_2dout_f32_c"
    proto = fn.c_prototype("b2s", "float")
    assert "const float* data_field, int64_t data_field_sj, int64_t data_field_sk, int64_t data_field_sb" in proto
    assert "float* out_field, int64_t out_field_sj, int64_t out_field_sb, void* stream" in proto
    assert generate.Argument("is", "int").name_sanitize == "_is"  # reserved names, argument.py:17-20
    # the reference's own test definition parses with this loader (MPI argument, 'arguments: None')
    y = tmp_path / "ref.yaml"
    y.write_text(
        "type: py_ftn_interface\nname: py_ftn_test\nbridge:\n"
        "    - name: check_mpi_translation\n      arguments:\n        inputs:\n"
        "            - !Argument\n              name: comm\n              type: MPI\n"
        "    - name: check_data\n      arguments:\n        inputs:\n"
        "            - !Argument\n              name: scalar\n              type: int\n"
        "            - !Argument\n              name: in_array\n              type: array_float\n              dims: 2\n"
        "        outputs:\n            - !Argument\n              name: out_array\n              type: array_float\n              dims: 2\n"
        "    - name: check_empty_function\n      arguments: None\n"
    )
    ref = generate.Bridge.from_yaml(str(y))
    assert list(ref.functions) == ["check_mpi_translation", "check_data", "check_empty_function"]
    assert ref.functions["check_data"].c_prototype("py_ftn_test", "double").startswith("int py_ftn_test_check_data_c(int scalar, const float* in_array")
    assert ref.functions["check_empty_function"].arguments == []
    with pytest.raises(RuntimeError):
        bad = tmp_path / "bad.yaml"
        bad.write_text("type: something_else\nname: x\nbridge: []\n")
        generate.Bridge.from_yaml(str(bad))


def test_fortran_interface_module(bridge):
    """The generated bind(c) module parses as free-form Fortran (tests/fortran_check.py: this image has no Fortran
    compiler) and every interface is a correct interoperable declaration of the C prototype of the same name in
    include/b200stencil.h -- the check the reference gets by compiling test/py_ftn_interface/data/fortran_program.f90
    against templates/interface.f90.jinja2:26-55."""
    import fortran_check as fc

    f90 = bridge.emit_fortran()
    mod = fc.parse_module(f90)
    assert mod.name == "b2s_interface_mod"
    header = open(os.path.join(ROOT, "include", "b200stencil.h")).read()
    n = fc.check_against_header(mod, header)
    assert n == len(bridge.symbols()) == len(mod.interfaces)
    assert max(len(l) for l in f90.splitlines()) <= 132
    # committed copies are what the generator produces now
    assert open(os.path.join(ROOT, "include", "b2s_interface_mod.f90")).read() == f90
    assert open(os.path.join(ROOT, "include", "b2s_example.f90")).read() == bridge.emit_fortran_example()
    # spot checks of the mapping
    halo_init = mod.interfaces["b2s_halo_init"]
    assert halo_init.decls["session"].base == "character" and halo_init.decls["session"].shape == "*"
    assert halo_init.decls["ctx"].attrs == ["intent(out)"] and halo_init.decls["ctx"].kind == "c_int64_t"
    fv = mod.interfaces["b2s_fv_tp2d_f32_c"]
    assert fv.dummies[:8] == ["ni", "nj", "nk", "nb", "i0", "i1", "j0", "j1"] and fv.dummies[-1] == "stream"
    assert fv.decls["q"].base == "type(c_ptr)" and fv.decls["q_sj"].kind == "c_int64_t"


def test_fortran_example_program(bridge):
    """The Fortran acceptance program: block structure, line rules, every library call imported from the module and
    made with as many actual arguments as the interface has dummies."""
    import fortran_check as fc

    mod = fc.parse_module(bridge.emit_fortran())
    calls = fc.check_program(bridge.emit_fortran_example(), mod)
    for must in ("b2s_init", "b2s_halo_init", "b2s_halo_alloc", "b2s_halo_plan", "b2s_halo_exchange", "b2s_halo_status",
                 "b2s_halo_finalize", "b2s_finalize"):  # fmt: skip
        assert must in calls, must


@pytest.mark.parametrize("mutation,message", [
    (lambda t: t.replace("integer(kind=c_int), value :: nk\n", "", 1), "without a declaration"),
    (lambda t: t.replace("type(c_ptr), value :: stream", "type(c_ptr) :: stream", 1), "by value"),
    (lambda t: t.replace("integer(kind=c_int), value :: device", "integer(kind=c_int64_t), value :: device", 1), "by value"),
    (lambda t: t.replace("end function b2s_init", "end function b2s_finalize", 1), "closes"),
    (lambda t: t.replace("   implicit none\n   private", "   implicit none\n   private\n" + "   ! " + "x" * 140, 1), "132"),
    (lambda t: t.replace("import c_int, ", "import ", 1), "not imported"),
    (lambda t: t.replace("function b2s_halo_rank(ctx)", "function b2s_halo_rank(ctx, extra)", 1), "without a declaration"),
])  # fmt: skip
def test_fortran_checker_catches_defects(bridge, mutation, message):
    """The checker is not a rubber stamp: seeded defects a compiler would reject are rejected here too."""
    import fortran_check as fc

    good = bridge.emit_fortran()
    bad = mutation(good)
    assert bad != good, "the mutation did not apply"
    header = open(os.path.join(ROOT, "include", "b200stencil.h")).read()
    with pytest.raises(fc.FortranError) as e:
        fc.check_against_header(fc.parse_module(bad), header)
    assert message in str(e.value), str(e.value)


def test_cmake_partial_builds_the_c_acceptance_program(bridge, tmp_path):
    """The build hint of the bridge (reference: cli.py:66-77 + templates/cmake.jinja2, "meant as a hint") is current and,
    unlike the reference's, is exercised: a CMake project that includes it builds tests/c_abi/abi_driver.c against the
    imported target, and the program fails loudly at b2s_init where there is no GPU."""
    import shutil

    with open(generate.DEFAULT_CMAKE) as f:
        assert f.read() == bridge.emit_cmake(), "include/CMakeLists_partial.txt is stale: run python -m b200stencil.bridge.generate"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("cmake") is None or shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("cmake, gcc or the CUDA runtime headers are not available")
    (tmp_path / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.18)\nproject(b2s_accept C)\n"
        f"include({generate.DEFAULT_CMAKE})\n"
        f"add_executable(abi_driver {os.path.join(ROOT, 'tests', 'c_abi', 'abi_driver.c')})\n"
        f"target_include_directories(abi_driver PRIVATE {cuda}/include)\n"
        f"target_link_libraries(abi_driver b2s_interface {cuda}/lib64/libcudart.so)\n"
    )
    build = tmp_path / "build"
    for cmd in (["cmake", "-S", str(tmp_path), "-B", str(build)], ["cmake", "--build", str(build)]):
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    exe = build / "abi_driver"
    assert exe.exists()
    try:
        import torch

        if torch.cuda.is_available():
            return  # the program's run on the device is tests/test_c_abi_driver.py
    except ImportError:
        pass
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert run.returncode != 0 and "no CPU fallback" in run.stdout + run.stderr


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_abi, "_state", {})
    with pytest.raises(_abi.LibraryMissing) as e:
        _abi.load()
    assert "no CPU fallback" in str(e.value)
