"""YAML -> C-ABI generator for libb200stencil.

Plays the role of the reference's ``tcn-fpy`` generator
(/root/reference/src/tcn/py_ftn_interface/cli.py:80-136, bridge.py:30-184) with the call
direction reversed: there Fortran calls *into* Python through generated C glue
(templates/interface.c.jinja2:8-30); here Python (or Fortran, through the same
``bind(c)`` symbols) calls *into* CUDA.  One YAML file, in the reference's schema
(argument.py:6-98), is the single source of truth for

* ``include/b200stencil.h``  - the C prototypes (``emit_header``),
* the cffi ``cdef`` the host package loads the library with (``emit_cdef``),
* the argument marshalling tables (``Bridge.functions``),
* a Fortran ``bind(c)`` interface module (``emit_fortran``, cf. templates/interface.f90.jinja2:1-56).

CLI:  python -m b200stencil.bridge.generate [DEF.yaml] [--header OUT.h] [--fortran OUT.f90]
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_YAML = os.path.join(HERE, "b200stencil.yaml")
REPO_ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
DEFAULT_HEADER = os.path.join(REPO_ROOT, "include", "b200stencil.h")
DEFAULT_GLUE = os.path.abspath(os.path.join(HERE, "..", "..", "csrc", "abi_glue.inc"))
DEFAULT_FORTRAN = os.path.join(REPO_ROOT, "include", "b2s_interface_mod.f90")
DEFAULT_FORTRAN_EXAMPLE = os.path.join(REPO_ROOT, "include", "b2s_example.f90")
DEFAULT_CMAKE = os.path.join(REPO_ROOT, "include", "CMakeLists_partial.txt")

PRECISIONS = {"double": ("double", "int64_t", ""), "float": ("float", "int32_t", "_f32")}

_SCALARS = {"int": "int", "float": "float", "double": "double", "int64": "int64_t"}
_ARRAYS = {"array_int": "int", "array_float": "float", "array_double": "double", "array_int64": "int64_t"}


@dataclass
class Argument:
    """One YAML ``!Argument`` (reference: argument.py:6-86)."""

    name: str
    type: str
    dims: Optional[int] = None
    intent: str = "in"  # in | inout | out, from the YAML section it sits in

    @property
    def name_sanitize(self) -> str:  # reference: argument.py:16-20
        return f"_{self.name}" if self.name in ("is", "in") else self.name

    @property
    def is_array(self) -> bool:
        return self.type.startswith("array_")

    @property
    def is_precision_dependent(self) -> bool:
        return self.type in ("real", "array_real", "array_index")

    def element_ctype(self, precision: str) -> str:
        real, index, _ = PRECISIONS[precision]
        if self.type in ("real", "array_real"):
            return real
        if self.type == "array_index":
            return index
        if self.type in _SCALARS:
            return _SCALARS[self.type]
        if self.type in _ARRAYS:
            return _ARRAYS[self.type]
        if self.type == "MPI":  # kept for schema compatibility (argument.py:58-59)
            return "void*"
        raise RuntimeError(f"ERROR_DEF_TYPE_TO_C: {self.type}")

    def stride_names(self) -> List[str]:
        if not self.is_array:
            return []
        return {3: ["sj", "sk", "sb"], 2: ["sj", "sb"]}.get(self.dims or 1, [])

    def c_parameters(self, precision: str) -> List[str]:
        t = self.element_ctype(precision)
        if not self.is_array:
            return [f"{t} {self.name}"]
        const = "const " if self.intent == "in" else ""
        return [f"{const}{t}* {self.name}"] + [f"int64_t {self.name}_{s}" for s in self.stride_names()]


@dataclass
class Function:
    """One bridge entry (reference: base.py:9-40)."""

    name: str
    inputs: List[Argument] = field(default_factory=list)
    inouts: List[Argument] = field(default_factory=list)
    outputs: List[Argument] = field(default_factory=list)
    cites: str = ""

    @property
    def arguments(self) -> List[Argument]:  # reference order: base.py:35-36
        return self.inputs + self.inouts + self.outputs

    @property
    def is_precision_dependent(self) -> bool:
        return any(a.is_precision_dependent for a in self.arguments)

    def precisions(self) -> List[str]:
        return list(PRECISIONS) if self.is_precision_dependent else ["double"]

    def symbol(self, prefix: str, precision: str) -> str:
        return f"{prefix}_{self.name}{PRECISIONS[precision][2]}_c"

    def c_prototype(self, prefix: str, precision: str) -> str:
        params: List[str] = []
        for a in self.arguments:
            params += a.c_parameters(precision)
        params.append("void* stream")
        return f"int {self.symbol(prefix, precision)}({', '.join(params)});"


def _argument_constructor(loader, node):
    return dict(loader.construct_mapping(node))


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_constructor("!Argument", _argument_constructor)  # reference: argument.py:89-98


class Bridge:
    """Parsed bridge definition (reference: Bridge.make_from_yaml, bridge.py:30-62)."""

    def __init__(self, prefix: str, functions: List[Function]):
        self.prefix = prefix
        self.functions: Dict[str, Function] = {f.name: f for f in functions}

    @classmethod
    def from_yaml(cls, path: str = DEFAULT_YAML) -> "Bridge":
        with open(path) as f:
            defs = yaml.load(f, Loader=_Loader)
        if defs.get("type") != "py_ftn_interface":
            raise RuntimeError(f"{path}: not a py_ftn_interface definition")
        functions = []
        for entry in defs["bridge"]:
            args = entry.get("arguments")
            if args in (None, "None"):
                args = {}
            fn = Function(entry["name"], cites=entry.get("cites", ""))
            for section, intent, dest in (
                ("inputs", "in", fn.inputs),
                ("inouts", "inout", fn.inouts),
                ("outputs", "out", fn.outputs),
            ):
                for a in args.get(section) or []:
                    dest.append(Argument(intent=intent, **a))
            functions.append(fn)
        return cls(defs["name"], functions)

    # ---- emitters -------------------------------------------------------------------------

    def prototypes(self) -> List[str]:
        out = []
        for fn in self.functions.values():
            for precision in fn.precisions():
                out.append(fn.c_prototype(self.prefix, precision))
        return out

    def symbols(self) -> List[str]:
        out = list(RUNTIME_SYMBOLS)
        for fn in self.functions.values():
            out += [fn.symbol(self.prefix, p) for p in fn.precisions()]
        return out

    def emit_cdef(self) -> str:
        return RUNTIME_CDEF + "\n".join(self.prototypes()) + "\n"

    def emit_header(self) -> str:
        lines = [HEADER_PROLOGUE]
        for fn in self.functions.values():
            lines.append(f"/* {fn.name}: {fn.cites} */")
            for precision in fn.precisions():
                lines.append("B2S_API " + fn.c_prototype(self.prefix, precision))
            lines.append("")
        lines.append(HEADER_EPILOGUE)
        return "\n".join(lines)

    def emit_glue(self) -> str:
        """``extern "C"`` definitions forwarding the flat arguments to ``b2s::impl::<fn><T>(...)``.

        The CUDA sources include this file (csrc/abi_glue.inc), so a stencil implementation whose
        signature drifts from the YAML fails to compile (the generated C shim of the reference,
        templates/interface.c.jinja2:8-30, plays the same forwarding role).
        """
        out = ["// GENERATED by b200stencil/bridge/generate.py from b200stencil.yaml -- do not edit.", ""]
        for fn in self.functions.values():
            for precision in fn.precisions():
                real, index, _ = PRECISIONS[precision]
                params, call = [], []
                for a in fn.arguments:
                    params += a.c_parameters(precision)
                    t = a.element_ctype(precision)
                    if a.is_array and (a.dims or 1) >= 2:
                        ct = f"const {t}" if a.intent == "in" else t
                        strides = ", ".join(f"{a.name}_{s}" for s in a.stride_names())
                        call.append(f"b2s::F{a.dims}<{ct}>{{{a.name}, {strides}}}")
                    else:
                        call.append(a.name)
                params.append("void* stream")
                call.append("static_cast<cudaStream_t>(stream)")
                targ = f"<{real}>" if fn.is_precision_dependent else ""
                out += [
                    f'extern "C" int {fn.symbol(self.prefix, precision)}({", ".join(params)}) {{',
                    f"  return b2s::impl::{fn.name}{targ}({', '.join(call)});",
                    "}",
                    "",
                ]
        return "\n".join(out)

    def emit_fortran(self) -> str:
        """``bind(c)`` interface module for Fortran callers (cf. templates/interface.f90.jinja2:1-56): the stencil entry
        points plus the runtime / halo lifecycle, lines wrapped to 100 columns (``bridge/fortran.py``)."""
        from . import fortran

        return fortran.emit_module(self)

    def emit_fortran_example(self) -> str:
        """The Fortran acceptance program (cf. test/py_ftn_interface/data/fortran_program.f90)."""
        from . import fortran

        return fortran.emit_example_program(self)

    def emit_cmake(self) -> str:
        """``CMakeLists_partial.txt``: the build hint the reference generator drops beside its bridge
        (cli.py:66-77 ``Build.generate_cmake``, templates/cmake.jinja2, with its [ADAPT] / [COPY] convention).

        There the partial builds a CFFI library out of Python at configure time and needs MPI and an interpreter; here the
        stencils already ARE a shared library, so the partial declares it as an IMPORTED target with the generated header's
        directory, names the generated Fortran module as a source to append, and needs nothing else."""
        P = self.prefix.upper()
        return f"""# GENERATED by b200stencil/bridge/generate.py from b200stencil.yaml -- do not edit.
# [ADAPT] comments point to lines to adapt in your CMakeLists.txt, [COPY] is code to copy as it stands
# (the convention of the reference generator's hint, src/tcn/py_ftn_interface/templates/cmake.jinja2).

#
## [ADAPT] Where the repository (or an installed copy of include/ and lib/) lives.
#

if(NOT DEFINED {P}_ROOT)
  get_filename_component({P}_ROOT "${{CMAKE_CURRENT_LIST_DIR}}/.." ABSOLUTE)
endif()
set({P}_INCLUDE_DIR "${{{P}_ROOT}}/include" CACHE PATH "directory of b200stencil.h and b2s_interface_mod.f90")
set({P}_LIBRARY "${{{P}_ROOT}}/geosongpu-ci_b200/b200stencil/lib/libb200stencil.so" CACHE FILEPATH "the C-ABI library")

#
## [COPY] {self.prefix} interface: the C-ABI library as an imported target.
#

message(STATUS "Using the {self.prefix} stencil library: ${{{P}_LIBRARY}}")
if(NOT EXISTS "${{{P}_LIBRARY}}")
  message(SEND_ERROR "libb200stencil.so is not built: run python __graft_entry__.py in ${{{P}_ROOT}}")
endif()
add_definitions(-DRUN_{self.prefix})
add_library({self.prefix}_interface SHARED IMPORTED GLOBAL)
set_target_properties({self.prefix}_interface PROPERTIES
  IMPORTED_LOCATION "${{{P}_LIBRARY}}"
  IMPORTED_NO_SONAME TRUE
  INTERFACE_INCLUDE_DIRECTORIES "${{{P}_INCLUDE_DIR}}")

# Fortran callers compile the generated bind(c) module with their own sources
set({self.prefix}_interface_sources "${{{P}_INCLUDE_DIR}}/b2s_interface_mod.f90")
# [ADAPT] : use this to append the interface source to your program
# list( APPEND sources ${{{self.prefix}_interface_sources}} )

#
## [ADAPT] Executable.
#

# [ADAPT] link your target against the interface (a device pointer and a cudaStream_t are all it takes from CUDA;
# callers that allocate through the CUDA runtime also link CUDA::cudart)
#target_link_libraries(test {self.prefix}_interface)
"""


RUNTIME_SYMBOLS = [
    "b2s_init",
    "b2s_finalize",
    "b2s_device",
    "b2s_last_error",
    "b2s_abi_version",
    "b2s_sm_count",
    "b2s_set_option",
    "b2s_get_option",
    "b2s_launch_count",
    "b2s_halo_init",
    "b2s_halo_finalize",
    "b2s_halo_rank",
    "b2s_halo_world",
    "b2s_halo_barrier",
    "b2s_halo_alloc",
    "b2s_halo_free",
    "b2s_halo_peer_ptr",
    "b2s_halo_plan",
    "b2s_halo_plan_remote_bytes",
    "b2s_halo_exchange",
    "b2s_halo_exchange_start",
    "b2s_halo_exchange_wait",
    "b2s_halo_gate",
    "b2s_halo_status",
    "b2s_halo_trace",
]

RUNTIME_CDEF = """
int b2s_init(int device);
int b2s_finalize(void);
int b2s_device(void);
const char* b2s_last_error(void);
int b2s_abi_version(void);
int b2s_sm_count(void);
int b2s_set_option(const char* name, int value);
int b2s_get_option(const char* name);
int64_t b2s_launch_count(void);
int b2s_halo_init(const char* session, int rank, int world, int device, int64_t* ctx);
int b2s_halo_finalize(int64_t ctx);
int b2s_halo_rank(int64_t ctx);
int b2s_halo_world(int64_t ctx);
int b2s_halo_barrier(int64_t ctx);
int b2s_halo_alloc(int64_t ctx, int64_t nbytes, void** ptr);
int b2s_halo_free(int64_t ctx, void* ptr);
int b2s_halo_peer_ptr(int64_t ctx, const void* ptr, int peer, void** peer_ptr);
int b2s_halo_plan(int64_t ctx, const void* field, int elem_size, int nk, int nlinks, const int64_t* links, int* plan);
int64_t b2s_halo_plan_remote_bytes(int64_t ctx, int plan);
int b2s_halo_exchange(int64_t ctx, int plan, void* stream);
int b2s_halo_exchange_start(int64_t ctx, int plan, int gated, void* stream);
int b2s_halo_exchange_wait(int64_t ctx, void* stream);
int b2s_halo_gate(int64_t ctx, int** gate);
int b2s_halo_status(int64_t ctx, int* epoch, int* status);
int b2s_halo_trace(int64_t ctx, int64_t* out6);
"""

HEADER_PROLOGUE = """/* b200stencil.h -- C-ABI of libb200stencil.so (GENERATED, do not edit).
 *
 * Generated by geosongpu-ci_b200/b200stencil/bridge/generate.py from
 * geosongpu-ci_b200/b200stencil/bridge/b200stencil.yaml, which uses the schema of the
 * reference's Fortran<->C<->Python bridge generator
 * (/root/reference/src/tcn/py_ftn_interface/README.md:9-22, argument.py:54-86).
 *
 * What each entry point replaces: the reference reaches a stencil through
 *   {prefix}_{fn}_f  -> bind(c) {prefix}_{fn}_c -> {prefix}_{fn}_py -> hook -> gt4py/NDSL stencil
 * (templates/interface.f90.jinja2:26-39, interface.c.jinja2:8-30, interface.py.jinja2:26-51).
 * Here {prefix}_{fn}_c *is* the stencil: a hand-written sm_100a kernel launch.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library allocates nothing
 *     persistent and performs no hidden host<->device copies;
 *   - fields are i-fastest: a dims-3 argument x is followed by x_sj, x_sk, x_sb (strides in
 *     elements of j, k and the tile/sub-domain batch axis b), a dims-2 argument by x_sj, x_sb;
 *     the pointer addresses compute-domain cell (0,0,0) (halo cells sit at negative offsets);
 *   - calls are asynchronous on `stream` (a cudaStream_t; NULL = the default stream);
 *   - return value: 0 ok, < 0 argument/state error, > 0 a cudaError_t; b2s_last_error() gives
 *     the message (thread-local).  The reference's void bridge has no error channel
 *     (interface.c.jinja2:8, SURVEY.md 8b) -- this one does;
 *   - {fn}_c computes in double (index outputs int64), {fn}_f32_c in float (index outputs int32).
 */
#ifndef B200STENCIL_H
#define B200STENCIL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2S_ABI_VERSION 1
#if defined(__GNUC__)
#define B2S_API __attribute__((visibility("default")))
#else
#define B2S_API
#endif

/* runtime: init / finalize triple of the GEOS bridge (example_def_dycore.yaml:4,21,71) */
B2S_API int b2s_init(int device);
B2S_API int b2s_finalize(void);
/* device bound by b2s_init, -1 before */
B2S_API int b2s_device(void);
B2S_API const char* b2s_last_error(void);
B2S_API int b2s_abi_version(void);
B2S_API int b2s_sm_count(void);
/* tuning knobs (kernel variant selection for benchmarking); unknown names return -1 */
B2S_API int b2s_set_option(const char* name, int value);
B2S_API int b2s_get_option(const char* name);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
B2S_API int64_t b2s_launch_count(void);

/* ---- multi-GPU halo exchange lifecycle (csrc/halo_ctx.cu) -------------------------------------------------------
 * The reference passes the communicator THROUGH the C boundary (argument type MPI: py_ftn_interface/argument.py:54-86,
 * MPI_Comm_f2c in base.py:78-96) and brackets the run with init / finalize (example_def_dycore.yaml:4,21,71); the halo
 * update itself is NDSL's HaloUpdater over mpi4py.  Here the exchange is owned by the library: one rank (process or
 * thread) per GPU of ONE node, peer memory over NVLink, no MPI, no NCCL, no Python.
 *   session   the analogue of the communicator: a name shared by the ranks of one run (POSIX shared-memory rendezvous;
 *             choose it unique per run, e.g. from the launcher's job id).  May be NULL when world == 1.
 *   ctx       opaque handle (int64 so that Fortran can hold it in an integer(c_int64_t)).
 * Collective calls (every rank, same order): b2s_halo_init, b2s_halo_alloc, b2s_halo_free, b2s_halo_barrier,
 * b2s_halo_finalize, and every exchange.  Host-side waits are bounded (B2S_RDV_TIMEOUT seconds, default 60). */
B2S_API int b2s_halo_init(const char* session, int rank, int world, int device, int64_t* ctx);
B2S_API int b2s_halo_finalize(int64_t ctx);
B2S_API int b2s_halo_rank(int64_t ctx);
B2S_API int b2s_halo_world(int64_t ctx);
B2S_API int b2s_halo_barrier(int64_t ctx);
/* symmetric device allocation: the same nbytes on every rank, every peer's buffer mapped into this process */
B2S_API int b2s_halo_alloc(int64_t ctx, int64_t nbytes, void** ptr);
B2S_API int b2s_halo_free(int64_t ctx, void* ptr);
/* address, in this process, of rank `peer`'s copy of the byte `ptr` points to (ptr inside a b2s_halo_alloc buffer) */
B2S_API int b2s_halo_peer_ptr(int64_t ctx, const void* ptr, int peer, void** peer_ptr);
/* bind a link table to a field.  links: HOST int64[nlinks][12]; words 0..9 as for b2s_halo_move (element offsets
 * relative to `field`, the same on every rank), [10] = rank that owns the source sub-domain, [11] = destination
 * sub-domain (batch index of the field, < 64; 0 for an unbatched field) in its low 16 bits.
 * Optional marks in word [11], for the push path of the ungated exchange (both ends of a strip must be marked alike on
 * their ranks; gated and fused exchanges pull every incoming strip in place, marked or not):
 *   B2S_HALO_LINK_OUT     the row is an OUTGOING strip: its source is on this rank, [10] = rank that owns the destination;
 *   B2S_HALO_LINK_PUSHED  an incoming strip its owner pushes (the owner's table has the matching OUT row);
 *   B2S_HALO_LINK_STAGED  a same-rank copy to run AFTER the deliveries of rank [10]: the owner pushed the strip, packed,
 *                         into a staging area of this rank's allocation (the OUT row's destination); this row unpacks it. */
#define B2S_HALO_LINK_OUT ((int64_t)1 << 32)
#define B2S_HALO_LINK_PUSHED ((int64_t)1 << 33)
#define B2S_HALO_LINK_STAGED ((int64_t)1 << 34)
B2S_API int b2s_halo_plan(int64_t ctx, const void* field, int elem_size, int nk, int nlinks, const int64_t* links, int* plan);
B2S_API int64_t b2s_halo_plan_remote_bytes(int64_t ctx, int plan);
/* halo update.  b2s_halo_exchange runs it on `stream`: a one-block handshake kernel (announce this rank's field, await the
 * neighbours' announcements) followed by the strip copies over peer memory (b2s_set_option("halo_variant", 1 | 2): ONE kernel with
 * the handshake inside);
 * _start forks it onto the context's own high-priority stream (ordered after the work already on `stream`), _wait
 * joins: what the caller enqueues on `stream` in between overlaps the exchange.  Both can be captured in a CUDA graph.
 * gated != 0: the kernel opens gate[b] when the halos of sub-domain b are complete (b = 0, 1, ... in turn); exactly one
 * gated stencil launch (b2s_fv_tp2d_gated_c, same batch) must consume the gates between _start and _wait. */
B2S_API int b2s_halo_exchange(int64_t ctx, int plan, void* stream);
B2S_API int b2s_halo_exchange_start(int64_t ctx, int plan, int gated, void* stream);
B2S_API int b2s_halo_exchange_wait(int64_t ctx, void* stream);
/* device address of the gate words (int32[66]: one flag per sub-domain, CTA counter, status) for gated stencil launches */
B2S_API int b2s_halo_gate(int64_t ctx, int** gate);
/* host-synchronising: exchanges completed; status bit 0 = a neighbour's announcement timed out on the device,
 * bit 1 = a gated stencil gave up waiting for the gate */
B2S_API int b2s_halo_status(int64_t ctx, int* epoch, int* status);
/* diagnostics, host-synchronising: device timeline (ns) of the last exchange + gated stencil: exchange start, exchange end,
 * gate 0 opened, first stencil CTA started, stencil CTA 0 passed gate 0, last stencil CTA finished */
B2S_API int b2s_halo_trace(int64_t ctx, int64_t* out6);
"""

HEADER_EPILOGUE = """#ifdef __cplusplus
}
#endif
#endif /* B200STENCIL_H */
"""


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("definition", nargs="?", default=DEFAULT_YAML)
    ap.add_argument("--header", default=DEFAULT_HEADER)
    ap.add_argument("--glue", default=DEFAULT_GLUE)
    ap.add_argument("--fortran", default=DEFAULT_FORTRAN)
    ap.add_argument("--fortran-example", default=DEFAULT_FORTRAN_EXAMPLE)
    ap.add_argument("--cmake", default=DEFAULT_CMAKE, help="build hint (CMakeLists_partial.txt); empty to skip")
    ap.add_argument("--check", action="store_true", help="fail if the header on disk is stale")
    ns = ap.parse_args(argv)
    bridge = Bridge.from_yaml(ns.definition)
    text = bridge.emit_header()
    if ns.check:
        with open(ns.header) as f:
            return 0 if f.read() == text else 1
    os.makedirs(os.path.dirname(ns.header), exist_ok=True)
    with open(ns.header, "w") as f:
        f.write(text)
    with open(ns.glue, "w") as f:
        f.write(bridge.emit_glue())
    if ns.fortran:
        with open(ns.fortran, "w") as f:
            f.write(bridge.emit_fortran())
    if ns.fortran_example:
        with open(ns.fortran_example, "w") as f:
            f.write(bridge.emit_fortran_example())
    if ns.cmake:
        with open(ns.cmake, "w") as f:
            f.write(bridge.emit_cmake())
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
