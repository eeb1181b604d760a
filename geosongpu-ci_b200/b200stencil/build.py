"""In-tree nvcc build of libb200stencil.so (sm_100a only).

``python -m b200stencil.build`` or ``__graft_entry__.build()``.  The library is written next to
the host package (``b200stencil/lib/libb200stencil.so``) so it travels with the repo snapshot
to the GPU box; there is no JIT cache and no other architecture in the fat binary.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.abspath(os.path.join(PKG, "..", "csrc"))
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libb200stencil.so")
OBJDIR = os.path.join(CSRC, "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]

# translation unit -> extra flags.  The vertical scans are compiled without FMA contraction:
# they are HBM-bound, and without contraction they reproduce the oracle bit for bit.
SOURCES = {
    "runtime.cu": [],
    "abi.cu": [],
    "k_patterns.cu": [],
    "k_moist.cu": [],
    "k_vertical.cu": ["-fmad=false"],
    "k_fv.cu": [],
    "k_fv_direct.cu": [],
    "k_fv_tma.cu": [],
    "k_fv_stream.cu": [],
    "tma_host.cu": [],
    "k_fv_split.cu": [],
    "k_fv_split_stream.cu": [],
    "k_remap_slab.cu": ["-fmad=false"],
    "k_remap_ppm.cu": [],
    "k_halo.cu": [],
    "halo_ctx.cu": [],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libb200stencil cannot be built (there is no CPU fallback)")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    from b200stencil.bridge import generate

    generate.main([])  # header + glue from the YAML, so they can never be stale
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    headers.append(generate.DEFAULT_HEADER)
    jobs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [nvcc(), *ARCH, *COMMON, *extra, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr}")
        with open(os.path.join(OBJDIR, src.replace(".cu", ".ptxas.log")), "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-Xlinker", "--exclude-libs,ALL"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
