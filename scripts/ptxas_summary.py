#!/usr/bin/env python
"""Registers / spills / shared memory per kernel from the nvcc -Xptxas -v logs of the last build
(geosongpu-ci_b200/csrc/_obj/*.ptxas.log, untracked).  Usage: python scripts/ptxas_summary.py [substring]"""
import glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", re.sub(r"\(anonymous namespace\)::|b2s::impl::|b2s::", "", o.replace("void ", ""))) for o in out]


def rows(pattern=""):
    res = []
    for log in sorted(glob.glob(os.path.join(ROOT, "geosongpu-ci_b200", "csrc", "_obj", "*.ptxas.log"))):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", txt):
            res.append((os.path.basename(log).replace(".ptxas.log", ""), m.group(1), int(m.group(5)), int(m.group(3)), int(m.group(4)), m.group(6)))
    names = demangle([r[1] for r in res])
    return [(f, n, regs, ss, sl, extra) for (f, _, regs, ss, sl, extra), n in zip(res, names) if pattern in n or pattern in f]


if __name__ == "__main__":
    for f, n, regs, ss, sl, extra in rows(sys.argv[1] if len(sys.argv) > 1 else ""):
        smem = re.search(r"(\d+) bytes smem", extra)
        print(f"{f:20s} {n:60s} regs {regs:3d} spill st/ld {ss}/{sl} smem {smem.group(1) if smem else 0}")
