#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libb200stencil, the Blackwell-specific instructions that prove how data
moves (UTMALDG = TMA tile load, SYNCS = mbarrier, LDGSTS = cp.async, UBLKCP = bulk copy) next to the arithmetic mix
(DFMA/DADD/DMUL, FFMA/FADD/FMUL, MUFU), plain global loads/stores, and registers / spills from the ptxas logs of the
same build.  Runs without a GPU (cuobjdump reads the objects nvcc cross-compiled).

usage: python scripts/sass_summary.py > profiles/rNN_sass_evidence.md
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import ptxas_summary  # noqa: E402

GROUPS = [
    ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"),
    ("LDG", r"\bLDG\b|\bLDG\."), ("STG", r"\bSTG\b|\bSTG\."), ("LDS", r"\bLDS\b|\bLDS\."), ("STS", r"\bSTS\b|\bSTS\."),
    ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\."), ("F64", r"\bD(FMA|ADD|MUL|SETP|MNMX)"), ("F32", r"\bF(FMA|ADD|MUL|SETP|MNMX|SEL)"),
    ("MUFU", r"\bMUFU"), ("ATOM/RED", r"\b(ATOMG|ATOM|RED)\b|\b(ATOMG|RED)\."), ("MEMBAR/FENCE", r"\b(MEMBAR|FENCE)"),
    ("HMMA/UTCMMA", r"\b(HMMA|UTCHMMA|UTCQMMA|UTCIMMA)"),
]  # fmt: skip


def sass_of(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
        kernels[name]["total"] += 1
        for label, pat in GROUPS:
            if re.match(pat, ins):
                kernels[name][label] += 1
    return kernels


def main():
    regs = {}
    for f, n, r, ss, sl, extra in ptxas_summary.rows(""):
        regs[n] = (r, ss, sl)
    print("# SASS evidence (cuobjdump -sass of the objects `__graft_entry__.build()` produces, sm_100a)\n")
    print("Instruction counts are static (per kernel body, not per executed instruction).  `UTMALDG` is the SASS of")
    print("`cp.async.bulk.tensor` (TMA tile load), `SYNCS` the mbarrier arrive/try_wait family, `LDGSTS` is `cp.async`.")
    print("No tensor-core instruction appears anywhere: nothing on this path is a contraction.\n")
    labels = [g[0] for g in GROUPS]
    print("| object | kernel | regs | spill B st/ld | SASS instrs | " + " | ".join(labels) + " |")
    print("|---|---|---|---|---|" + "---|" * len(labels))
    for obj in sorted(glob.glob(os.path.join(ROOT, "geosongpu-ci_b200", "csrc", "_obj", "*.o"))):
        ks = sass_of(obj)
        if not ks:
            continue
        names = ptxas_summary.demangle(list(ks))
        for (mangled, c), n in zip(ks.items(), names):
            r = regs.get(n, ("?", "?", "?"))
            print(f"| {os.path.basename(obj)} | `{n}` | {r[0]} | {r[1]}/{r[2]} | {c['total']} | " + " | ".join(str(c[l]) for l in labels) + " |")


if __name__ == "__main__":
    main()
