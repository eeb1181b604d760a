"""Time every remap kernel variant on the same resident workloads (one process, one allocation):
python scripts/remap_variants.py [--configs C720x137,C384x72] [--out gpurun_out/remap_variants.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "geosongpu-ci_b200"))

import torch  # noqa: E402

from b200stencil import _abi  # noqa: E402
from b200stencil.bench import harness, workloads  # noqa: E402

NAMES = {1: "nested", 2: "slab+cp.async", 3: "slab+TMA"}


def parse_variant(x):
    """'3' or '3:16:1' = remap_variant[:remap_nw:remap_cg]"""
    p = [int(y) for y in x.split(":")]
    return tuple(p + [0] * (3 - len(p)))



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C720x137,C384x72")
    ap.add_argument("--stencils", default="remap,remap_delp")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--variants", default="1,2,3:8:1,3:16:1,3:8:2")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    ns = ap.parse_args()
    peaks = harness.measured_peaks()
    rows = []
    for cfg in ns.configs.split(","):
        tiles, n, nk = workloads.CONFIGS[cfg]
        for d in ns.dtypes.split(","):
            dtype = torch.float64 if d == "f64" else torch.float32
            for stencil in ns.stencils.split(","):
                wl = workloads.make(stencil, tiles, n, nk, dtype)
                for v, nw, cg in (parse_variant(x) for x in ns.variants.split(",")):
                    _abi.set_option("remap_variant", v)
                    _abi.set_option("remap_nw", nw)
                    _abi.set_option("remap_cg", cg)
                    t = harness.time_kernel(wl.run, iters=ns.iters, warmup=3, rotate=wl.slots)
                    rf = harness.roofline(wl.bytes_per_launch, t["median_ms"], peaks["hbm_gbs"])
                    row = {"stencil": stencil, "config": cfg, "dtype": d, "variant": NAMES[v] + (f" {nw}x{cg}" if nw else ""), "median_ms": round(t["median_ms"], 4),
                           "min_ms": round(t["min_ms"], 4), "GBps": round(rf["achieved"], 1), "frac_measured_peak": round(rf["frac"], 4)}
                    rows.append(row)
                    print(json.dumps(row), flush=True)
                for o in ("remap_variant", "remap_nw", "remap_cg"):
                    _abi.set_option(o, 0)
                del wl
                torch.cuda.empty_cache()
    if ns.out:
        with open(ns.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
