"""Per-stencil throughput / roofline sweep:  python -m b200stencil.bench.sweep [--config NAME] [--stencils a,b]"""
from __future__ import annotations

import argparse
import json
import sys

import torch

from . import harness, workloads


def run_one(stencil: str, config: str, dtype, iters: int, warmup: int, graph: bool = False, sub=None) -> dict:
    if sub:  # explicit sub-domain batch: ni,nj,nb,nk
        ni, nj, tiles, nk = sub
        wl = workloads.make(stencil, tiles, ni, nk, dtype, ni=ni, nj=nj)
        config = f"{tiles}x{ni}x{nj}x{nk}"
    else:
        tiles, n, nk = workloads.CONFIGS[config]
        wl = workloads.make(stencil, tiles, n, nk, dtype)
    if graph:
        t = harness.time_graph(wl.run, rotate=wl.slots, launches_per_graph=max(wl.slots * 4, 16), iters=iters, warmup=warmup)
    else:
        t = harness.time_kernel(wl.run, iters=iters, warmup=warmup, rotate=wl.slots)
    peaks = harness.measured_peaks()
    rf = harness.roofline(wl.bytes_per_launch, t["median_ms"], peaks["hbm_gbs"])
    out = {
        "stencil": stencil, "config": config, "timing": "cuda-graph replay" if graph else "events per launch", "dtype": "f64" if dtype == torch.float64 else "f32",
        "points": wl.points, "bytes_per_point": round(wl.bytes_per_point, 3), "slots": wl.slots,
        "median_ms": round(t["median_ms"], 4), "min_ms": round(t["min_ms"], 4),
        "gpts_per_s": round(wl.points / (t["median_ms"] * 1e-3) / 1e9, 3),
        "GBps": round(rf["achieved"], 1), "frac_measured_peak": round(rf["frac"], 4),
        "frac_nominal_8TBs": round(rf["frac_of_nominal_8TBs"], 4), "peak_source": peaks["source"],
    }  # fmt: skip
    if wl.notes:
        out["notes"] = wl.notes
    del wl
    torch.cuda.empty_cache()
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None, help="one of %s (default: each stencil's BASELINE config)" % list(workloads.CONFIGS))
    ap.add_argument("--stencils", default=",".join(workloads.ALL_STENCILS))
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--sub", default=None, help="explicit sub-domain batch ni,nj,nb,nk instead of a named config")
    ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays (removes host launch overhead)")
    ap.add_argument("--option", action="append", default=[], help="libb200stencil option name=value")
    ns = ap.parse_args(argv)
    from .. import _abi

    for opt in ns.option:
        name, value = opt.split("=")
        _abi.set_option(name, int(value))
    rows = []
    for stencil in ns.stencils.split(","):
        for d in ns.dtypes.split(","):
            dtype = torch.float64 if d == "f64" else torch.float32
            cfg = ns.config or workloads.DEFAULT_CONFIG[stencil]
            sub = tuple(int(x) for x in ns.sub.split(",")) if ns.sub else None
            row = run_one(stencil, cfg, dtype, ns.iters, ns.warmup, ns.graph, sub)
            if ns.option:
                row["options"] = ns.option
            rows.append(row)
            print(json.dumps(row), flush=True)
    if ns.out:
        with open(ns.out, "w") as f:
            json.dump(rows, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
