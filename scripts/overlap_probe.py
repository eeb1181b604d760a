#!/usr/bin/env python
"""Cost of the overlapped (gated) transport step on ONE GPU, by parts.  For a cube C<n> x <nk> hosted entirely on this
GPU (world = 1: every link is local) and each TMA kernel variant, CUDA-graph replays of
  serial      exchange kernel, then fv_tp2d                      (b2s_halo_exchange + b2s_fv_tp2d)
  gated_serial exchange kernel (gates raised), then fv_tp2d_gated -> what the gate checks alone cost
  overlapped  exchange forked beside fv_tp2d_gated
  fused       b2s_halo_fv_tp2d: exchange + stencil in ONE launch
  stencil     fv_tp2d alone
  exchange    the exchange kernel alone
Usage: python scripts/overlap_probe.py [--n 192] [--nk 72] [--nb 6|3] [--dtype f64] [--option name=value ...]"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "geosongpu-ci_b200")):
    sys.path.insert(0, p)

import torch

from b200stencil import _abi, fields, stencils
from b200stencil.halo.device import HaloContext
from b200stencil.halo.partitioner import CubedSpherePartitioner


def timed_graph(fn, reps=30, per_graph=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(per_graph):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b) / per_graph)
    return statistics.median(ms) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=192)
    ap.add_argument("--nk", type=int, default=72)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--variants", default="2,3")
    ap.add_argument("--option", action="append", default=[])
    ns = ap.parse_args()
    dtype = torch.float64 if ns.dtype == "f64" else torch.float32
    for opt in ns.option:
        k, v = opt.split("=")
        _abi.set_option(k, int(v))
    n, nk = ns.n, ns.nk
    part = CubedSpherePartitioner(n)
    ctx = HaloContext(0, 1, 0)
    q = ctx.field((n + 6, n + 6, nk), 6, dtype)
    g = torch.Generator(device="cuda").manual_seed(1)
    q.uniform_(0.5, 1.5, generator=g)
    ex = ctx.plan(q, part)
    mk = lambda s, lo, hi: fields.empty(s, dtype, batch=6).uniform_(lo, hi, generator=g)  # noqa: E731
    crx, cry = mk((n + 1, n, nk), -0.9, 0.9), mk((n, n + 1, nk), -0.9, 0.9)
    xfx, yfx, rarea = mk((n + 1, n, nk), -1, 1), mk((n, n + 1, nk), -1, 1), mk((n, n), 0.9, 1.1)
    out = fields.empty((n, n, nk), dtype, batch=6)
    es = 8 if dtype == torch.float64 else 4
    gbytes = 6 * n * n * nk * (6 * es + es / nk) / 1e9
    for variant in [int(v) for v in ns.variants.split(",")]:
        _abi.set_option("fv_variant", variant)
        full = stencils.prepare_fv_tp2d(q, crx, xfx, cry, yfx, rarea, out)
        gated = stencils.prepare_fv_tp2d_gated(q, crx, xfx, cry, yfx, rarea, out, ctx.gate)

        def serial():
            ex.update()
            full()

        def gated_serial():
            ex.start(gated=True)
            ex.wait()
            gated()

        def overlapped():
            ex.start(gated=True)
            gated()
            ex.wait()

        fused = stencils.prepare_halo_fv_tp2d(ex, q, crx, xfx, cry, yfx, rarea, out)

        row = {"cube": n, "nk": nk, "dtype": ns.dtype, "variant": variant, "options": ns.option,
               "stencil_us": round(timed_graph(full), 2), "exchange_us": round(timed_graph(ex.update), 2),
               "serial_us": round(timed_graph(serial), 2), "gated_serial_us": round(timed_graph(gated_serial), 2),
               "overlapped_us": round(timed_graph(overlapped), 2), "fused_us": round(timed_graph(fused), 2)}  # fmt: skip
        row["stencil_GBps"] = round(gbytes / (row["stencil_us"] * 1e-6), 1)
        row["exchange_gated_us"] = round(timed_graph(lambda: (ex.start(gated=True), ex.wait())), 2)
        torch.cuda.synchronize()
        gated()  # lowers the gates the line above left open
        for _ in range(3):
            overlapped()
        row["trace_overlapped_ns"] = ctx.trace()
        for _ in range(3):
            gated_serial()
        row["trace_gated_serial_ns"] = ctx.trace()
        for _ in range(3):
            fused()
        row["trace_fused_ns"] = ctx.trace()
        ctx.check()
        print(json.dumps(row), flush=True)
    ctx.finalize()


if __name__ == "__main__":
    main()
