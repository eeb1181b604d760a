#!/usr/bin/env python
"""A/B of the launch modes and exchange-kernel options of the C384x72 transport step inside ONE job (under torchrun, one
rank per GPU): the fields, the halo context and the process group are set up once, then every configuration is warmed up,
captured into a CUDA graph and timed like bench.py times it (K replays between two CUDA events, barrier + synchronize on
both sides, max over ranks, median of the regions).  bench.py pays ~20 s of start-up per configuration; on an 8-GPU box
that is what the GPU budget goes to.

usage: torchrun --nproc-per-node N scripts/step_modes_probe.py [--steps 1000] [--regions 3] [--configs name,name,...]
A configuration is  mode[:option=value[:option=value...]]  with mode = serial | overlap | fused, or one of the two halves
of the serial step alone: stencil (no exchange: the slowest rank's kernel, free of the lock-step) | exchange.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "geosongpu-ci_b200")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

from b200stencil import _abi, fields
from b200stencil.halo.device import HaloContext
from b200stencil.halo.partitioner import CubedSpherePartitioner, layout_for
from b200stencil.halo.transport import FvTransport

DEFAULT = ("serial,serial:halo_variant=1,serial:halo_levels_per_unit=1,serial:halo_blocks_per_sm=2,overlap,"
           "overlap:halo_blocks_per_sm=4,overlap:halo_variant=2,fused")
OPTIONS = ("halo_variant", "halo_levels_per_unit", "halo_blocks_per_sm", "halo_push", "halo_handshake", "halo_levels", "fv_variant")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--regions", type=int, default=3)
    ap.add_argument("--configs", default=DEFAULT)
    ap.add_argument("--cube", type=int, default=384)
    ap.add_argument("--nk", type=int, default=72)
    ns = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    part = CubedSpherePartitioner(ns.cube, layout_for(world), 3)
    nsub, ni, nj, nk = part.subdomains_per_gpu(world), part.nx, part.ny, ns.nk
    g = torch.Generator(device=dev).manual_seed(20240724 + 1000 * rank)
    mk = lambda s, lo, hi: fields.empty(s, torch.float64, dev, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    ctx = HaloContext(rank, world, local)
    q = ctx.field((ni + 6, nj + 6, nk), nsub, torch.float64)
    q.uniform_(0.5, 1.5, generator=g)
    ex = ctx.plan(q, part, push=False)
    crx, cry = mk((ni + 1, nj, nk), -0.9, 0.9), mk((ni, nj + 1, nk), -0.9, 0.9)
    xfx, yfx, rarea = mk((ni + 1, nj, nk), 0.9, 1.1).mul_(crx), mk((ni, nj + 1, nk), 0.9, 1.1).mul_(cry), mk((ni, nj), 0.9, 1.1)
    out = fields.empty((ni, nj, nk), torch.float64, dev, batch=nsub)
    args = (q, crx, xfx, cry, yfx, rarea, out)
    want = None
    for cfg in ns.configs.split(","):
        mode, *opts = cfg.split(":")
        for name in OPTIONS:
            _abi.set_option(name, 1 if name == "halo_push" else 0)
        for o in opts:
            k, v = o.split("=")
            _abi.set_option(k, int(v))
        tr = FvTransport(part, world, rank, exchange="device", halo_exchange=ex, overlap=mode not in ("serial", "stencil", "exchange"),
                         fused=mode == "fused")  # fmt: skip
        step = {"stencil": tr.calls(*args)[0], "exchange": ex.update}.get(mode, lambda: tr.step(*args))
        for _ in range(5):
            step()
        barrier()
        ctx.check()
        same = True
        if want is None and mode != "exchange":
            want = out.clone()
        elif mode != "exchange":
            same = bool(torch.equal(out, want))
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            step()
        for _ in range(3):
            graph.replay()
        barrier()
        region_ms = []
        for _ in range(ns.regions):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(ns.steps):
                graph.replay()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            region_ms.append(ms)
        ctx.check()
        trace = ctx.trace()
        del graph
        if rank == 0:
            print(json.dumps({"n_gpus": world, "config": cfg, "us_per_step": round(statistics.median(region_ms) / ns.steps * 1e3, 2),
                              "regions_us": [round(x / ns.steps * 1e3, 2) for x in region_ms], "same_bits_as_first": same,
                              "trace_ns": trace}), flush=True)  # fmt: skip
        barrier()
    ctx.finalize()
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
