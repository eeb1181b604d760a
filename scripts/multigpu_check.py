#!/usr/bin/env python
"""Multi-GPU correctness check, run under torchrun (one rank per GPU); prints one line per check and a final
JSON line, exits non-zero on the first failure.  Every wait on the device is bounded, so a protocol error shows up
here as an exception, not as a hung GPU.

  1. library-owned exchange (b2s_halo_init / alloc / plan / exchange over NVLink peer memory, csrc/halo_ctx.cu):
     global-id field, every halo cell must hold its geometric neighbour's id -- three epochs, edges and corners;
  2. the overlapped transport step (exchange forked + fv_tp2d_gated) == exchange-then-stencil, eagerly and from a
     replayed CUDA graph, with both TMA kernels;
  3. NCCL baseline: packed-strip exchange adjacency, and NCCL step == device-exchange step bit for bit.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "geosongpu-ci_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

from b200stencil import _abi, fields
from b200stencil.halo.device import HaloContext
from b200stencil.halo.partitioner import CubedSpherePartitioner, global_id_field, layout_for
from b200stencil.halo.transport import FvTransport
from b200stencil.halo.updater import HaloUpdater
from halo_util import batch_field, check_field


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    results = {}

    def say(msg):
        if rank == 0:
            print(f"multigpu_check ok on {world} GPUs: {msg}", flush=True)

    ctx = HaloContext(rank, world, local)  # session name broadcast through torch.distributed (bootstrap only)
    for corners in (False, True):
        N, nk = 96, 5
        part = CubedSpherePartitioner(N, layout_for(world), corners=corners)
        nsub, ni, nj = part.subdomains_per_gpu(world), part.nx, part.ny
        f = ctx.field((ni + 6, nj + 6, nk), nsub, torch.float64, part=part if corners else None)  # staged pushes with corners, in place without
        ex = ctx.plan(f, part, push=True)  # ungated exchanges: same-GPU strips pulled, the rest pushed by their owner
        for rep in range(3):
            for b in range(nsub):
                f[b].copy_(torch.from_numpy(global_id_field(part, rank * nsub + b, nk)))
            torch.cuda.synchronize()
            ctx.barrier()
            if rep == 1:
                ex.start()
                ex.wait()
            else:
                ex.update()
            torch.cuda.synchronize()
            ctx.check()
            check_field(part, world, rank, f, nk)
            ctx.barrier()
        say(f"device exchange adjacency, corners={corners}, 3 epochs, {ex.remote_bytes} B over NVLink per update")
    results["device_exchange_adjacency"] = "ok"

    # ---- transport steps ----
    N, nk = 384, 4
    part = CubedSpherePartitioner(N, layout_for(world))
    nsub, ni, nj = part.subdomains_per_gpu(world), part.nx, part.ny
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    mk = lambda s, lo, hi: fields.empty(s, torch.float64, dev, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    q = ctx.field((ni + 6, nj + 6, nk), nsub, torch.float64, part=part)
    q.uniform_(0.5, 1.5, generator=g)
    ex = ctx.plan(q, part)
    crx, cry = mk((ni + 1, nj, nk), -0.9, 0.9), mk((ni, nj + 1, nk), -0.9, 0.9)
    xfx, yfx, rarea = mk((ni + 1, nj, nk), -1, 1), mk((ni, nj + 1, nk), -1, 1), mk((ni, nj), 0.9, 1.1)
    q_nccl = fields.empty((ni + 6, nj + 6, nk), torch.float64, dev, batch=nsub)
    q_nccl.copy_(q)
    o_serial = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
    FvTransport(part, world, rank, exchange="device", halo_exchange=ex, overlap=False).step(q, crx, xfx, cry, yfx, rarea, o_serial)
    torch.cuda.synchronize()
    ctx.check()
    for variant in (2, 3):
        _abi.set_option("fv_variant", variant)
        tr = FvTransport(part, world, rank, exchange="device", halo_exchange=ex, overlap=True)
        for rep in range(3):
            o = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
            tr.step(q, crx, xfx, cry, yfx, rarea, o)
            torch.cuda.synchronize()
            assert torch.equal(o, o_serial), f"gated step (variant {variant}, rep {rep}) differs: {(o - o_serial).abs().max().item()}"
        o = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            tr.step(q, crx, xfx, cry, yfx, rarea, o)
        for rep in range(5):
            o.zero_()
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(o, o_serial), f"graph replay of the gated step (variant {variant}) differs"
        ctx.check()
        del graph
    _abi.set_option("fv_variant", 0)
    say("overlapped step (exchange || fv_tp2d_gated) == exchange-then-stencil, tile + streaming kernels, eager + graph replay")
    results["gated_step"] = "ok"

    # ---- NCCL baseline ----
    f = batch_field(CubedSpherePartitioner(96, layout_for(world)), world, rank, 5, device=dev, pad=2)
    HaloUpdater(CubedSpherePartitioner(96, layout_for(world)), world, rank).update(f)
    torch.cuda.synchronize()
    check_field(CubedSpherePartitioner(96, layout_for(world)), world, rank, f, 5)
    o_nccl = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
    o_nccl2 = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
    FvTransport(part, world, rank, overlap=True).step(q_nccl, crx, xfx, cry, yfx, rarea, o_nccl)
    FvTransport(part, world, rank, overlap=False).step(q_nccl, crx, xfx, cry, yfx, rarea, o_nccl2)
    torch.cuda.synchronize()
    assert torch.equal(o_nccl, o_nccl2), "NCCL overlapped vs plain step differ"
    assert torch.equal(q_nccl[:, :, 3:-3], q[:, :, 3:-3]) and torch.equal(q_nccl[:, 3:-3], q[:, 3:-3]), "NCCL halos differ from the device exchange"
    assert torch.equal(o_nccl, o_serial), "NCCL step differs from the device-exchange step"
    say("NCCL baseline adjacency; NCCL step == device-exchange step bit for bit")
    results["nccl_equals_device"] = "ok"

    dist.barrier()
    ctx.finalize()
    if rank == 0:
        print(json.dumps({"multigpu_check": "ok", "n_gpus": world, **results}), flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
