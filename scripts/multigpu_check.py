#!/usr/bin/env python
"""Multi-GPU correctness check, run under torchrun (one rank per GPU):
  1. global-id halo exchange over NCCL: every halo cell holds its geometric neighbour's id;
  2. the overlapped transport step (interior | exchange, then frame) equals exchange-then-full-stencil bit for bit.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "geosongpu-ci_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

from b200stencil import fields, stencils
from b200stencil.halo.partitioner import CubedSpherePartitioner, layout_for
from b200stencil.halo.transport import FvTransport
from b200stencil.halo.updater import HaloUpdater
from halo_util import batch_field, check_field


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    N, nk = 96, 5
    part = CubedSpherePartitioner(N, layout_for(world))
    f = batch_field(part, world, rank, nk, device=dev, pad=2)
    up = HaloUpdater(part, world, rank)
    up.update(f)
    torch.cuda.synchronize()
    check_field(part, world, rank, f, nk)

    nsub, ni, nj = part.subdomains_per_gpu(world), part.nx, part.ny
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    mk = lambda s, lo, hi: fields.empty(s, torch.float64, dev, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    q = mk((ni + 6, nj + 6, nk), 0.5, 1.5)
    crx, cry = mk((ni + 1, nj, nk), -0.9, 0.9), mk((ni, nj + 1, nk), -0.9, 0.9)
    xfx, yfx, rarea = mk((ni + 1, nj, nk), -1, 1), mk((ni, nj + 1, nk), -1, 1), mk((ni, nj), 0.9, 1.1)
    q2 = q.clone()
    o1 = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
    o2 = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
    FvTransport(part, world, rank, overlap=True).step(q, crx, xfx, cry, yfx, rarea, o1)
    FvTransport(part, world, rank, overlap=False).step(q2, crx, xfx, cry, yfx, rarea, o2)
    torch.cuda.synchronize()
    assert torch.equal(q, q2), "halos differ between the overlapped and the plain exchange"
    assert torch.equal(o1, o2), f"overlap vs plain: max diff {(o1 - o2).abs().max().item()}"
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok on {world} GPUs: NCCL halo adjacency + overlapped step == plain step", flush=True)

    # ---- peer-memory path: one pull kernel over NVLink ----
    from b200stencil.halo.p2p import P2PHaloUpdater, SymmetricField
    from b200stencil.halo.partitioner import global_id_field

    sf = SymmetricField((ni + 6, nj + 6, nk), nsub, torch.float64, dev)
    for b in range(nsub):
        sf.field[b].copy_(torch.from_numpy(global_id_field(part, rank * nsub + b, nk)))
    torch.cuda.synchronize()
    dist.barrier()
    P2PHaloUpdater(part, world, rank, sf).update()
    torch.cuda.synchronize()
    check_field(part, world, rank, sf.field, nk)
    # p2p transport step == nccl transport step
    sf.field.copy_(q2)
    sf.field[:, :3, 3:-3] = -1.0  # scrub the west halo strips (not the corners) so the pull has to refill them
    torch.cuda.synchronize()
    dist.barrier()
    o3 = fields.zeros((ni, nj, nk), device=dev, batch=nsub)
    FvTransport(part, world, rank, exchange="p2p", symmetric_q=sf).step(sf.field, crx, xfx, cry, yfx, rarea, o3)
    torch.cuda.synchronize()
    assert torch.equal(sf.field, q2), "p2p halos differ from the NCCL exchange"
    assert torch.equal(o3, o2), "p2p step differs from the NCCL step"
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok on {world} GPUs: peer-memory halo_pull adjacency + p2p step == NCCL step", flush=True)

    # ---- EXPERIMENTAL one-launch handshake + pull (halo_pull_sync): only with B2S_CHECK_FUSED=1.  Three updates in a
    #      row exercise the device-resident epoch; every wait in the kernel is bounded (status word), so a protocol
    #      error shows up as an exception here, not as a hung GPU.
    if os.environ.get("B2S_CHECK_FUSED") == "1":
        fused = P2PHaloUpdater(part, world, rank, sf, fused_signal=True)
        for rep in range(3):
            for b in range(nsub):
                sf.field[b].copy_(torch.from_numpy(global_id_field(part, rank * nsub + b, nk)))
            torch.cuda.synchronize()
            dist.barrier()
            fused.update()
            torch.cuda.synchronize()
            fused.check()
            check_field(part, world, rank, sf.field, nk)
            dist.barrier()
        assert int(fused.sync_state[0].item()) == 3 and int(fused.sync_state[1].item()) == 0
        if rank == 0:
            print(f"multigpu_check ok on {world} GPUs: fused handshake + pull (halo_pull_sync), 3 epochs", flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
