"""Device-initiated halo exchange over NVLink peer memory.

The field whose halos are exchanged lives in torch *symmetric memory*: every rank allocates the same
layout and maps every peer's buffer into its own address space.  A halo update is then

    device-side barrier (signal pads over NVLink)  ->  ONE ``halo_pull`` kernel

that reads each neighbour's edge strip straight out of the owner's field -- over NVLink for remote
sub-domains, from local HBM for sub-domains on the same GPU -- and writes it into the halo cells.
No pack buffer, no NCCL launch, no unpack.  The barrier orders the pull after every peer's previous
writes of the field and, in a ping-pong time loop, every peer's pull before the next overwrite.

This is the B200-native exchange; ``updater.HaloUpdater`` (packed strips + one grouped NCCL
send/recv) is the portable baseline it is measured against.  torch.distributed supplies the
rendezvous and the barrier kernel; the data path is libb200stencil's kernel.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .. import stencils
from .partitioner import CubedSpherePartitioner
from .updater import FieldGeometry, HaloPlan

PULL_WORDS = 11


def _padded(ni: int, dtype: torch.dtype) -> int:
    per = 16 // torch.empty((), dtype=dtype).element_size()
    return (ni + per - 1) // per * per


class SymmetricField:
    """A halo-padded batch field [b, i, j, k] (i-fastest) allocated in symmetric memory."""

    def __init__(self, shape_ijk: Sequence[int], batch: int, dtype, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        ni, nj, nk = (int(s) for s in shape_ijk)
        nip = _padded(ni, dtype)
        numel = batch * nk * nj * nip
        self.flat = symm_mem.empty(numel, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.flat, group=group or dist.group.WORLD)
        self.field = self.flat.view(batch, nk, nj, nip).permute(0, 3, 2, 1)[:, :ni]
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]


class P2PHaloUpdater:
    """barrier + one halo_pull kernel per update, for a :class:`SymmetricField`."""

    def __init__(self, part: CubedSpherePartitioner, n_gpus: int, gpu: int, sfield: SymmetricField,
                 group_ranks: Optional[Sequence[int]] = None):
        self.part, self.n_gpus, self.gpu = part, n_gpus, gpu
        self.plan = HaloPlan(part, n_gpus, gpu)
        self.sfield = sfield
        ranks = list(group_ranks) if group_ranks is not None else list(range(n_gpus))
        field = sfield.field
        geo = FieldGeometry(field, part.halo)
        es = field.element_size()
        rows = []
        self.remote_bytes = 0
        for l in part.all_links():  # every link whose destination is one of my sub-domains
            if part.gpu_of(l.dst, n_gpus) != gpu:
                continue
            owner = part.gpu_of(l.src, n_gpus)
            b_src, b_dst = part.local_index(l.src, n_gpus), part.local_index(l.dst, n_gpus)
            rows.append([
                geo.cell(b_src, l.si0, l.sj0), geo.step(l.sdi, l.sdj), geo.step(l.spi, l.spj), geo.sk,
                geo.cell(b_dst, l.di0, l.dj0), geo.step(l.ddi, l.ddj), geo.step(l.dpi, l.dpj), geo.sk,
                l.nd, l.np_, sfield.peer_ptrs[ranks[owner]],
            ])  # fmt: skip
            if owner != gpu:
                self.remote_bytes += l.nd * l.np_ * geo.nk * es
        table = np.asarray(rows, dtype=np.int64).reshape(-1, PULL_WORDS)
        self.links = torch.from_numpy(table).to(field.device)
        self.max_strip = int((table[:, 8] * table[:, 9]).max()) if len(table) else 0
        self.nk = geo.nk
        self._call = stencils.prepare_halo_pull(self.links, self.nk, sfield.flat, self.max_strip) if len(table) else None
        self.bytes_sent_per_update = self.remote_bytes  # pulled, not sent: same volume crosses NVLink

    def update(self) -> None:
        if self.n_gpus > 1:
            self.sfield.handle.barrier(channel=0)
        if self._call is not None:
            self._call()
