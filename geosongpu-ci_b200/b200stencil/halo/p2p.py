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
PULL_SYNC_WORDS = 12  # + the rank that owns the source sub-domain (-1: this GPU, nothing to wait for)


def build_pull_table(part: CubedSpherePartitioner, n_gpus: int, gpu: int, geo: FieldGeometry, elem_size: int,
                     peer_ptrs: Sequence[int], ranks: Sequence[int], with_source_rank: bool = False):
    """(int64 table [nlinks, 11 | 12], bytes pulled over NVLink per update) of the links that fill ``gpu``'s halos."""
    rows, remote_bytes = [], 0
    for l in part.all_links():  # every link whose destination is one of my sub-domains
        if part.gpu_of(l.dst, n_gpus) != gpu:
            continue
        owner = part.gpu_of(l.src, n_gpus)
        b_src, b_dst = part.local_index(l.src, n_gpus), part.local_index(l.dst, n_gpus)
        row = [
            geo.cell(b_src, l.si0, l.sj0), geo.step(l.sdi, l.sdj), geo.step(l.spi, l.spj), geo.sk,
            geo.cell(b_dst, l.di0, l.dj0), geo.step(l.ddi, l.ddj), geo.step(l.dpi, l.dpj), geo.sk,
            l.nd, l.np_, int(peer_ptrs[ranks[owner]]),
        ]  # fmt: skip
        if with_source_rank:
            row.append(ranks[owner] if owner != gpu else -1)
        rows.append(row)
        if owner != gpu:
            remote_bytes += l.nd * l.np_ * geo.nk * elem_size
    words = PULL_SYNC_WORDS if with_source_rank else PULL_WORDS
    return np.asarray(rows, dtype=np.int64).reshape(-1, words), remote_bytes


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
    """barrier + one halo_pull kernel per update, for a :class:`SymmetricField`.

    ``fused_signal=True`` (EXPERIMENTAL, opt-in: the kernel has not yet been run on more than one GPU) replaces the
    pair by ONE launch, ``halo_pull_sync``: neighbours announce their field through int32 flags in symmetric memory
    and every block waits only for the peer its strip comes from (csrc/k_halo.cu)."""

    def __init__(self, part: CubedSpherePartitioner, n_gpus: int, gpu: int, sfield: SymmetricField,
                 group_ranks: Optional[Sequence[int]] = None, fused_signal: bool = False, group=None):
        self.part, self.n_gpus, self.gpu = part, n_gpus, gpu
        self.plan = HaloPlan(part, n_gpus, gpu)
        self.sfield = sfield
        ranks = list(group_ranks) if group_ranks is not None else list(range(n_gpus))
        field = sfield.field
        geo = FieldGeometry(field, part.halo)
        self.fused_signal = bool(fused_signal) and n_gpus > 1
        table, self.remote_bytes = build_pull_table(part, n_gpus, gpu, geo, field.element_size(), sfield.peer_ptrs, ranks,
                                                    with_source_rank=self.fused_signal)
        self.links = torch.from_numpy(table).to(field.device)
        self.max_strip = int((table[:, 8] * table[:, 9]).max()) if len(table) else 0
        self.nk = geo.nk
        self._call = None
        if len(table) and self.fused_signal:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem

            self.flags = symm_mem.empty(n_gpus, dtype=torch.int32, device=field.device).zero_()
            self._flags_handle = symm_mem.rendezvous(self.flags, group=group or dist.group.WORLD)
            self.peer_flags = torch.tensor([int(p) for p in self._flags_handle.buffer_ptrs], dtype=torch.int64, device=field.device)
            self.sync_state = torch.zeros(3, dtype=torch.int32, device=field.device)  # epoch, blocks done, status
            self._flags_handle.barrier(channel=0)  # every rank's flags are zero before the first announcement
            self._call = stencils.prepare_halo_pull_sync(self.links, self.nk, sfield.flat, self.max_strip, ranks[gpu], n_gpus,
                                                         self.peer_flags, self.sync_state)
        elif len(table):
            self._call = stencils.prepare_halo_pull(self.links, self.nk, sfield.flat, self.max_strip)
        self.bytes_sent_per_update = self.remote_bytes  # pulled, not sent: same volume crosses NVLink

    def update(self) -> None:
        if self.n_gpus > 1 and not self.fused_signal:
            self.sfield.handle.barrier(channel=0)
        if self._call is not None:
            self._call()

    def check(self) -> None:
        """fused_signal: raise if a block gave up waiting for a neighbour (host-synchronising; call outside timed loops)."""
        if self.fused_signal and int(self.sync_state[2].item()) != 0:
            raise RuntimeError("halo_pull_sync: a neighbour's announcement did not arrive within the device-side timeout")
