"""One flux-form transport step on a GPU's batch of sub-domains: halo update of q + fv_tp2d.

This is the unit bench.py times (BASELINE config 4: "FV3-style horizontal finite-volume
flux/advection stencil with 3-point halo on C384x72, halo exchange at 2/4/8 GPUs").  The exchange runs
on its own stream.  Product path (library-owned exchange, halo/device.py): ONE gated stencil launch that
computes sub-domain b as soon as its halos have landed while those of the next sub-domains are still in
flight.  NCCL baseline: an interior launch beside the exchange, then frame launches.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import stencils
from .partitioner import CubedSpherePartitioner
from .updater import HaloUpdater

Rect = Tuple[int, int, int, int]


def split_regions(ni: int, nj: int, halo: int = 3, side: int = 32) -> Tuple[Rect, List[Rect]]:
    """Interior rectangle (reads no halo cell) and the boundary frame as (i0, i1, j0, j1) rectangles.

    The west/east frame strips are ``side`` columns wide (not ``halo``): a 3-column rectangle would
    leave 29 of every 32 lanes idle, a 32-column one runs the 32-column TMA tile at full width.
    """
    w = side if ni >= 4 * side else halo
    w = max(w, halo)
    if ni <= 2 * w or nj <= 2 * halo:
        return (0, 0, 0, 0), [(0, ni, 0, nj)]
    interior = (w, ni - w, halo, nj - halo)
    frame = [(0, ni, 0, halo), (0, ni, nj - halo, nj), (0, w, halo, nj - halo), (ni - w, ni, halo, nj - halo)]
    return interior, frame


class FvTransport:
    """One transport step on this GPU's batch of sub-domains.

    ``exchange`` =
      "device"  (the product path) the library-owned exchange of ``halo/device.py``: a handshake kernel + the strip copies per halo update
                (neighbour handshake + pull over NVLink peer memory).  ``q`` must be the field ``halo_exchange`` was
                planned for (``HaloContext.field`` + ``HaloContext.plan``).  With ``overlap`` the exchange is forked onto
                the context's stream and opens one gate per sub-domain as its halos land; ``fv_tp2d_gated`` walks the
                batch in the same order, so sub-domain b is computed while the halos of b+1.. are in flight -- one
                stencil launch, its DRAM-friendly item order untouched (interior-cells-first was measured: it costs
                13-28 % of the stencil, profiles/r02_overlap.md).  With ``fused`` the whole step is ONE launch
                (``b2s_halo_fv_tp2d``): the CTAs of the stencil grid share the exchange among themselves first;
      "nccl"    (portable baseline, torch.distributed) packed strips + grouped NCCL send/recv, optionally overlapped
                with an interior launch followed by four frame launches.
    """

    def __init__(self, part: CubedSpherePartitioner, n_gpus: int, gpu: int, process_group=None,
                 overlap: bool = True, side: int = 32, exchange: str = "nccl", halo_exchange=None, fused: bool = False):
        self.part, self.n_gpus, self.gpu = part, n_gpus, gpu
        self.exchange = exchange
        self.dev_exchange = None
        self.updater = None
        self.fused = False
        if exchange == "device":
            if halo_exchange is None:
                raise ValueError('exchange="device" needs the HaloExchange (HaloContext.plan) of the field that holds q')
            self.dev_exchange = halo_exchange
            self.fused = bool(fused)
            self.overlap = bool(overlap) or self.fused
        elif exchange == "nccl":
            self.updater = HaloUpdater(part, n_gpus, gpu, process_group=process_group)
            self.overlap = overlap and bool(self.updater.plan.peers)
        else:
            raise ValueError(f"unknown exchange {exchange!r}: 'device' or 'nccl'")
        self.interior, self.frame = split_regions(part.nx, part.ny, part.halo, side)
        self._calls = {}

    def calls(self, q, crx, xfx, cry, yfx, rarea, q_out, q_out_halo: int = 0):
        """(full, interior, [frame...], gated) PreparedCalls for this set of fields, marshalled once."""
        fs = (q, crx, xfx, cry, yfx, rarea, q_out)
        key = tuple((t.data_ptr(), tuple(t.stride()), tuple(t.shape)) for t in fs) + (q_out_halo,)
        if key not in self._calls:
            mk = lambda region: stencils.prepare_fv_tp2d(*fs, region=region, q_out_halo=q_out_halo)  # noqa: E731
            if self.dev_exchange is not None:
                if self.fused:
                    gated = stencils.prepare_halo_fv_tp2d(self.dev_exchange, *fs, q_out_halo=q_out_halo)
                elif self.overlap:
                    gated = stencils.prepare_fv_tp2d_gated(*fs, gate=self.dev_exchange.ctx.gate, q_out_halo=q_out_halo)
                else:
                    gated = None
                self._calls[key] = (mk(None), None, [], gated)
            else:
                interior = mk(self.interior) if self.interior[1] > self.interior[0] else None
                self._calls[key] = (mk(None), interior, [mk(r) for r in self.frame], None)
        return self._calls[key]

    def step(self, q, crx, xfx, cry, yfx, rarea, q_out, q_out_halo: int = 0) -> None:
        """q (halo-padded batch field) -> q_out; q's halos are refreshed from the neighbours first."""
        full, interior, frame, gated = self.calls(q, crx, xfx, cry, yfx, rarea, q_out, q_out_halo)
        if self.dev_exchange is not None:
            ex = self.dev_exchange
            if q.data_ptr() != ex.field.data_ptr():
                raise ValueError("device exchange: q is not the field this transport's HaloExchange was planned for")
            if self.fused:
                gated()
            elif gated is not None:
                ex.start(gated=True)
                gated()
                ex.wait()
            else:
                ex.update()
                full()
            return
        if not self.overlap:
            self.updater.update(q)
            full()
            return
        self.updater.start(q)
        if interior is not None:
            interior()
        self.updater.wait()
        for call in frame:
            call()


class SplitTransport:
    """One FV3-style transport step with the inner/outer operator splitting (S5b, ``fv_tp2d_split``):

        halo update of q INCLUDING the corner blocks  ->  q_out = fv_tp2d_split(q, ...)

    The partitioner must have been built with ``corners=True``: corner blocks then arrive from the diagonal
    neighbour in the same exchange as the edge strips, and at the eight cube corners the exchange writes FV3's
    copy_corners values for x-sweeps while the kernel derives the y-sweep values itself (``corner_flags``).
    ``exchange`` = "nccl" (packed strips, grouped send/recv) or "device" (library-owned peer-memory exchange;
    ``halo_exchange`` = the ``HaloContext.plan`` of the field that holds q).
    """

    def __init__(self, part: CubedSpherePartitioner, n_gpus: int, gpu: int, process_group=None, exchange: str = "nccl",
                 halo_exchange=None):
        if not part.corners:
            raise ValueError("fv_tp2d_split reads the halo corners: build the partitioner with corners=True")
        self.part, self.n_gpus, self.gpu = part, n_gpus, gpu
        self.updater = None
        self.dev_exchange = None
        if exchange == "device":
            if halo_exchange is None:
                raise ValueError('exchange="device" needs the HaloExchange (HaloContext.plan) of the field that holds q')
            self.dev_exchange = halo_exchange
        else:
            self.updater = HaloUpdater(part, n_gpus, gpu, process_group=process_group)
        nsub = part.subdomains_per_gpu(n_gpus)
        self.corner_flags_host = [part.cube_corner_flags(gpu * nsub + b) for b in range(nsub)]
        self._flags = {}

    def corner_flags(self, device) -> torch.Tensor:
        if device not in self._flags:
            self._flags[device] = torch.tensor(self.corner_flags_host, dtype=torch.int32, device=device)
        return self._flags[device]

    def step(self, q, crx, xfx, cry, yfx, area, rarea, q_out, fx_out=None, fy_out=None) -> None:
        if self.dev_exchange is not None:
            if q.data_ptr() != self.dev_exchange.field.data_ptr():
                raise ValueError("device exchange: q is not the field this transport's HaloExchange was planned for")
            self.dev_exchange.update()
        else:
            self.updater.update(q)
        stencils.fv_tp2d_split(q, crx, xfx, cry, yfx, area, rarea, q_out, fx_out, fy_out, corner_flags=self.corner_flags(q.device))


class DycoreChain:
    """BASELINE config 5: horizontal FV transport followed by the vertical remap scan, per step

        halo update of q  ->  q_adv = fv_tp2d(q, ...)  ->  pe1 = pe_prefix(delp, ptop)  ->  q_new = remap(pe1, q_adv, pe2)

    on the batch of sub-domains a GPU hosts.  The three stencils run back to back on one stream (the
    vertical kernels are column-local and need no exchange); ``step()`` can be captured into a CUDA graph.
    Algorithmic bytes/point (fp64): 48.1 + 16 + 32 = 96.1.
    """

    def __init__(self, transport: FvTransport, ptop: float = 1.0, fused: bool = False):
        """``fused``: use ``remap_delp`` (pe_prefix folded into the remap; pe1 is then not written)."""
        self.transport = transport
        self.ptop = float(ptop)
        self.fused = fused
        self._calls = {}

    def step(self, q, crx, xfx, cry, yfx, rarea, delp, pe2, q_adv, pe1, q_new) -> None:
        key = tuple(t.data_ptr() for t in (delp, pe2, q_adv, pe1, q_new))
        if key not in self._calls:
            from .. import _abi
            from ..fields import shape3

            ni, nj, nk, nb = shape3(delp)
            nk2 = shape3(q_new)[2]
            prec = _abi.precision_of(delp)
            if self.fused:
                self._calls[key] = (
                    None,
                    _abi.prepare("remap_delp", prec, dict(ni=ni, nj=nj, nk1=nk, nk2=nk2, nb=nb, ptop=self.ptop,
                                                          delp=delp, q1=q_adv, pe2=pe2, q2=q_new)),
                )  # fmt: skip
            else:
                self._calls[key] = (
                    _abi.prepare("pe_prefix", prec, dict(ni=ni, nj=nj, nk=nk, nb=nb, ptop=self.ptop, delp=delp, pe=pe1)),
                    _abi.prepare("remap", prec, dict(ni=ni, nj=nj, nk1=nk, nk2=nk2, nb=nb, pe1=pe1, q1=q_adv, pe2=pe2, q2=q_new)),
                )  # fmt: skip
        pe_prefix, remap = self._calls[key]
        self.transport.step(q, crx, xfx, cry, yfx, rarea, q_adv)
        if pe_prefix is not None:
            pe_prefix()
        remap()
