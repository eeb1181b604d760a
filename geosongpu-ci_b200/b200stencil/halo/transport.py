"""One flux-form transport step on a GPU's batch of sub-domains: halo update of q + fv_tp2d.

This is the unit bench.py times (BASELINE config 4: "FV3-style horizontal finite-volume
flux/advection stencil with 3-point halo on C384x72, halo exchange at 2/4/8 GPUs").  With more than
one GPU the exchange runs on a communication stream while the interior rectangle -- whose stencil
never reads a halo cell -- is computed; the boundary frame follows once the halos have landed.
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
    def __init__(self, part: CubedSpherePartitioner, n_gpus: int, gpu: int, process_group=None,
                 overlap: bool = True, side: int = 32):
        self.part, self.n_gpus, self.gpu = part, n_gpus, gpu
        self.updater = HaloUpdater(part, n_gpus, gpu, process_group=process_group)
        self.overlap = overlap and bool(self.updater.plan.peers)
        self.interior, self.frame = split_regions(part.nx, part.ny, part.halo, side)
        self.kernel_launches_per_step = 0

    def step(self, q, crx, xfx, cry, yfx, rarea, q_out, q_out_halo: int = 0) -> None:
        """q (halo-padded batch field) -> q_out; q's halos are refreshed from the neighbours first."""
        if not self.overlap:
            self.updater.update(q)
            stencils.fv_tp2d(q, crx, xfx, cry, yfx, rarea, q_out, q_out_halo=q_out_halo)
            return
        self.updater.start(q)
        if self.interior[1] > self.interior[0]:
            stencils.fv_tp2d(q, crx, xfx, cry, yfx, rarea, q_out, region=self.interior, q_out_halo=q_out_halo)
        self.updater.wait()
        for rect in self.frame:
            stencils.fv_tp2d(q, crx, xfx, cry, yfx, rarea, q_out, region=rect, q_out_halo=q_out_halo)
