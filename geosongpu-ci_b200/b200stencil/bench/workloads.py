"""Synthetic device workloads for every stencil at the BASELINE configs.

Inputs are generated ON THE DEVICE with torch's RNG following the recipes of oracle/inputs.py
(same distributions; the bit-exact seeded NumPy generators are for the parity tests, these are
for timing at sizes where host generation would take minutes).  Each workload knows its
algorithmic bytes per point (fp64 figures of SURVEY.md 8d scale with the element size).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import torch

from .. import fields, stencils

CONFIGS = {
    # name: (tiles, n, nk)
    "cfg1_24x24x72": (1, 24, 72),
    "C96x72": (6, 96, 72),
    "C180x72": (6, 180, 72),
    "C384x72": (6, 384, 72),
    "C720x137": (6, 720, 137),
}


@dataclass
class Workload:
    name: str
    stencil: str
    points: int  # grid cells x levels per launch
    bytes_per_point: float  # algorithmic, for this dtype
    run: Callable[[int], None]  # run(slot)
    slots: int
    notes: str = ""
    keep: List[object] = field(default_factory=list)

    @property
    def bytes_per_launch(self) -> float:
        return self.points * self.bytes_per_point


def _rand(shape, dtype, nb, lo=0.0, hi=1.0, gen=None):
    t = fields.empty(shape, dtype, batch=nb)
    t.uniform_(lo, hi, generator=gen)
    return t


def _slots_for(bytes_per_set: float, want_bytes: float = 2.5 * 126e6, cap: int = 8) -> int:
    return int(max(1, min(cap, math.ceil(want_bytes / max(bytes_per_set, 1.0)))))


def make(stencil: str, tiles: int, n: int, nk: int, dtype=torch.float64, seed: int = 20240724,
         ni: Optional[int] = None, nj: Optional[int] = None, slots: Optional[int] = None) -> Workload:
    """Build the device workload of one stencil on ``tiles`` sub-domains of ni x nj x nk."""
    ni = ni or n
    nj = nj or n
    es = torch.empty((), dtype=dtype).element_size()
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    cols = tiles * ni * nj
    pts = cols * nk
    shp3, shp2 = (ni, nj, nk), (ni, nj)
    idt = torch.int64 if dtype == torch.float64 else torch.int32
    name = f"{stencil}[{tiles}x{ni}x{nj}x{nk},{'f64' if es == 8 else 'f32'}]"

    if stencil == "top_of_column":
        bpp = es + 2 * es / nk
        ns = slots or _slots_for(pts * bpp)
        sets = [(_rand(shp3, dtype, tiles, 0, 1000, g), fields.zeros(shp2, dtype, batch=tiles), fields.empty(shp3, dtype, batch=tiles)) for _ in range(ns)]
        return Workload(name, stencil, pts, bpp, lambda s: stencils.top_of_column(*sets[s]), ns, keep=sets)

    if stencil == "while_in_function":
        bpp = 2 * es
        ns = slots or _slots_for(pts * bpp)
        sets = []
        for _ in range(ns):
            f = _rand(shp3, dtype, tiles, 0.0, 3.999, g)
            hit = torch.rand(f.shape, device="cuda", generator=g) < (2.0 / nk)
            f[hit] = 4.0 + 38.0 * torch.rand(int(hit.sum()), device="cuda", generator=g).to(dtype)
            f[..., nk - 1] = 42.0
            sets.append((f, fields.empty(shp3, dtype, batch=tiles)))
        return Workload(name, stencil, pts, bpp, lambda s: stencils.while_in_function(*sets[s]), ns, keep=sets)

    if stencil == "hybrid_index_2dout":
        bpp = 2 * es + 2 * es / nk
        ns = slots or _slots_for(pts * bpp)
        sets = []
        for _ in range(ns):
            data = _rand(shp3, dtype, tiles, 800, 900, g).floor_()
            kmask = fields.empty(shp3, dtype, batch=tiles)
            kmask[...] = torch.arange(nk, device="cuda", dtype=dtype)
            kidx = _rand(shp2, dtype, tiles, 0, nk, g).floor_().clamp_(0, nk - 1)
            sets.append((data, kmask, kidx, fields.zeros(shp2, dtype, batch=tiles)))
        return Workload(name, stencil, pts, bpp, lambda s: stencils.hybrid_index_2dout(*sets[s]), ns, keep=sets,
                        notes="algorithmic bytes count data_field in full (SURVEY 8d); the kernel only touches the sectors holding a match")

    if stencil in ("find_klcl", "cloud_top", "saturation_adjust"):
        k = torch.arange(nk, device="cuda", dtype=torch.float64)
        pcol = (100.0 * (100.0 + 900.0 * (k + 0.5) / nk)).to(dtype)
        if stencil == "find_klcl":
            bpp = es + 3 * es / nk
            ns = slots or _slots_for(pts * bpp)
            sets = []
            for _ in range(ns):
                p = fields.empty(shp3, dtype, batch=tiles)
                p[...] = pcol
                plcl = _rand(shp2, dtype, tiles, 600e2, 950e2, g)
                sets.append((p, plcl, fields.zeros(shp2, idt, batch=tiles), fields.zeros(shp2, dtype, batch=tiles)))
            return Workload(name, stencil, pts, bpp, lambda s: stencils.find_klcl(*sets[s]), ns, keep=sets,
                            notes="algorithmic bytes = full column; the warp-level early exit reads only the levels below the LCL")
        if stencil == "cloud_top":
            bpp = es + es / nk
            ns = slots or _slots_for(pts * bpp)
            sets = []
            for _ in range(ns):
                ql = fields.empty(shp3, dtype, batch=tiles)
                ql.normal_(0.0, 1e-4, generator=g).clamp_(min=0.0)
                ql[..., : nk // 2] = 0.0  # clear sky aloft: the search runs half the column
                sets.append((ql, fields.zeros(shp2, idt, batch=tiles)))
            return Workload(name, stencil, pts, bpp, lambda s: stencils.cloud_top(*sets[s]), ns, keep=sets,
                            notes="algorithmic bytes = full column; early exit after the first cloudy level")
        bpp = 7 * es
        ns = slots or _slots_for(pts * bpp)
        sets = []
        for _ in range(ns):
            p = fields.empty(shp3, dtype, batch=tiles)
            p[...] = pcol
            T = fields.empty(shp3, dtype, batch=tiles)
            T.normal_(0.0, 2.0, generator=g)
            T += (210.0 + 90.0 * k / nk).to(dtype)
            es_ = 611.2 * torch.exp(17.67 * (T - 273.15) / (T - 29.65))
            qs = 0.622 * es_ / (p - 0.378 * es_)
            q = _rand(shp3, dtype, tiles, 0, 1.2, g).mul_(qs)
            ql = fields.empty(shp3, dtype, batch=tiles)
            ql.normal_(0.0, 1e-4, generator=g).clamp_(min=0.0)
            sets.append((T, q, ql, p))
            del es_, qs
        return Workload(name, stencil, pts, bpp, lambda s: stencils.saturation_adjust(*sets[s]), ns, keep=sets)

    if stencil == "fv_tp2d":
        bpp = 6 * es + es / nk
        ns = slots or _slots_for(pts * bpp, cap=4)
        sets = []
        for _ in range(ns):
            q = _rand((ni + 6, nj + 6, nk), dtype, tiles, 0.5, 1.5, g)
            crx = _rand((ni + 1, nj, nk), dtype, tiles, -0.9, 0.9, g)
            cry = _rand((ni, nj + 1, nk), dtype, tiles, -0.9, 0.9, g)
            xfx = _rand((ni + 1, nj, nk), dtype, tiles, 0.9, 1.1, g).mul_(crx)  # in place: keeps the padded rows
            yfx = _rand((ni, nj + 1, nk), dtype, tiles, 0.9, 1.1, g).mul_(cry)
            rarea = _rand(shp2, dtype, tiles, 0.9, 1.1, g)
            sets.append((q, crx, xfx, cry, yfx, rarea, fields.empty(shp3, dtype, batch=tiles)))
        return Workload(name, stencil, pts, bpp, lambda s: stencils.fv_tp2d(*sets[s]), ns, keep=sets)

    if stencil == "fv_tp2d_split":  # FV3 fv_tp_2d: q, crx, xfx, cry, yfx in (+ area, rarea IJ); q_out out
        bpp = 6 * es + 2 * es / nk
        ns = slots or _slots_for(pts * bpp, cap=4)
        sets = []
        for _ in range(ns):
            q = _rand((ni + 6, nj + 6, nk), dtype, tiles, 0.5, 1.5, g)
            crx = _rand((ni + 1, nj + 6, nk), dtype, tiles, -0.45, 0.45, g)
            cry = _rand((ni + 6, nj + 1, nk), dtype, tiles, -0.45, 0.45, g)
            xfx = _rand((ni + 1, nj + 6, nk), dtype, tiles, 0.9, 1.1, g).mul_(crx)
            yfx = _rand((ni + 6, nj + 1, nk), dtype, tiles, 0.9, 1.1, g).mul_(cry)
            area = _rand((ni + 6, nj + 6), dtype, tiles, 0.9, 1.1, g)
            rarea = fields.empty(shp2, dtype, batch=tiles)
            rarea[...] = 1.0 / area[:, 3:3 + ni, 3:3 + nj]
            sets.append((q, crx, xfx, cry, yfx, area, rarea, fields.empty(shp3, dtype, batch=tiles)))
        return Workload(name, stencil, pts, bpp, lambda s: stencils.fv_tp2d_split(*sets[s]), ns, keep=sets)

    if stencil == "remap_delp":  # pe_prefix fused into the remap: delp, q1, pe2 in; q2 out
        bpp = 4 * es
        ns = slots or _slots_for(pts * bpp)
        sets = []
        for _ in range(ns):
            delp = _rand(shp3, dtype, tiles, 0.5 * 1e5 / nk, 1.5 * 1e5 / nk, g)
            pe1 = fields.empty((ni, nj, nk + 1), dtype, batch=tiles)
            stencils.pe_prefix(delp, 1.0, pe1)
            sig = (torch.arange(nk + 1, device="cuda", dtype=torch.float64) / nk).to(dtype)
            pe2 = fields.empty((ni, nj, nk + 1), dtype, batch=tiles)
            pe2[...] = pe1[..., :1] + (pe1[..., -1:] - pe1[..., :1]) * sig
            pe2[..., -1] = pe1[..., -1]
            del pe1
            q1 = _rand(shp3, dtype, tiles, 1.0, 2.0, g)
            sets.append((delp, 1.0, q1, pe2, fields.empty(shp3, dtype, batch=tiles)))
        return Workload(name, stencil, pts, bpp, lambda s: stencils.remap_delp(*sets[s]), ns, keep=sets)

    if stencil == "remap_ppm":  # FV3-style PPM remap: pe1, q1, pe2 in; q2 out
        bpp = 4 * es
        ns = slots or _slots_for(pts * bpp)
        sets = []
        for _ in range(ns):
            delp = _rand(shp3, dtype, tiles, 0.5 * 1e5 / nk, 1.5 * 1e5 / nk, g)
            pe1 = fields.empty((ni, nj, nk + 1), dtype, batch=tiles)
            stencils.pe_prefix(delp, 1.0, pe1)
            del delp
            sig = (torch.arange(nk + 1, device="cuda", dtype=torch.float64) / nk).to(dtype)
            pe2 = fields.empty((ni, nj, nk + 1), dtype, batch=tiles)
            pe2[...] = pe1[..., :1] + (pe1[..., -1:] - pe1[..., :1]) * sig
            pe2[..., -1] = pe1[..., -1]
            pm = (0.5 * (pe1[..., 1:] + pe1[..., :-1]) / pe1[..., -1:]).to(torch.float64)
            q1 = fields.empty(shp3, dtype, batch=tiles)
            q1[...] = (1.0 + 0.6 * torch.sin(2 * math.pi * pm) + 0.3 * torch.sin(7 * math.pi * pm * pm)).to(dtype)
            q1 += _rand(shp3, dtype, tiles, -0.02, 0.02, g)
            del pm
            sets.append((pe1, q1, pe2, fields.empty(shp3, dtype, batch=tiles)))
        return Workload(name, stencil, pts, bpp, lambda s: stencils.remap_ppm(*sets[s]), ns, keep=sets)

    if stencil in ("pe_prefix", "remap", "tridiag"):
        if stencil == "pe_prefix":
            bpp = 2 * es
            ns = slots or _slots_for(pts * bpp)
            sets = [(_rand(shp3, dtype, tiles, 0.5 * 1e5 / nk, 1.5 * 1e5 / nk, g), 1.0, fields.empty((ni, nj, nk + 1), dtype, batch=tiles)) for _ in range(ns)]
            return Workload(name, stencil, pts, bpp, lambda s: stencils.pe_prefix(*sets[s]), ns, keep=sets)
        if stencil == "remap":
            bpp = 4 * es
            ns = slots or _slots_for(pts * bpp)
            sets = []
            for _ in range(ns):
                delp = _rand(shp3, dtype, tiles, 0.5 * 1e5 / nk, 1.5 * 1e5 / nk, g)
                pe1 = fields.empty((ni, nj, nk + 1), dtype, batch=tiles)
                stencils.pe_prefix(delp, 1.0, pe1)
                sig = (torch.arange(nk + 1, device="cuda", dtype=torch.float64) / nk).to(dtype)
                pe2 = fields.empty((ni, nj, nk + 1), dtype, batch=tiles)
                pe2[...] = pe1[..., :1] + (pe1[..., -1:] - pe1[..., :1]) * sig
                pe2[..., -1] = pe1[..., -1]
                q1 = _rand(shp3, dtype, tiles, 1.0, 2.0, g)
                sets.append((pe1, q1, pe2, fields.empty(shp3, dtype, batch=tiles)))
                del delp
            return Workload(name, stencil, pts, bpp, lambda s: stencils.remap(*sets[s]), ns, keep=sets)
        bpp = 9 * es
        ns = slots or _slots_for(pts * bpp, cap=3)
        sets = []
        for _ in range(ns):
            a = _rand(shp3, dtype, tiles, 0, 1, g).neg_()
            c = _rand(shp3, dtype, tiles, 0, 1, g).neg_()
            b = _rand(shp3, dtype, tiles, 2, 3, g)
            d = _rand(shp3, dtype, tiles, -1, 1, g)
            sets.append((a, b, c, d, fields.empty(shp3, dtype, batch=tiles), fields.empty(shp3, dtype, batch=tiles)))
        return Workload(name, stencil, pts, bpp, lambda s: stencils.tridiag(*sets[s]), ns, keep=sets)

    raise KeyError(stencil)


ALL_STENCILS = [
    "top_of_column", "while_in_function", "hybrid_index_2dout", "find_klcl", "saturation_adjust", "cloud_top",
    "fv_tp2d", "fv_tp2d_split", "pe_prefix", "remap", "remap_delp", "remap_ppm", "tridiag",
]  # fmt: skip

# stencil -> BASELINE config it is quoted on
DEFAULT_CONFIG: Dict[str, str] = {
    "top_of_column": "C96x72", "while_in_function": "C96x72", "hybrid_index_2dout": "C96x72",
    "find_klcl": "C180x72", "saturation_adjust": "C180x72", "cloud_top": "C180x72",
    "fv_tp2d": "C384x72", "fv_tp2d_split": "C384x72", "pe_prefix": "C720x137", "remap": "C720x137", "remap_delp": "C720x137", "remap_ppm": "C720x137", "tridiag": "C720x137",
}  # fmt: skip
