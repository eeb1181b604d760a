"""Stencil entry points on device fields (torch CUDA tensors, i-fastest; see fields.py).

One function per row of SURVEY.md 8(a).  Argument names and order follow the reference stencil
signatures where a reference exists (dsl_patterns/*.py) and the oracle spec otherwise.  Every
function only validates, marshals and calls the C-ABI (``_abi.call``); the arithmetic is in
csrc/*.cu.  Fields may carry a leading batch axis ``b`` (tiles / sub-domains).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _abi
from .fields import shape3

FV_HALO = 3


def _index_dtype(t: torch.Tensor) -> torch.dtype:
    return torch.int64 if t.dtype == torch.float64 else torch.int32


def _expect(fn: str, lead: torch.Tensor, **args) -> None:
    """Shape check of every argument against the domain derived from the leading field.

    The C-ABI carries no extents with a pointer (like the reference's bridge, SURVEY.md 8b "no shape information
    travels with a pointer"), so a wrongly shaped tensor would be a silent out-of-bounds access on the device:
    the host wrapper is the only place that can refuse it.  ``args``: name -> (tensor | None, expected [i,j(,k)] shape);
    in a batched call (``lead`` indexed [b,i,j,k]) a field carries the same ``b`` or no batch axis at all (shared)."""
    batched = lead.dim() == 4
    nb = lead.shape[0] if batched else None
    for name, (t, shape) in args.items():
        if t is None:
            continue
        core = tuple(int(x) for x in shape)
        want = ((nb,) if batched else ()) + core
        # a field without the batch axis is shared by every sub-domain of a batched call (batch stride 0)
        if tuple(t.shape) != want and not (batched and tuple(t.shape) == core):
            raise ValueError(f"{fn}: {name} has shape {tuple(t.shape)}, expected {want} "
                             f"({'[b,]' if batched else ''}i,j{',k' if len(shape) == 3 else ''} for this domain)")  # fmt: skip


def _expect_flat(fn: str, name: str, t, numel: int, dtype=None) -> None:
    if t is None:
        return
    if t.numel() < numel:
        raise ValueError(f"{fn}: {name} holds {t.numel()} elements, the call reads {numel}")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{fn}: {name} must be {dtype}, got {t.dtype}")


def top_of_column(PLEmb, PLEmb_top, out_field, stream: Optional[int] = None) -> None:
    """dsl_patterns/Do__get_top_of_the_column.py:33-38 -- K1 (csrc/k_patterns.cu)."""
    ni, nj, nk, nb = shape3(PLEmb)
    _expect("top_of_column", PLEmb, PLEmb_top=(PLEmb_top, (ni, nj)), out_field=(out_field, (ni, nj, nk)))
    _abi.call(
        "top_of_column", _abi.precision_of(PLEmb),
        dict(ni=ni, nj=nj, nk=nk, nb=nb, PLEmb=PLEmb, PLEmb_top=PLEmb_top, out_field=out_field), stream,
    )  # fmt: skip


def while_in_function(in_field, out_field, threshold: float = 4.0, undefined_count=None, stream=None) -> None:
    """dsl_patterns/Do__while_in_gt_functions.py:22-32 -- K2.

    ``undefined_count`` (optional int64 device tensor of one element, caller-zeroed) receives the
    number of points whose search ran off the column (undefined behaviour in the reference).
    """
    ni, nj, nk, nb = shape3(in_field)
    _expect("while_in_function", in_field, out_field=(out_field, (ni, nj, nk)))
    _expect_flat("while_in_function", "undefined_count", undefined_count, 1, torch.int64)
    _abi.call(
        "while_in_function", _abi.precision_of(in_field),
        dict(ni=ni, nj=nj, nk=nk, nb=nb, threshold=float(threshold), in_field=in_field, out_field=out_field,
             undefined_count=undefined_count), stream,
    )  # fmt: skip


def hybrid_index_2dout(data_field, k_mask, k_index_desired, out_field, stream=None) -> None:
    """dsl_patterns/WIP__hybrid_index_2dout.py:34-42 -- K3."""
    ni, nj, nk, nb = shape3(data_field)
    _expect("hybrid_index_2dout", data_field, k_mask=(k_mask, (ni, nj, nk)), k_index_desired=(k_index_desired, (ni, nj)),
            out_field=(out_field, (ni, nj)))
    _abi.call(
        "hybrid_index_2dout", _abi.precision_of(data_field),
        dict(ni=ni, nj=nj, nk=nk, nb=nb, data_field=data_field, k_mask=k_mask, k_index_desired=k_index_desired,
             out_field=out_field), stream,
    )  # fmt: skip


def find_klcl(PLmb, PLCL, KLCL, PLmb_at_KLCL, stream=None) -> None:
    """S4a (spec: oracle/numpy_oracle.py find_klcl) -- K4a (csrc/k_moist.cu)."""
    ni, nj, nk, nb = shape3(PLmb)
    _expect("find_klcl", PLmb, PLCL=(PLCL, (ni, nj)), KLCL=(KLCL, (ni, nj)), PLmb_at_KLCL=(PLmb_at_KLCL, (ni, nj)))
    _abi.call(
        "find_klcl", _abi.precision_of(PLmb),
        dict(ni=ni, nj=nj, nk=nk, nb=nb, PLmb=PLmb, PLCL=PLCL, PLmb_at_KLCL=PLmb_at_KLCL, KLCL=KLCL), stream,
    )  # fmt: skip


def saturation_adjust(T, q, ql, p, stream=None) -> None:
    """S4b (spec: oracle/numpy_oracle.py saturation_adjust) -- K4b; in place on T, q, ql."""
    ni, nj, nk, nb = shape3(T)
    _expect("saturation_adjust", T, q=(q, (ni, nj, nk)), ql=(ql, (ni, nj, nk)), p=(p, (ni, nj, nk)))
    _abi.call("saturation_adjust", _abi.precision_of(T), dict(ni=ni, nj=nj, nk=nk, nb=nb, p=p, T=T, q=q, ql=ql), stream)


def cloud_top(ql, ktop, ql_min: float = 1.0e-8, stream=None) -> None:
    """S4c (spec: oracle/numpy_oracle.py cloud_top) -- K4c."""
    ni, nj, nk, nb = shape3(ql)
    _expect("cloud_top", ql, ktop=(ktop, (ni, nj)))
    _abi.call("cloud_top", _abi.precision_of(ql), dict(ni=ni, nj=nj, nk=nk, nb=nb, ql_min=float(ql_min), ql=ql, ktop=ktop), stream)


def _expect_fv(fn, q, crx, xfx, cry, yfx, rarea, q_out, ni, nj, nk, q_out_halo=0):
    if ni <= 0 or nj <= 0:
        raise ValueError(f"{fn}: q has shape {tuple(q.shape)}; it must carry a {FV_HALO}-cell halo around a non-empty domain")
    g = 2 * q_out_halo
    _expect(fn, q, crx=(crx, (ni + 1, nj, nk)), xfx=(xfx, (ni + 1, nj, nk)), cry=(cry, (ni, nj + 1, nk)), yfx=(yfx, (ni, nj + 1, nk)),
            rarea=(rarea, (ni, nj)), q_out=(q_out, (ni + g, nj + g, nk)))  # fmt: skip


def prepare_fv_tp2d_gated(q, crx, xfx, cry, yfx, rarea, q_out, gate, q_out_halo: int = 0) -> "_abi.PreparedCall":
    """fv_tp2d on the whole batch overlapped with the halo update of ``q`` in flight (``HaloExchange.start(gated=True)``):
    sub-domain b is computed once ``gate[b]`` (``HaloContext.gate``) says its halos have landed, while those of b+1..
    are still arriving.  ``q`` must be the whole batch field of the exchange.  Same bits as fv_tp2d."""
    h = FV_HALO
    nip, njp, nk, nb = shape3(q)
    ni, nj = nip - 2 * h, njp - 2 * h
    _expect_fv("fv_tp2d_gated", q, crx, xfx, cry, yfx, rarea, q_out, ni, nj, nk, q_out_halo)
    _expect_flat("fv_tp2d_gated", "gate", gate, 66, torch.int32)
    return _abi.prepare(
        "fv_tp2d_gated", _abi.precision_of(q),
        dict(ni=ni, nj=nj, nk=nk, nb=nb, q=q, crx=crx, xfx=xfx, cry=cry, yfx=yfx, rarea=rarea, gate=gate, q_out=q_out),
        origins={"q": (h, h, 0), "q_out": (q_out_halo, q_out_halo, 0)},
    )  # fmt: skip


def prepare_halo_fv_tp2d(halo_exchange, q, crx, xfx, cry, yfx, rarea, q_out, q_out_halo: int = 0) -> "_abi.PreparedCall":
    """The transport step as ONE launch (``b2s_halo_fv_tp2d``): halo update of ``q`` over NVLink peer memory fused with
    fv_tp2d on the whole batch -- every CTA of the persistent stencil grid first takes its share of the exchange
    (neighbour handshake + strip copies), then walks its stencil items in sub-domain order behind the gates the
    exchange opens.  ``halo_exchange`` = the ``HaloContext.plan`` of ``q``.  Same bits as exchange-then-fv_tp2d."""
    h = FV_HALO
    nip, njp, nk, nb = shape3(q)
    ni, nj = nip - 2 * h, njp - 2 * h
    _expect_fv("halo_fv_tp2d", q, crx, xfx, cry, yfx, rarea, q_out, ni, nj, nk, q_out_halo)
    if q.data_ptr() != halo_exchange.field.data_ptr():
        raise ValueError("halo_fv_tp2d: q is not the field the HaloExchange was planned for")
    return _abi.prepare(
        "halo_fv_tp2d", _abi.precision_of(q),
        dict(ctx=halo_exchange.ctx.handle, plan=halo_exchange.plan, ni=ni, nj=nj, nk=nk, nb=nb, crx=crx, xfx=xfx, cry=cry, yfx=yfx,
             rarea=rarea, q=q, q_out=q_out),
        origins={"q": (h, h, 0), "q_out": (q_out_halo, q_out_halo, 0)},
    )  # fmt: skip


def prepare_fv_tp2d(q, crx, xfx, cry, yfx, rarea, q_out, region=None, q_out_halo: int = 0) -> "_abi.PreparedCall":
    """fv_tp2d with its arguments marshalled once (see :class:`_abi.PreparedCall`)."""
    h = FV_HALO
    nip, njp, nk, nb = shape3(q)
    ni, nj = nip - 2 * h, njp - 2 * h
    i0, i1, j0, j1 = (0, ni, 0, nj) if region is None else region
    _expect_fv("fv_tp2d", q, crx, xfx, cry, yfx, rarea, q_out, ni, nj, nk, q_out_halo)
    return _abi.prepare(
        "fv_tp2d", _abi.precision_of(q),
        dict(ni=ni, nj=nj, nk=nk, nb=nb, i0=i0, i1=i1, j0=j0, j1=j1, q=q, crx=crx, xfx=xfx, cry=cry, yfx=yfx,
             rarea=rarea, q_out=q_out),
        origins={"q": (h, h, 0), "q_out": (q_out_halo, q_out_halo, 0)},
    )  # fmt: skip


def fv_tp2d(q, crx, xfx, cry, yfx, rarea, q_out, region=None, q_out_halo: int = 0, stream=None) -> None:
    """S5 (spec: oracle/numpy_oracle.py fv_tp2d) -- K5 (csrc/k_fv*.cu).

    ``q`` carries a 3-cell halo on every horizontal side ([b,] ni+6, nj+6, nk); ``q_out`` is
    compute-domain shaped, or halo-padded like ``q`` when ``q_out_halo`` = 3 (time stepping).
    ``region`` = (i0, i1, j0, j1) restricts the update to a sub-rectangle (interior / boundary
    split for halo-exchange overlap); default is the whole domain.
    """
    prepare_fv_tp2d(q, crx, xfx, cry, yfx, rarea, q_out, region, q_out_halo)(stream)


def fv_tp2d_split(q, crx, xfx, cry, yfx, area, rarea, q_out, fx_out=None, fy_out=None, corner_flags=None, stream=None) -> None:
    """S5b FV3 fv_tp_2d, inner/outer operator splitting (spec: oracle/numpy_oracle.py fv_tp2d_split) -- csrc/k_fv_split.cu.

    ``q`` and ``area`` carry a 3-cell halo on every side INCLUDING the corners ([b,] ni+6, nj+6[, nk]); ``crx``/``xfx``
    are ([b,] ni+1, nj+6, nk) (x-interfaces of the rows -3 .. nj+2), ``cry``/``yfx`` ([b,] ni+6, nj+1, nk); ``rarea`` and
    ``q_out`` are compute-domain shaped; ``fx_out`` ([b,] ni+1, nj, nk) / ``fy_out`` ([b,] ni, nj+1, nk) optionally receive
    the averaged fluxes.  ``corner_flags`` (int32 device tensor, one entry per sub-domain of the batch; see
    ``CubedSpherePartitioner.cube_corner_flags``) marks halo corners that are cube corners: the corner cells of ``q``
    then hold FV3's copy_corners values for x-sweeps (what the halo update writes) and the inner y-sweep takes the
    direction-2 values from the sub-domain's own south / north halo."""
    h = FV_HALO
    nip, njp, nk, nb = shape3(q)
    ni, nj = nip - 2 * h, njp - 2 * h
    _expect("fv_tp2d_split", q, crx=(crx, (ni + 1, nj + 6, nk)), xfx=(xfx, (ni + 1, nj + 6, nk)), cry=(cry, (ni + 6, nj + 1, nk)),
            yfx=(yfx, (ni + 6, nj + 1, nk)), area=(area, (ni + 6, nj + 6)), rarea=(rarea, (ni, nj)), q_out=(q_out, (ni, nj, nk)),
            fx_out=(fx_out, (ni + 1, nj, nk)), fy_out=(fy_out, (ni, nj + 1, nk)))  # fmt: skip
    _expect_flat("fv_tp2d_split", "corner_flags", corner_flags, nb, torch.int32)
    _abi.call(
        "fv_tp2d_split", _abi.precision_of(q),
        dict(ni=ni, nj=nj, nk=nk, nb=nb, q=q, crx=crx, xfx=xfx, cry=cry, yfx=yfx, area=area, rarea=rarea,
             corner_flags=corner_flags, q_out=q_out, fx_out=fx_out, fy_out=fy_out),
        stream, origins={"q": (h, h, 0), "area": (h, h), "crx": (0, h, 0), "xfx": (0, h, 0), "cry": (h, 0, 0), "yfx": (h, 0, 0)},
    )  # fmt: skip


def pe_prefix(delp, ptop: float, pe, stream=None) -> None:
    """S6a (spec: oracle/numpy_oracle.py pe_prefix) -- K6a (csrc/k_vertical.cu); pe has nk+1 levels."""
    ni, nj, nk, nb = shape3(delp)
    _expect("pe_prefix", delp, pe=(pe, (ni, nj, nk + 1)))
    _abi.call("pe_prefix", _abi.precision_of(delp), dict(ni=ni, nj=nj, nk=nk, nb=nb, ptop=float(ptop), delp=delp, pe=pe), stream)


def remap(pe1, q1, pe2, q2, stream=None) -> None:
    """S6b (spec: oracle/numpy_oracle.py remap_column) -- K6b."""
    ni, nj, nk1, nb = shape3(q1)
    nk2 = shape3(q2)[2]
    _expect("remap", q1, pe1=(pe1, (ni, nj, nk1 + 1)), pe2=(pe2, (ni, nj, nk2 + 1)), q2=(q2, (ni, nj, nk2)))
    _abi.call("remap", _abi.precision_of(q1), dict(ni=ni, nj=nj, nk1=nk1, nk2=nk2, nb=nb, pe1=pe1, q1=q1, pe2=pe2, q2=q2), stream)


def remap_delp(delp, ptop: float, q1, pe2, q2, stream=None) -> None:
    """pe_prefix fused into remap (csrc/k_vertical.cu k_remap_delp): bit-identical to
    ``pe_prefix(delp, ptop, pe1); remap(pe1, q1, pe2, q2)`` without materialising pe1."""
    ni, nj, nk1, nb = shape3(q1)
    nk2 = shape3(q2)[2]
    _expect("remap_delp", q1, delp=(delp, (ni, nj, nk1)), pe2=(pe2, (ni, nj, nk2 + 1)), q2=(q2, (ni, nj, nk2)))
    _abi.call(
        "remap_delp", _abi.precision_of(q1),
        dict(ni=ni, nj=nj, nk1=nk1, nk2=nk2, nb=nb, ptop=float(ptop), delp=delp, q1=q1, pe2=pe2, q2=q2), stream,
    )  # fmt: skip


def remap_ppm(pe1, q1, pe2, q2, kord: int = 4, iv: int = 1, stream=None) -> None:
    """S6d FV3-style PPM remap (spec: oracle/numpy_oracle.py ppm_profile + remap_ppm_column) -- csrc/k_remap_ppm.cu.

    ``kord`` 4 | 5 | 6: interior limiter (monotone | positive definite | none); ``iv`` 0 for positive
    definite scalars, 1 otherwise.  Edges must be strictly increasing with ``pe2`` inside ``pe1``."""
    ni, nj, nk1, nb = shape3(q1)
    nk2 = shape3(q2)[2]
    _expect("remap_ppm", q1, pe1=(pe1, (ni, nj, nk1 + 1)), pe2=(pe2, (ni, nj, nk2 + 1)), q2=(q2, (ni, nj, nk2)))
    _abi.call("remap_ppm", _abi.precision_of(q1),
              dict(ni=ni, nj=nj, nk1=nk1, nk2=nk2, nb=nb, kord=int(kord), iv=int(iv), pe1=pe1, q1=q1, pe2=pe2, q2=q2), stream)


def tridiag(a, b, c, d, x, w=None, stream=None) -> None:
    """S6c (spec: oracle/numpy_oracle.py tridiag) -- K6c; ``w`` is scratch shaped like ``x``."""
    ni, nj, nk, nb = shape3(b)
    if w is None:
        w = torch.empty_like(x)
    _expect("tridiag", b, a=(a, (ni, nj, nk)), c=(c, (ni, nj, nk)), d=(d, (ni, nj, nk)), w=(w, (ni, nj, nk)), x=(x, (ni, nj, nk)))
    _abi.call("tridiag", _abi.precision_of(b), dict(ni=ni, nj=nj, nk=nk, nb=nb, a=a, b=b, c=c, d=d, w=w, x=x), stream)


def halo_move(links, nk: int, src, dst, max_strip: int, stream=None) -> None:
    """K7 (csrc/k_halo.cu): run the affine strip copies described by ``links`` (int64 [nlinks, 10]);
    ``max_strip`` = the largest nd*np of the table (sizes the grid without a device read-back)."""
    if int(links.shape[0]) == 0:
        return
    prepare_halo_move(links, nk, src, dst, max_strip)(stream)


def _expect_links(fn: str, links, words: int) -> None:
    if links.dim() != 2 or links.shape[1] != words or links.dtype != torch.int64 or not links.is_contiguous():
        raise ValueError(f"{fn}: links must be a contiguous int64 tensor [nlinks, {words}], got {tuple(links.shape)} {links.dtype}")


def prepare_halo_move(links, nk: int, src, dst, max_strip: int) -> "_abi.PreparedCall":
    _expect_links("halo_move", links, 10)
    return _abi.prepare(
        "halo_move", _abi.precision_of(src),
        dict(nlinks=int(links.shape[0]), nk=int(nk), max_strip=int(max_strip), links=links, src=src, dst=dst),
    )  # fmt: skip


def prepare_halo_pull(links, nk: int, dst, max_strip: int) -> "_abi.PreparedCall":
    """One-kernel halo update over peer memory (csrc/k_halo.cu halo_pull); links int64 [nlinks, 11] with the source
    buffer's base address in word [10].  The caller orders it against the peers' writes (the library-owned exchange,
    ``halo/device.py``, carries the handshake inside its kernel instead)."""
    _expect_links("halo_pull", links, 11)
    return _abi.prepare(
        "halo_pull", _abi.precision_of(dst),
        dict(nlinks=int(links.shape[0]), nk=int(nk), max_strip=int(max_strip), links=links, dst=dst),
    )  # fmt: skip
