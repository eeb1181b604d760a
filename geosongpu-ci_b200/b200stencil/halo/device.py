"""Library-owned halo exchange over NVLink peer memory: the Python face of ``b2s_halo_*`` (csrc/halo_ctx.cu).

Rendezvous, peer mapping, the neighbour handshake and the pull all live behind the C-ABI
(include/b200stencil.h "multi-GPU halo exchange lifecycle"); this module only

* picks a session name every rank agrees on (the analogue of the communicator the reference passes through its
  bridge, /root/reference/src/tcn/py_ftn_interface/argument.py:54-86),
* wraps the symmetric allocation in a torch tensor (device storage only),
* turns the partitioner's links into the host table ``b2s_halo_plan`` wants.

A halo update on the caller's stream is a one-block handshake kernel (announce, await the neighbours) followed by the
strip copies; forked, it is ONE kernel with the handshake inside.  ``start()`` forks it onto the context's own stream so it overlaps
whatever the caller launches before ``wait()``; with ``gated=True`` the kernel opens one gate per sub-domain as its
halos land and a gated stencil (``stencils.prepare_fv_tp2d_gated``) computes sub-domain b while the halos of b+1..
are still in flight.
A C or Fortran caller drives the same entry points without Python (tests/c_abi/abi_driver.c).
"""
from __future__ import annotations

import os
import uuid
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _abi
from .partitioner import CubedSpherePartitioner

PLAN_WORDS = 12


def make_session(group=None) -> str:
    """A name unique to this run and equal on every rank.

    With an initialised torch.distributed group (an optional bootstrap helper, nothing on the data path) rank 0
    draws a uuid and broadcasts it; otherwise B2S_SESSION, else the launcher's rendezvous identity."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            box = [uuid.uuid4().hex if dist.get_rank(group) == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            return f"{box[0]}"
    except Exception:
        pass
    if os.environ.get("B2S_SESSION"):
        return os.environ["B2S_SESSION"]
    env = os.environ
    return "_".join(str(env.get(k, "x")) for k in ("MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID", "TORCHELASTIC_RESTART_COUNT")).replace("/", "-")


class _RawDeviceMemory:
    """``__cuda_array_interface__`` holder for a library-owned allocation (torch keeps it alive with the tensor)."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
        self._owner = owner


def _padded(ni: int, dtype: torch.dtype) -> int:
    per = 16 // torch.empty((), dtype=dtype).element_size()
    return (ni + per - 1) // per * per


class HaloContext:
    """One rank of a halo-exchange session (``b2s_halo_init`` ... ``b2s_halo_finalize``)."""

    def __init__(self, rank: int = 0, world: int = 1, device: Optional[int] = None, session: Optional[str] = None):
        ffi, lib = _abi.load()
        self._ffi, self._lib = ffi, lib
        self.rank, self.world = int(rank), int(world)
        self.device = int(device if device is not None else torch.cuda.current_device())
        _abi.ensure_init(self.device)
        if self.world > 1 and not session:
            session = make_session()
        out = ffi.new("int64_t*")
        _abi.check("b2s_halo_init", lib.b2s_halo_init((session or "").encode(), self.rank, self.world, self.device, out))
        self.handle = int(out[0])
        self._gate = None
        self._staging = {}  # data_ptr of a field -> (offset, length) in elements of the staging area behind it

    # ---- memory ---------------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> int:
        """Collective symmetric allocation; returns this rank's device address."""
        p = self._ffi.new("void**")
        _abi.check("b2s_halo_alloc", self._lib.b2s_halo_alloc(self.handle, int(nbytes), p))
        return int(self._ffi.cast("uintptr_t", p[0]))

    def free(self, ptr: int) -> None:
        _abi.check("b2s_halo_free", self._lib.b2s_halo_free(self.handle, self._ffi.cast("void*", int(ptr))))

    def peer_ptr(self, ptr: int, peer: int) -> int:
        p = self._ffi.new("void**")
        _abi.check("b2s_halo_peer_ptr", self._lib.b2s_halo_peer_ptr(self.handle, self._ffi.cast("void*", int(ptr)), int(peer), p))
        return int(self._ffi.cast("uintptr_t", p[0]))

    def _tensor(self, ptr: int, nbytes: int, dtype: torch.dtype) -> torch.Tensor:
        raw = torch.as_tensor(_RawDeviceMemory(ptr, nbytes, self), device=torch.device("cuda", self.device))
        return raw.view(dtype)

    def field(self, shape_ijk: Sequence[int], batch: int, dtype=torch.float64, fill: Optional[float] = 0.0,
              part: Optional[CubedSpherePartitioner] = None, n_gpus: Optional[int] = None) -> torch.Tensor:
        """A batch field [b, i, j, k] (i-fastest, rows padded to 16 bytes) in symmetric memory (collective).

        ``part``: also reserve, behind the field in the SAME allocation, the staging area :meth:`plan` needs to let the
        peers deliver their strips packed (:func:`build_plan_table`); without it the crossing strips are pushed in place."""
        ni, nj, nk = (int(s) for s in shape_ijk)
        nip = _padded(ni, dtype)
        numel = int(batch) * nk * nj * nip
        es = torch.empty((), dtype=dtype).element_size()
        align = 128 // es
        staging_at = (numel + align - 1) // align * align
        staging = staging_elements(part, self.world if n_gpus is None else n_gpus, nk) if part is not None else 0
        total = staging_at + staging if staging else numel
        flat = self._tensor(self.alloc(total * es), total * es, dtype)
        if fill is not None:
            flat.fill_(fill)
        out = flat[:numel].view(int(batch), nk, nj, nip).permute(0, 3, 2, 1)[:, :ni]
        if staging:
            self._staging[out.data_ptr()] = (staging_at, staging)
        return out

    # ---- exchange -------------------------------------------------------------------------------
    @property
    def gate(self) -> torch.Tensor:
        """int32[66] device words the gated stencils poll (``b2s_halo_gate``): one flag per sub-domain, CTA counter, status."""
        if self._gate is None:
            p = self._ffi.new("int**")
            _abi.check("b2s_halo_gate", self._lib.b2s_halo_gate(self.handle, p))
            self._gate = self._tensor(int(self._ffi.cast("uintptr_t", p[0])), 66 * 4, torch.int32)
        return self._gate

    def plan(self, field: torch.Tensor, part: CubedSpherePartitioner, n_gpus: Optional[int] = None, gpu: Optional[int] = None,
             ranks: Optional[Sequence[int]] = None, push: bool = False) -> "HaloExchange":
        """Bind the links that fill this GPU's halos to ``field`` (a tensor returned by :meth:`field`).

        ``ranks[g]`` = session rank that hosts GPU ``g`` of the decomposition (default: identity).  ``push`` (every rank
        alike): the table also carries this GPU's outgoing strips, so that the ungated exchange pushes what crosses
        NVLink instead of pulling it (:func:`build_plan_table`; packed, if :meth:`field` reserved a staging area);
        gated and fused exchanges pull regardless.  Off by default: on 2 and 8 B200s the pushed exchanges measured
        3 - 19 % SLOWER per step than the in-place pull (profiles/README.md, round 2) -- each rank's update then ends
        only after every neighbour has started its own and delivered, a tighter coupling than the pull's single
        announcement per neighbour."""
        n_gpus = self.world if n_gpus is None else n_gpus
        gpu = self.rank if gpu is None else gpu
        ranks = list(ranks) if ranks is not None else list(range(n_gpus))
        at = self._staging.get(field.data_ptr())
        if at is not None and at[1] < staging_elements(part, n_gpus, int(field.shape[3])):
            raise ValueError("the staging area reserved behind this field is too small for this partitioner")
        table = build_plan_table(part, n_gpus, gpu, field, ranks, push=bool(push) and n_gpus > 1,
                                 staging_offset=at[0] if at is not None else None)
        out = self._ffi.new("int*")
        tbl = np.ascontiguousarray(table, dtype=np.int64)
        _abi.check("b2s_halo_plan", self._lib.b2s_halo_plan(
            self.handle, self._ffi.cast("void*", field.data_ptr()), field.element_size(), int(field.shape[3]), len(tbl),
            self._ffi.cast("int64_t*", tbl.ctypes.data) if len(tbl) else self._ffi.NULL, out))  # fmt: skip
        return HaloExchange(self, int(out[0]), field)

    def barrier(self) -> None:
        _abi.check("b2s_halo_barrier", self._lib.b2s_halo_barrier(self.handle))

    def status(self):
        """(exchanges completed, status bits); host-synchronising.  Status != 0: a device-side wait timed out."""
        e, s = self._ffi.new("int*"), self._ffi.new("int*")
        _abi.check("b2s_halo_status", self._lib.b2s_halo_status(self.handle, e, s))
        return int(e[0]), int(s[0])

    def trace(self):
        """Device timeline (ns, relative to the exchange start) of the last exchange + gated stencil; diagnostics."""
        out = self._ffi.new("int64_t[6]")
        _abi.check("b2s_halo_trace", self._lib.b2s_halo_trace(self.handle, out))
        t = [int(out[i]) for i in range(6)]
        names = ("exchange_start", "exchange_end", "gate0_open", "stencil_start", "stencil_gate0", "stencil_end")
        return {n: (v - t[0] if v else None) for n, v in zip(names, t)}

    def check(self) -> None:
        epoch, status = self.status()
        if status:
            raise RuntimeError(f"halo exchange: a device-side wait gave up (status {status}, epoch {epoch}): "
                               "a neighbour did not announce its field or the gate never opened")  # fmt: skip

    def finalize(self) -> None:
        if self.handle:
            h, self.handle = self.handle, 0
            self._gate = None
            _abi.check("b2s_halo_finalize", self._lib.b2s_halo_finalize(h))


LINK_OUT = 1 << 32     # b2s_halo_plan: the row is an outgoing strip, [10] = rank that owns the destination sub-domain
LINK_PUSHED = 1 << 33  # b2s_halo_plan: an incoming strip its owner pushes
LINK_STAGED = 1 << 34  # b2s_halo_plan: same-rank unpack of a strip rank [10] delivered into this rank's staging area


def _crossing_links(part: CubedSpherePartitioner, n_gpus: int, gpu: int):
    """Links whose destination is on ``gpu`` and whose source is on another GPU, in the partitioner's global order --
    every rank enumerates them alike, which is what lets the two ends of a strip agree on its place in the staging area."""
    return [l for l in part.all_links() if part.gpu_of(l.dst, n_gpus) == gpu and part.gpu_of(l.src, n_gpus) != gpu]


def staging_elements(part: CubedSpherePartitioner, n_gpus: int, nk: int) -> int:
    """Elements of staging area a GPU needs for the packed strips its peers deliver (the largest over the GPUs: the
    allocation is symmetric)."""
    if n_gpus <= 1:
        return 0
    return max(sum(l.nd * l.np_ for l in _crossing_links(part, n_gpus, g)) for g in range(n_gpus)) * int(nk)


def build_plan_table(part: CubedSpherePartitioner, n_gpus: int, gpu: int, field: torch.Tensor, ranks: Sequence[int],
                     push: bool = False, staging_offset: Optional[int] = None) -> np.ndarray:
    """int64 [nlinks, 12] host table of ``b2s_halo_plan``: element offsets relative to ``field``'s first element,
    [10] = session rank owning the source sub-domain, [11] = destination sub-domain (batch index).

    ``push``: every strip that crosses GPUs is marked at both ends -- LINK_PUSHED on the incoming row of the GPU that
    owns the destination, and an extra LINK_OUT row ([10] = rank owning the destination) in the table of the GPU that
    owns the source -- so that the ungated exchange pulls the same-GPU strips and lets the owners push the rest
    (stores over NVLink are posted, loads are round trips: csrc/halo_device.cuh, version 3).

    ``staging_offset`` (with ``push``; elements from ``field``'s first element to a staging area inside the same
    symmetric allocation, :func:`staging_elements` long): the owner pushes every crossing strip PACKED -- level by level,
    in the order its own memory is contiguous in -- into the destination rank's staging area, so that everything that
    travels over NVLink is contiguous (a west/east strip in place is 3 elements out of every 3 KB row), and the
    destination rank unpacks it locally (LINK_STAGED rows) once the delivery flag has arrived."""
    from .updater import FieldGeometry

    geo = FieldGeometry(field, part.halo)
    staged = bool(push) and staging_offset is not None

    def segment(g_dst: int, link) -> int:  # first element of the link's packed strip in the staging area of g_dst
        at = 0
        for l in _crossing_links(part, n_gpus, g_dst):
            if l.key == link.key:
                return int(staging_offset) + at * geo.nk
            at += l.nd * l.np_
        raise KeyError(link.key)

    rows = []
    for l in part.all_links():
        g_src, g_dst = part.gpu_of(l.src, n_gpus), part.gpu_of(l.dst, n_gpus)
        if g_dst != gpu and not (push and g_src == gpu):
            continue
        b_src, b_dst = part.local_index(l.src, n_gpus), part.local_index(l.dst, n_gpus)
        src = [geo.cell(b_src, l.si0, l.sj0), geo.step(l.sdi, l.sdj), geo.step(l.spi, l.spj), geo.sk]
        dst = [geo.cell(b_dst, l.di0, l.dj0), geo.step(l.ddi, l.ddj), geo.step(l.dpi, l.dpj), geo.sk]
        size = [l.nd, l.np_]
        crossing = g_src != g_dst
        if g_dst == gpu:  # a strip into one of my sub-domains
            rows.append(src + dst + size + [int(ranks[g_src]), b_dst | (LINK_PUSHED if push and crossing else 0)])
        if not (push and crossing):
            continue
        # packed layout of the strip: one level after the other, depth fastest where the source's depth runs along i
        n = l.nd * l.np_
        packed = [segment(g_dst, l)] + ([1, l.nd] if abs(src[1]) == 1 else [l.np_, 1]) + [n] if staged else None
        if g_src == gpu:  # a strip of mine that a peer needs
            rows.append(src + (packed if staged else dst) + size + [int(ranks[g_dst]), b_dst | LINK_OUT])
        if g_dst == gpu and staged:  # ... and its unpacking at the destination, once delivered
            rows.append(packed + dst + size + [int(ranks[g_src]), b_dst | LINK_STAGED])
    return np.asarray(rows, dtype=np.int64).reshape(-1, PLAN_WORDS)


class HaloExchange:
    """A link table bound to one field: ``update()`` = exchange on the current stream; ``start()/wait()`` = forked."""

    def __init__(self, ctx: HaloContext, plan: int, field: torch.Tensor):
        self.ctx, self.plan, self.field = ctx, plan, field
        self.remote_bytes = int(ctx._lib.b2s_halo_plan_remote_bytes(ctx.handle, plan))
        self.bytes_sent_per_update = self.remote_bytes  # pulled, not sent: the same volume crosses NVLink

    def _stream(self, stream):
        if stream is None:
            stream = torch.cuda.current_stream(self.field.device).cuda_stream
        return self.ctx._ffi.cast("void*", int(stream))

    def update(self, stream=None) -> None:
        _abi.check("b2s_halo_exchange", self.ctx._lib.b2s_halo_exchange(self.ctx.handle, self.plan, self._stream(stream)))

    def start(self, gated: bool = False, stream=None) -> None:
        _abi.check("b2s_halo_exchange_start",
                   self.ctx._lib.b2s_halo_exchange_start(self.ctx.handle, self.plan, 1 if gated else 0, self._stream(stream)))

    def wait(self, stream=None) -> None:
        _abi.check("b2s_halo_exchange_wait", self.ctx._lib.b2s_halo_exchange_wait(self.ctx.handle, self._stream(stream)))
