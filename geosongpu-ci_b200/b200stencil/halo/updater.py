"""Halo exchange for sub-domain batches resident on GPUs.

Stands where NDSL's HaloUpdater over mpi4py stands in the reference stack (not vendored; the
reference only builds the CUDA-aware OpenMPI/UCX it runs on,
/root/reference/sw_stack/discover/sles15/src/2024.04.00/build_0_on-node.sh:12-44).  B200 design:

* one process per GPU, each hosting ``S = 6*lx*ly / G`` sub-domains as the batch axis of one
  halo-padded field ``[b, i, j, k]`` (i-fastest);
* neighbours on the SAME GPU are filled by one ``halo_move`` kernel straight from the neighbour's
  interior (tile-edge rotation folded into the affine strides) -- no staging;
* neighbours on OTHER GPUs: one ``halo_move`` packs every outgoing strip into one contiguous
  segment per peer GPU, a single grouped NCCL send/recv (``torch.distributed.batch_isend_irecv``)
  moves the segments over NVLink on a dedicated communication stream, one ``halo_move`` unpacks;
* ``start()`` / ``wait()`` split so interior compute overlaps the exchange.

PyTorch provides streams, device memory and the NCCL plumbing only; packing, unpacking and local
copies are kernels of libb200stencil (csrc/k_halo.cu).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .partitioner import CubedSpherePartitioner, Link

LINK_WORDS = 10
Mover = Callable[..., None]  # mover(links, nk, src, dst, max_strip)


def _cuda_mover(links: torch.Tensor, nk: int, src: torch.Tensor, dst: torch.Tensor, max_strip: int = 0) -> None:
    from .. import stencils

    stencils.halo_move(links, nk, src, dst, max_strip)


class FieldGeometry:
    """Strides of a halo-padded batch field [b, i, j, k] (elements), offsets relative to data_ptr."""

    def __init__(self, field: torch.Tensor, halo: int):
        if field.dim() != 4:
            raise ValueError("halo exchange works on batch fields indexed [b, i, j, k]")
        if field.shape[1] > 1 and field.stride(1) != 1:
            raise ValueError("fields must be i-fastest")
        self.sb, self.sj, self.sk = field.stride(0), field.stride(2), field.stride(3)
        self.nk = field.shape[3]
        self.halo = halo
        self.key = (self.sb, self.sj, self.sk, self.nk)

    def cell(self, b: int, i: int, j: int) -> int:
        return b * self.sb + (i + self.halo) + (j + self.halo) * self.sj

    def step(self, di: int, dj: int) -> int:
        return di + dj * self.sj


class HaloPlan:
    """Link tables of one GPU (``gpu`` of ``n_gpus``) for one field geometry."""

    def __init__(self, part: CubedSpherePartitioner, n_gpus: int, gpu: int):
        self.part, self.n_gpus, self.gpu = part, n_gpus, gpu
        self.nsub = part.subdomains_per_gpu(n_gpus)
        links = part.all_links()
        mine = lambda r: part.gpu_of(r, n_gpus) == gpu  # noqa: E731
        self.local: List[Link] = [l for l in links if mine(l.dst) and mine(l.src)]
        self.send: Dict[int, List[Link]] = {}
        self.recv: Dict[int, List[Link]] = {}
        for l in links:  # `links` is sorted by Link.key, so both ends see the same order
            if mine(l.src) and not mine(l.dst):
                self.send.setdefault(part.gpu_of(l.dst, n_gpus), []).append(l)
            if mine(l.dst) and not mine(l.src):
                self.recv.setdefault(part.gpu_of(l.src, n_gpus), []).append(l)
        self.peers = sorted(set(self.send) | set(self.recv))

    def _b(self, rank: int) -> int:
        return self.part.local_index(rank, self.n_gpus)

    def segment_elems(self, links: Sequence[Link], nk: int) -> int:
        return sum(l.nd * l.np_ * nk for l in links)

    def tables(self, geo: FieldGeometry):
        """NumPy int64 tables: local [n,10]; pack [n,10] + per-peer (offset, count); unpack likewise."""
        nk = geo.nk

        def field_side(rank, i0, j0, di, dj, pi, pj):
            return [geo.cell(self._b(rank), i0, j0), geo.step(di, dj), geo.step(pi, pj), geo.sk]

        local = [
            field_side(l.src, l.si0, l.sj0, l.sdi, l.sdj, l.spi, l.spj)
            + field_side(l.dst, l.di0, l.dj0, l.ddi, l.ddj, l.dpi, l.dpj)
            + [l.nd, l.np_]
            for l in self.local
        ]
        pack, unpack = [], []
        send_seg, recv_seg = {}, {}
        off = 0
        for peer in self.peers:
            start = off
            for l in self.send.get(peer, []):
                pack.append(field_side(l.src, l.si0, l.sj0, l.sdi, l.sdj, l.spi, l.spj)
                            + [off, l.np_, 1, l.nd * l.np_] + [l.nd, l.np_])  # fmt: skip
                off += l.nd * l.np_ * nk
            send_seg[peer] = (start, off - start)
        send_total, off = off, 0
        for peer in self.peers:
            start = off
            for l in self.recv.get(peer, []):
                unpack.append([off, l.np_, 1, l.nd * l.np_]
                              + field_side(l.dst, l.di0, l.dj0, l.ddi, l.ddj, l.dpi, l.dpj) + [l.nd, l.np_])  # fmt: skip
                off += l.nd * l.np_ * nk
            recv_seg[peer] = (start, off - start)
        as_np = lambda rows: np.asarray(rows, dtype=np.int64).reshape(-1, LINK_WORDS)  # noqa: E731
        return {
            "local": as_np(local), "pack": as_np(pack), "unpack": as_np(unpack),
            "send_seg": send_seg, "recv_seg": recv_seg, "send_total": send_total, "recv_total": off,
        }  # fmt: skip


class HaloUpdater:
    """start()/wait() halo update of one GPU's batch field over torch.distributed.

    ``group_ranks[g]`` is the torch.distributed rank that owns GPU ``g`` (default: identity).
    ``mover`` runs a link table (default: the CUDA ``halo_move`` kernel; the gloo tests inject a CPU one).
    """

    def __init__(self, part: CubedSpherePartitioner, n_gpus: int, gpu: int, mover: Optional[Mover] = None,
                 process_group=None, group_ranks: Optional[Sequence[int]] = None, use_comm_stream: bool = True):
        self.plan = HaloPlan(part, n_gpus, gpu)
        self.part, self.n_gpus, self.gpu = part, n_gpus, gpu
        self.mover = mover or _cuda_mover
        self.pg = process_group
        self.group_ranks = list(group_ranks) if group_ranks is not None else list(range(n_gpus))
        self._cache: Dict[tuple, dict] = {}
        self._use_comm_stream = use_comm_stream
        self._comm_stream = None
        self._pending = None
        self.bytes_sent_per_update = 0

    def _prepared(self, field: torch.Tensor) -> dict:
        geo = FieldGeometry(field, self.part.halo)
        key = geo.key + (field.dtype, field.device)
        if key not in self._cache:
            t = self.plan.tables(geo)
            dev = field.device
            prep = {k: torch.from_numpy(t[k]).to(dev) for k in ("local", "pack", "unpack")}
            prep.update({k: t[k] for k in ("send_seg", "recv_seg", "send_total", "recv_total")})
            prep["send_buf"] = torch.empty(max(t["send_total"], 1), dtype=field.dtype, device=dev)
            prep["recv_buf"] = torch.empty(max(t["recv_total"], 1), dtype=field.dtype, device=dev)
            prep["nk"] = geo.nk
            for k in ("local", "pack", "unpack"):  # largest strip of each table, known on the host
                prep["max_" + k] = int((t[k][:, 8] * t[k][:, 9]).max()) if len(t[k]) else 0
            self._cache[key] = prep
        return self._cache[key]

    def _movers(self, prep: dict, field: torch.Tensor, flat: torch.Tensor):
        """(pack, local, unpack) callables for this field.  With the CUDA mover they are PreparedCalls
        bound to the field's storage (argument marshalling done once, ~1 us of host time per launch)."""
        if self.mover is not _cuda_mover:
            return (
                lambda: self.mover(prep["pack"], prep["nk"], flat, prep["send_buf"], prep["max_pack"]),
                lambda: self.mover(prep["local"], prep["nk"], flat, flat, prep["max_local"]),
                lambda: self.mover(prep["unpack"], prep["nk"], prep["recv_buf"], flat, prep["max_unpack"]),
            )
        key = (field.data_ptr(), field.dtype)
        calls = prep.setdefault("calls", {})
        if key not in calls:
            from .. import stencils

            noop = lambda: None  # noqa: E731
            mk = lambda tbl, src, dst, mx: (  # noqa: E731
                stencils.prepare_halo_move(prep[tbl], prep["nk"], src, dst, prep[mx]) if len(prep[tbl]) else noop
            )
            calls[key] = (
                mk("pack", flat, prep["send_buf"], "max_pack"),
                mk("local", flat, flat, "max_local"),
                mk("unpack", prep["recv_buf"], flat, "max_unpack"),
            )
        return calls[key]

    def _flat(self, field: torch.Tensor) -> torch.Tensor:
        """1-D alias of the storage behind ``field`` starting at its first element (what link offsets index)."""
        extent = 1 + sum((n - 1) * s for n, s in zip(field.shape, field.stride()))
        return field.as_strided((extent,), (1,))

    def start(self, field: torch.Tensor) -> None:
        """Fill same-GPU halos, pack and post the cross-GPU exchange (asynchronous)."""
        import torch.distributed as dist

        if self._pending is not None:
            raise RuntimeError("HaloUpdater.start() called twice without wait()")
        prep = self._prepared(field)
        flat = self._flat(field)
        cuda = field.is_cuda
        ctx = None
        if cuda and self._use_comm_stream and self.plan.peers:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=field.device)
            self._comm_stream.wait_stream(torch.cuda.current_stream(field.device))
            ctx = torch.cuda.stream(self._comm_stream)
            ctx.__enter__()
        try:
            reqs = []
            pack, local, unpack = self._movers(prep, field, flat)
            if self.plan.peers:
                pack()
                ops = []
                for peer in self.plan.peers:
                    so, sn = prep["send_seg"][peer]
                    ro, rn = prep["recv_seg"][peer]
                    dst_rank = self.group_ranks[peer]
                    if rn:
                        ops.append(dist.P2POp(dist.irecv, prep["recv_buf"][ro : ro + rn], dst_rank, self.pg))
                    if sn:
                        ops.append(dist.P2POp(dist.isend, prep["send_buf"][so : so + sn], dst_rank, self.pg))
                reqs = dist.batch_isend_irecv(ops) if ops else []
                self.bytes_sent_per_update = prep["send_total"] * field.element_size()
            # same-GPU neighbours: straight copies, concurrent with the transfer
            local()
            self._pending = (prep, flat, reqs, field, unpack)
        finally:
            if ctx is not None:
                ctx.__exit__(None, None, None)

    def wait(self) -> None:
        """Complete the exchange: unpack received strips; the current stream then sees full halos."""
        if self._pending is None:
            return
        prep, flat, reqs, field, unpack = self._pending
        self._pending = None
        cuda = field.is_cuda
        on_comm = cuda and self._comm_stream is not None and self.plan.peers
        if on_comm:
            with torch.cuda.stream(self._comm_stream):
                for r in reqs:
                    r.wait()
                unpack()
            torch.cuda.current_stream(field.device).wait_stream(self._comm_stream)
        else:
            for r in reqs:
                r.wait()
            if self.plan.peers:
                unpack()

    def update(self, field: torch.Tensor) -> None:
        self.start(field)
        self.wait()


def exchange_in_process(part: CubedSpherePartitioner, n_gpus: int, fields: Sequence[torch.Tensor],
                        mover: Optional[Mover] = None) -> None:
    """Run the full multi-GPU exchange with every "GPU" held by this process (virtual ranks).

    ``fields[g]`` is GPU g's batch field.  Pack tables, message layout and unpack tables are exactly
    the ones the distributed path uses; only the transport is a tensor copy.  This is how the
    adjacency tests cover 2/4/8-GPU decompositions on one device.
    """
    mover = mover or _cuda_mover
    ups = [HaloUpdater(part, n_gpus, g, mover) for g in range(n_gpus)]
    preps = [u._prepared(f) for u, f in zip(ups, fields)]
    flats = [u._flat(f) for u, f in zip(ups, fields)]
    for g in range(n_gpus):
        if ups[g].plan.peers:
            mover(preps[g]["pack"], preps[g]["nk"], flats[g], preps[g]["send_buf"], preps[g]["max_pack"])
    for g in range(n_gpus):
        for peer in ups[g].plan.peers:
            so, sn = preps[g]["send_seg"][peer]
            ro, rn = preps[peer]["recv_seg"][g]
            assert sn == rn, "send and receive segments of a GPU pair must agree"
            preps[peer]["recv_buf"][ro : ro + rn].copy_(preps[g]["send_buf"][so : so + sn])
    for g in range(n_gpus):
        mover(preps[g]["local"], preps[g]["nk"], flats[g], flats[g], preps[g]["max_local"])
        if ups[g].plan.peers:
            mover(preps[g]["unpack"], preps[g]["nk"], preps[g]["recv_buf"], flats[g], preps[g]["max_unpack"])
