"""Host <-> device staging at the boundary: the part of the reference's bridge that owns layout and
ownership (/root/reference/src/tcn/py_ftn_interface/templates/data_conversion.py:26-191).

Reference contract kept:
  * the caller (Fortran / NumPy) owns host buffers for the whole run; this side borrows them
    (``np.frombuffer`` over ``ffi.buffer``, data_conversion.py:106-111);
  * Fortran column-major ``(i, j, k)`` memory is an i-fastest ``[i, j, k]`` field without a copy
    (data_conversion.py:134-148);
  * NO type casting on the way in (data_conversion.py:30); uploads alternate between two
    non-blocking streams and the caller must ``sync()`` (data_conversion.py:42-44, 54-57, 127-131).
Reference defects not reproduced: the copy-back used ``4 * size`` bytes for every dtype
(data_conversion.py:95) -- here the byte count follows the dtype; ``sync()`` dereferenced streams
that only exist on the GPU path -- here there is only the GPU path.

Also here: ``fv_tp2d_host``, the end-to-end form of the transport stencil for callers whose fields
live in (pinned) host memory: per-sub-domain chunks are uploaded, computed and downloaded on three
streams so the PCIe transfers of neighbouring chunks overlap the kernel.
"""
from __future__ import annotations

import os
from math import prod
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import fields, stencils

_TYPEMAP = {"float": np.float32, "double": np.float64, "int": np.int32, "int64_t": np.int64}


class HostDeviceConversion:
    """Fortran/C pointer <-> i-fastest device field (successor of ``FortranPythonConversion``)."""

    def __init__(self, device: str = "cuda"):
        import cffi

        if not torch.cuda.is_available():
            raise RuntimeError("HostDeviceConversion needs a CUDA device: b200stencil has no CPU path")
        self._ffi = cffi.FFI()
        self._device = torch.device(device)
        self._stream_A = torch.cuda.Stream(device=self._device)
        self._stream_B = torch.cuda.Stream(device=self._device)
        self._current = self._stream_A

    def sync(self) -> None:
        """Synchronize the two working streams (reference: data_conversion.py:54-57)."""
        self._stream_A.synchronize()
        self._stream_B.synchronize()

    def _swap(self) -> torch.cuda.Stream:
        s = self._current
        self._current = self._stream_B if s is self._stream_A else self._stream_A
        return s

    def _borrow(self, fptr, dim: Sequence[int]) -> np.ndarray:
        ftype = self._ffi.getctype(self._ffi.typeof(fptr).item)
        if ftype not in _TYPEMAP:
            raise TypeError(f"unsupported C element type {ftype}")
        nbytes = prod(dim) * self._ffi.sizeof(ftype)
        flat = np.frombuffer(self._ffi.buffer(fptr, nbytes), dtype=_TYPEMAP[ftype])
        return flat.reshape(tuple(reversed(dim))).transpose()  # i-fastest [i,j,k] view, zero copy

    def fortran_to_device(self, fptr, dim: Sequence[int], swap_axes: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """Upload Fortran memory of shape ``dim`` (column-major) as an i-fastest device field."""
        host = self._borrow(fptr, dim)
        if swap_axes:
            host = np.swapaxes(host, swap_axes[0], swap_axes[1])
        # allocated on the CALLER's stream (that is where the field will be used and, later, freed: the caching
        # allocator may then hand the block back to that stream only), uploaded on a working stream
        consumer = torch.cuda.current_stream(self._device)
        dev = fields.empty(host.shape, torch.from_numpy(np.empty(0, host.dtype)).dtype, self._device)
        up = self._swap()
        up.wait_stream(consumer)  # the block may still be in use by earlier work of the caller
        with torch.cuda.stream(up):
            dev.copy_(torch.from_numpy(host), non_blocking=True)
        dev.record_stream(up)
        return dev

    def device_to_fortran(self, field: torch.Tensor, fptr, swap_axes: Optional[Tuple[int, int]] = None) -> None:
        """Copy a device field back into caller-owned Fortran memory (no cast: dtypes must agree)."""
        dim = list(field.shape)
        if swap_axes:
            dim[swap_axes[0]], dim[swap_axes[1]] = dim[swap_axes[1]], dim[swap_axes[0]]
        host = self._borrow(fptr, dim)
        if swap_axes:
            host = np.swapaxes(host, swap_axes[0], swap_axes[1])
        if torch.from_numpy(np.empty(0, host.dtype)).dtype != field.dtype:
            raise TypeError(f"device field is {field.dtype}, host buffer is {host.dtype}: the bridge does not cast")
        producer = torch.cuda.current_stream(field.device)  # the stream the caller produced `field` on
        down = self._swap()
        down.wait_stream(producer)
        with torch.cuda.stream(down):
            torch.from_numpy(host).copy_(field, non_blocking=False)
        field.record_stream(down)


# ---- pinned host fields ------------------------------------------------------------------------------


def pinned_like(field: torch.Tensor) -> torch.Tensor:
    """Pinned host tensor with the same shape AND the same strides as a device field, so that a
    host<->device copy is one flat memcpy."""
    extent = 1 + sum((n - 1) * s for n, s in zip(field.shape, field.stride()))
    base = torch.empty(extent, dtype=field.dtype, pin_memory=True)
    return base.as_strided(tuple(field.shape), tuple(field.stride()))


def _flat(t: torch.Tensor) -> torch.Tensor:
    extent = 1 + sum((n - 1) * s for n, s in zip(t.shape, t.stride()))
    return t.as_strided((extent,), (1,))


def copy_flat(dst: torch.Tensor, src: torch.Tensor) -> int:
    """dst <- src for two tensors of identical shape/strides as one contiguous copy; returns bytes."""
    if tuple(dst.stride()) != tuple(src.stride()) or tuple(dst.shape) != tuple(src.shape):
        dst.copy_(src, non_blocking=True)
        return dst.numel() * dst.element_size()
    d, s = _flat(dst), _flat(src)
    d.copy_(s, non_blocking=True)
    return d.numel() * d.element_size()


class FvTp2dHost:
    """fv_tp2d for HOST-resident batch fields [b, i, j, k] (pinned, i-fastest): a 3-stream pipeline
    upload(b+1) | compute(b) | download(b-1) over the sub-domains of the batch.

    The device working set is two sub-domains, whatever the batch size.
    """

    NAMES = ("q", "crx", "xfx", "cry", "yfx", "rarea")

    def __init__(self, ni: int, nj: int, nk: int, dtype=torch.float64, device="cuda"):
        self.device = torch.device(device)
        self.shapes = {
            "q": (ni + 6, nj + 6, nk), "crx": (ni + 1, nj, nk), "xfx": (ni + 1, nj, nk),
            "cry": (ni, nj + 1, nk), "yfx": (ni, nj + 1, nk), "rarea": (ni, nj), "q_out": (ni, nj, nk),
        }  # fmt: skip
        self.slots = [{n: fields.empty(s, dtype, device) for n, s in self.shapes.items()} for _ in range(2)]
        self.s_up = torch.cuda.Stream(device=self.device)
        self.s_run = torch.cuda.Stream(device=self.device)
        self.s_down = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def host_fields(self, nb: int) -> Dict[str, torch.Tensor]:
        """Pinned host batch fields laid out like the device slots (one flat memcpy per sub-domain)."""
        out = {}
        for n, s in self.shapes.items():
            proto = self.slots[0][n]
            per = 1 + sum((m - 1) * st for m, st in zip(proto.shape, proto.stride()))
            base = torch.empty(nb * per, dtype=proto.dtype, pin_memory=True)
            out[n] = base.as_strided((nb,) + tuple(proto.shape), (per,) + tuple(proto.stride()))
        return out

    def __call__(self, host: Dict[str, torch.Tensor], q_dev: Optional[torch.Tensor] = None, exchange=None, sync=None) -> None:
        """host['q_out'][b] <- fv_tp2d(host inputs [b]) for every b; returns when the data is on the host.

        With ``q_dev`` (a resident halo-padded batch field, e.g. ``HaloContext.field``) and ``exchange`` (a callable
        that refreshes its halos on the current stream: ``HaloExchange.update`` or ``lambda: updater.update(q_dev)``)
        the step includes the halo update: the whole of q is uploaded first, its halos are exchanged with the
        neighbouring sub-domains / GPUs, and the per-sub-domain pipeline carries the other five inputs.
        ``sync`` (e.g. ``HaloContext.barrier``) is called first when peers read ``q_dev``: nobody may still be pulling
        the previous step's values while this step's upload overwrites them."""
        nb = host["q"].shape[0]
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_up, self.s_run, self.s_down):
            s.wait_stream(cur)
        up_done: List[torch.cuda.Event] = []
        run_done: List[torch.cuda.Event] = []
        down_done: List[Optional[torch.cuda.Event]] = [None, None]
        self.h2d_bytes = self.d2h_bytes = 0
        names = self.NAMES
        if q_dev is not None:
            if sync is not None:
                sync()
            names = tuple(n for n in self.NAMES if n != "q")
            with torch.cuda.stream(self.s_up):
                for b in range(nb):
                    self.h2d_bytes += copy_flat(q_dev[b], host["q"][b])
                q_up = torch.cuda.Event()
                q_up.record()
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(q_up)
                if exchange is not None:
                    exchange()
        for b in range(nb):
            slot = self.slots[b & 1]
            with torch.cuda.stream(self.s_up):
                if b >= 2:
                    self.s_up.wait_event(run_done[b - 2])  # inputs of this slot are free again
                for n in names:
                    self.h2d_bytes += copy_flat(slot[n], host[n][b])
                ev = torch.cuda.Event()
                ev.record()
                up_done.append(ev)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(up_done[b])
                if down_done[b & 1] is not None:
                    self.s_run.wait_event(down_done[b & 1])  # q_out of this slot has been downloaded
                stencils.fv_tp2d(slot["q"] if q_dev is None else q_dev[b], slot["crx"], slot["xfx"], slot["cry"], slot["yfx"],
                                 slot["rarea"], slot["q_out"])
                ev = torch.cuda.Event()
                ev.record()
                run_done.append(ev)
            with torch.cuda.stream(self.s_down):
                self.s_down.wait_event(run_done[b])
                self.d2h_bytes += copy_flat(host["q_out"][b], slot["q_out"])
                ev = torch.cuda.Event()
                ev.record()
                down_done[b & 1] = ev
        self.s_down.synchronize()
        cur.wait_stream(self.s_run)


# ---- host placement and the link the e2e numbers are bounded by ---------------------------------------


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus += list(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node the GPU hangs off (sysfs), or None when the platform does not say."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the cores of the GPU's NUMA node, so that pinned host buffers allocated afterwards are
    first-touched in the DRAM next to the GPU's PCIe root (one process per GPU: without it eight ranks can end up
    streaming through one socket's memory controllers).  Returns the node, or None if nothing was changed."""
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set(_parse_cpulist(f.read())) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def measure_pcie(device, nbytes: int = 256 << 20, reps: int = 5) -> Dict[str, float]:
    """Pinned host <-> device copy rate of this GPU's link in GB/s (best of ``reps``, CUDA events): the roofline of
    every number measured with host-resident data."""
    device = torch.device(device)
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
    out = {}
    for name, (dst, src) in {"h2d": (dev, host), "d2h": (host, dev)}.items():
        best = 0.0
        for _ in range(reps + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dst.copy_(src, non_blocking=True)
            b.record()
            b.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        out[name] = round(best, 2)
    out["bytes"] = nbytes
    return out
