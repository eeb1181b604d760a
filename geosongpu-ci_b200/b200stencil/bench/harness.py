"""CUDA-event timing harness and roofline arithmetic.

Successor of the reference's profiling helper
(/root/reference/src/tcn/py_ftn_interface/templates/cuda_profiler.py:22-75): same context-manager
surface (``CUDAProfiler(label)``, ``TimedCUDAProfiler(label, timings)``, NVTX ranges), but timed
with CUDA events on the launching stream instead of ``perf_counter`` around a device sync, which
over-counts at the microsecond scale of these kernels.
"""
from __future__ import annotations

import json
import os
import statistics
from typing import Callable, Dict, List, Optional

import torch

NOMINAL_HBM_GBS = 8000.0  # BASELINE.json north_star "~8 TB/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peaks(repo_root: Optional[str] = None) -> Dict[str, object]:
    """HBM roofline denominator: MEASURED_PEAKS.json when present, else the recipe's fallback."""
    root = repo_root or os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
    path = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": FALLBACK_HBM_GBS, "source": "fallback (B200_PROFILING.md)"}


class CUDAProfiler:
    """NVTX range + device sync on entry/exit (reference: cuda_profiler.py:22-36)."""

    def __init__(self, label: str) -> None:
        self.label = label

    def __enter__(self):
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push(self.label)
        return self

    def __exit__(self, _type, _val, _traceback):
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()

    @classmethod
    def sync_device(cls):
        torch.cuda.synchronize()

    @classmethod
    def start_cuda_profiler(cls):
        torch.cuda.cudart().cudaProfilerStart()

    @classmethod
    def stop_cuda_profiler(cls):
        torch.cuda.cudart().cudaProfilerStop()

    @classmethod
    def mark_cuda_profiler(cls, message: str):
        torch.cuda.nvtx.mark(message)


class TimedCUDAProfiler(CUDAProfiler):
    """Appends the elapsed seconds of the block to ``timings[label]`` (reference: cuda_profiler.py:59-75),
    measured with CUDA events on the current stream."""

    def __init__(self, label: str, timings: Dict[str, List[float]]) -> None:
        super().__init__(label)
        self._timings = timings
        self._t0 = torch.cuda.Event(enable_timing=True)
        self._t1 = torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        super().__enter__()
        self._t0.record()
        return self

    def __exit__(self, _type, _val, _traceback):
        self._t1.record()
        super().__exit__(_type, _val, _traceback)
        self._timings.setdefault(self.label, []).append(self._t0.elapsed_time(self._t1) * 1e-3)


class L2Flusher:
    """Writes a buffer larger than the 126 MB L2 between timed iterations."""

    def __init__(self, nbytes: int = 256 << 20, device="cuda"):
        self.buf = torch.empty(nbytes // 4, dtype=torch.int32, device=device)

    def __call__(self):
        self.buf.zero_()


def time_kernel(fn: Callable[[int], None], iters: int = 20, warmup: int = 3, flush: Optional[L2Flusher] = None,
                rotate: int = 1) -> Dict[str, float]:
    """Per-launch device time of ``fn(slot)`` in milliseconds: CUDA events around every launch.

    ``rotate`` distinct buffer sets are cycled (``slot = it % rotate``) so that the working set
    exceeds L2; alternatively ``flush`` is run (untimed) before every launch.
    """
    for it in range(warmup):
        fn(it % rotate)
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for it in range(iters):
        if flush is not None:
            flush()
        starts[it].record()
        fn(it % rotate)
        stops[it].record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    return {"min_ms": min(ms), "median_ms": statistics.median(ms), "mean_ms": sum(ms) / len(ms), "iters": iters}


def time_graph(fn: Callable[[int], None], rotate: int = 1, launches_per_graph: int = 20, iters: int = 10,
               warmup: int = 3) -> Dict[str, float]:
    """Per-launch device time with launch overhead removed: ``launches_per_graph`` calls of
    ``fn(slot)`` (slots rotating, so the working set exceeds L2) are captured into one CUDA graph and
    the replays are timed with CUDA events.  For kernels of a few microseconds, where the host
    cannot issue launches as fast as the device retires them."""
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for it in range(max(warmup, rotate)):
            fn(it % rotate)
    stream.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        for it in range(launches_per_graph):
            fn(it % rotate)
    for _ in range(warmup):
        graph.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b) / launches_per_graph)
    return {"min_ms": min(ms), "median_ms": statistics.median(ms), "mean_ms": sum(ms) / len(ms), "iters": iters,
            "launches_per_graph": launches_per_graph}


def roofline(bytes_per_launch: float, ms: float, peak_gbs: float) -> Dict[str, float]:
    achieved = bytes_per_launch / (ms * 1e-3) / 1e9
    return {"achieved": achieved, "peak": peak_gbs, "frac": achieved / peak_gbs, "frac_of_nominal_8TBs": achieved / NOMINAL_HBM_GBS}
