"""Energy of a sampled run from an hws dump.

The reference integrates power over the SAMPLE INDEX and then guesses a time base from the default sample
rate (its author flags the kWh figure "Wrong?!", /root/reference/src/tcn/hws/analysis.py:27,38-50).  Dumps
written by this package carry timestamps and ``dt``, so the integral is taken over time and the energy is
exact; the sample-index integral is still reported (``*_kw_samples``) for comparison with old dumps.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import numpy as np

_trapezoid = getattr(np, "trapezoid", None) or np.trapz


@dataclass
class EnergyReport:
    samples: int
    duration_s: float
    cpu_wh: float
    gpu_wh: float
    cpu_kw_samples: float  # reference unit: kW x sample count
    gpu_kw_samples: float

    @property
    def total_wh(self) -> float:
        return self.cpu_wh + self.gpu_wh

    @property
    def total_kwh(self) -> float:
        return self.total_wh / 1000.0

    def describe(self) -> str:
        lines = [
            f"{self.samples} samples over {self.duration_s:.2f} s",
            f"  CPU  {self.cpu_wh:10.4f} Wh   ({self.cpu_kw_samples:.2f} kW.samples)",
            f"  GPU  {self.gpu_wh:10.4f} Wh   ({self.gpu_kw_samples:.2f} kW.samples)",
            f"  all  {self.total_wh:10.4f} Wh",
        ]
        return "\n".join(lines)


def energy_report(cpu_w, gpu_w, timestamps: Optional[np.ndarray] = None, dt: Optional[float] = None) -> EnergyReport:
    """cpu_w [samples] and gpu_w [samples] or [samples, gpus] in watts -> energy in Wh."""
    cpu = np.asarray(cpu_w, dtype=float)
    gpu = np.asarray(gpu_w, dtype=float)
    if gpu.ndim == 2:
        gpu = gpu.sum(axis=1)  # every GPU of the node
    n = int(cpu.shape[0])
    if timestamps is not None and len(timestamps) == n:
        t = np.asarray(timestamps, dtype=float)
        t = t - t[0] if n else t
    else:
        t = np.arange(n) * (0.1 if dt is None else dt)
    joule = lambda w: float(_trapezoid(w, t)) if n > 1 else 0.0  # noqa: E731
    return EnergyReport(
        samples=n,
        duration_s=float(t[-1]) if n else 0.0,
        cpu_wh=joule(cpu) / 3600.0,
        gpu_wh=joule(gpu) / 3600.0,
        cpu_kw_samples=float(_trapezoid(cpu / 1000.0)) if n > 1 else 0.0,
        gpu_kw_samples=float(_trapezoid(gpu / 1000.0)) if n > 1 else 0.0,
    )


def load(path: str) -> Dict[str, Any]:
    if not path.endswith(".npz"):
        raise NotImplementedError("only npz dumps can be analysed")
    return dict(np.load(path, allow_pickle=False))
