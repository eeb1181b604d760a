"""Energy envelope of an hws dump (reference: /root/reference/src/tcn/hws/analysis.py:20-72).

The reference integrates power over SAMPLE INDEX and then guesses the time base from the default
sample rate (flagged "Wrong?!" by its author, :38-50).  Dumps written here carry timestamps, so the
integral is taken over time and kWh is exact; the sample-index figures are kept for comparison.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, Optional

import numpy as np


@dataclasses.dataclass
class EnergyReport:
    CPU_envelop_integrated: float = 0  # kW * sample_count   (reference unit)
    CPU_envelop_kWh: float = 0
    GPU_envelop_integrated: float = 0
    GPU_envelop_kWh: float = 0
    overall_envelop_integrated: float = 0
    overall_envelop_kWh: float = 0
    duration_s: float = 0
    samples: int = 0


def _trapz(y, x=None):
    fn = getattr(np, "trapezoid", None) or np.trapz
    return float(fn(y, x)) if x is not None else float(fn(y))


def energy_envelop_calculation(cpu_psu_data, gpu_psu_data, timestamps: Optional[np.ndarray] = None,
                               dt: Optional[float] = None, verbose: bool = True) -> EnergyReport:
    """cpu_psu_data [samples] in W, gpu_psu_data [samples] or [samples, gpus] in W."""
    cpu = np.asarray(cpu_psu_data, dtype=float)
    gpu = np.asarray(gpu_psu_data, dtype=float)
    if gpu.ndim == 2:
        gpu = gpu.sum(axis=1)  # all GPUs of the node
    n = len(cpu)
    if timestamps is not None and len(timestamps) == n:
        t = np.asarray(timestamps, dtype=float) - float(timestamps[0])
    else:
        t = np.arange(n) * (dt if dt is not None else 0.1)
    r = EnergyReport(samples=n, duration_s=float(t[-1]) if n else 0.0)
    r.GPU_envelop_integrated = _trapz(gpu / 1000)
    r.CPU_envelop_integrated = _trapz(cpu / 1000)
    r.overall_envelop_integrated = r.GPU_envelop_integrated + r.CPU_envelop_integrated
    r.GPU_envelop_kWh = _trapz(gpu / 1000, t) / 3600
    r.CPU_envelop_kWh = _trapz(cpu / 1000, t) / 3600
    r.overall_envelop_kWh = r.GPU_envelop_kWh + r.CPU_envelop_kWh
    if verbose:
        print(
            f"Number of samples: {n} over {r.duration_s:.2f} s\n"
            f"CPU envelop: {r.CPU_envelop_kWh * 1000:.4f} Wh\n"
            f"GPU envelop: {r.GPU_envelop_kWh * 1000:.4f} Wh\n"
            f"Overall envelop: {r.overall_envelop_kWh * 1000:.4f} Wh"
        )
    return r


def load_data(data_filepath: str, data_format: str = "npz") -> Dict[str, Any]:
    if data_format != "npz":
        raise NotImplementedError(f"Format {data_format} not implemented")
    return np.load(data_filepath, allow_pickle=False)
