"""Power / memory figures used for the CPU power model and for labelling dumps.

The reference hard-codes an A100-40GB and two EPYC parts (/root/reference/src/tcn/hws/constants.py:48-63) and
selects them with ``HWS_HW_CPU`` / ``HWS_HW_GPU``; the same variables work here, with B200 and a generic Xeon
added as defaults.  GPU memory size is read from NVML at run time; ``max_vram_mib`` is only a label.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional


@dataclass(frozen=True)
class HardwareSpec:
    tdp_w: float
    idle_w: Optional[float] = None
    max_vram_mib: Optional[int] = None


SPECS: Dict[str, HardwareSpec] = {
    "EPYC 7402": HardwareSpec(tdp_w=180, idle_w=60),
    "EPYC 7763": HardwareSpec(tdp_w=280, idle_w=60),
    "Xeon (generic)": HardwareSpec(tdp_w=350, idle_w=80),
    "A100_SX40": HardwareSpec(tdp_w=400, max_vram_mib=40536),
    # 1000 W limit, 183,359 MiB reported by the driver (B200_PROFILING.md)
    "B200_SXM": HardwareSpec(tdp_w=1000, max_vram_mib=183359),
}

CPU_LABEL = os.getenv("HWS_HW_CPU", "Xeon (generic)")
GPU_LABEL = os.getenv("HWS_HW_GPU", "B200_SXM")


def cpu_power_w(util_pct: float, label: str = CPU_LABEL) -> float:
    """Linear utilisation -> watts model of the reference sampler (server.py:55-58), floored at idle."""
    spec = SPECS[label]
    return max(util_pct / 100.0 * spec.tdp_w, spec.idle_w or 0.0)
