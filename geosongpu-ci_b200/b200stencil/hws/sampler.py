"""The sampling core shared by the hws server and by in-process users (bench.py).

Reference behaviour (/root/reference/src/tcn/hws/server.py:35-61): every ``dt`` seconds read NVML
power [W], utilisation (gpu, memory) [%], memory used [MiB], psutil CPU % and a CPU power estimate
``max(util/100 * TDP, idle)``.  Differences, all from SURVEY.md section 5:
  * every visible GPU is sampled (the reference asserts ``deviceCount == 1``, server.py:90);
  * SM clock and the clock-event (throttle) reasons are recorded too -- the B200 timing rules need them;
  * wall-clock timestamps and ``dt`` are part of the record (TODO at hws/analysis.py:27);
  * TICK stores the current sample index (the reference stored ``len(dict)``, server.py:139).
NVML access goes through a small provider so the tests run without a GPU.
"""
from __future__ import annotations

import threading
import time
from typing import Dict, List, Optional, Sequence

from .protocol import SERIES_CPU, SERIES_GPU
from .specs import CPU_LABEL, cpu_power_w

# nvmlClocksEventReasons bits (nvml.h)
REASON_BITS = {
    "gpu_idle": 0x1, "applications_clocks_setting": 0x2, "sw_power_cap": 0x4, "hw_slowdown": 0x8,
    "sync_boost": 0x10, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
    "hw_power_brake_slowdown": 0x80, "display_clock_setting": 0x100,
}  # fmt: skip
REJECT_REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")


def decode_reasons(mask: int) -> List[str]:
    return [name for name, bit in REASON_BITS.items() if mask & bit]


class NVMLProvider:
    """Real NVML through pynvml."""

    def __init__(self, indices: Optional[Sequence[int]] = None):
        import pynvml

        self._n = pynvml
        pynvml.nvmlInit()
        count = pynvml.nvmlDeviceGetCount()
        self.indices = list(indices) if indices is not None else list(range(count))
        self.handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]

    def names(self) -> List[str]:
        out = []
        for h in self.handles:
            n = self._n.nvmlDeviceGetName(h)
            out.append(n.decode() if isinstance(n, bytes) else n)
        return out

    def driver(self) -> str:
        v = self._n.nvmlSystemGetDriverVersion()
        return v.decode() if isinstance(v, bytes) else v

    def memory_total_mib(self) -> List[float]:
        return [self._n.nvmlDeviceGetMemoryInfo(h).total / (1024 * 1024) for h in self.handles]

    def max_sm_mhz(self) -> List[float]:
        return [float(self._n.nvmlDeviceGetMaxClockInfo(h, self._n.NVML_CLOCK_SM)) for h in self.handles]

    def read(self) -> Dict[str, List[float]]:
        n = self._n
        out = {k: [] for k in SERIES_GPU}
        for h in self.handles:
            out["gpu_psu"].append(n.nvmlDeviceGetPowerUsage(h) / 1000)
            u = n.nvmlDeviceGetUtilizationRates(h)
            out["gpu_exe_utl"].append(float(u.gpu))
            out["gpu_mem_utl"].append(float(u.memory))
            out["gpu_mem"].append(n.nvmlDeviceGetMemoryInfo(h).used / (1024 * 1024))
            out["gpu_sm_mhz"].append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            out["gpu_throttle"].append(float(mask))
        return out


class FakeNVML:
    """Deterministic stand-in for tests on GPU-less machines."""

    def __init__(self, n_gpus: int = 2):
        self.indices = list(range(n_gpus))
        self._t = 0

    def names(self):
        return ["FAKE B200"] * len(self.indices)

    def driver(self):
        return "0.0-fake"

    def memory_total_mib(self):
        return [183359.0] * len(self.indices)

    def max_sm_mhz(self):
        return [1965.0] * len(self.indices)

    def read(self):
        self._t += 1
        g = len(self.indices)
        return {
            "gpu_psu": [100.0 + 10 * i + self._t for i in range(g)],
            "gpu_exe_utl": [50.0] * g, "gpu_mem_utl": [25.0] * g, "gpu_mem": [1024.0 * (i + 1) for i in range(g)],
            "gpu_sm_mhz": [1900.0] * g, "gpu_throttle": [4.0 if self._t % 2 else 0.0] * g,
        }  # fmt: skip


class Sampler:
    """Collects samples from a provider; usable as a background thread (``start``/``stop``) or
    driven externally (``sample_once``, used by the asyncio server)."""

    def __init__(self, provider=None, dt: float = 0.1, cpu_label: str = CPU_LABEL):
        self.provider = provider if provider is not None else NVMLProvider()
        self.dt = dt
        self.cpu_label = cpu_label
        self.data: Dict[str, list] = {k: [] for k in SERIES_GPU + SERIES_CPU}
        self.timestamps: List[float] = []
        self.ticks: List[int] = []
        self._thread: Optional[threading.Thread] = None
        self._stop = threading.Event()
        try:
            import psutil

            self._psutil = psutil
            psutil.cpu_percent()  # prime the interval counter
        except Exception:
            self._psutil = None

    # ---- one sample ------------------------------------------------------------------------------
    def sample_once(self) -> None:
        r = self.provider.read()
        for k in SERIES_GPU:
            self.data[k].append(r[k])
        cpu_use = float(self._psutil.cpu_percent()) if self._psutil else 0.0
        self.data["cpu_exe_utl"].append(cpu_use)
        self.data["cpu_psu"].append(cpu_power_w(cpu_use, self.cpu_label))
        self.timestamps.append(time.time())

    def tick(self) -> int:
        self.ticks.append(len(self.timestamps))
        return self.ticks[-1]

    # ---- thread mode -------------------------------------------------------------------------------
    def start(self) -> "Sampler":
        self._stop.clear()

        def loop():
            while not self._stop.is_set():
                try:
                    self.sample_once()
                except Exception:
                    pass
                self._stop.wait(self.dt)

        self._thread = threading.Thread(target=loop, name="hws-sampler", daemon=True)
        self._thread.start()
        return self

    def stop(self) -> "Sampler":
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=5)
            self._thread = None
        return self

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()

    # ---- output -------------------------------------------------------------------------------------
    def dump_dict(self) -> Dict[str, object]:
        out: Dict[str, object] = dict(self.data)
        out["timestamps"] = self.timestamps
        out["ticks"] = self.ticks
        out["dt"] = self.dt
        out["gpu_indices"] = list(self.provider.indices)
        out["gpu_names"] = self.provider.names()
        out["gpu_mem_total"] = self.provider.memory_total_mib()
        return out

    def clocks_summary(self, gpu: int = 0, since: Optional[float] = None, until: Optional[float] = None) -> Dict[str, object]:
        """The ``clocks`` object of the bench line: median SM clock under load, max SM clock, reasons seen."""
        import statistics

        sel = [
            i for i, t in enumerate(self.timestamps)
            if (since is None or t >= since) and (until is None or t <= until)
        ]  # fmt: skip
        mhz = [self.data["gpu_sm_mhz"][i][gpu] for i in sel]
        mask = 0
        for i in sel:
            mask |= int(self.data["gpu_throttle"][i][gpu])
        reasons = [r for r in decode_reasons(mask) if r != "gpu_idle"]
        return {
            "sm_mhz": statistics.median(mhz) if mhz else None,
            "sm_max_mhz": self.provider.max_sm_mhz()[gpu],
            "reasons": reasons,
            "samples": len(sel),
            "power_w_max": max((self.data["gpu_psu"][i][gpu] for i in sel), default=None),
            "rejected": any(r in REJECT_REASONS for r in reasons),
        }
