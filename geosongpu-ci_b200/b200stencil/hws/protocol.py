"""Wire protocol of the hardware sampler: one JSON object per UNIX-socket connection.

Kept byte-compatible with the reference daemon so its launcher scripts keep working
(/root/reference/src/tcn/hws/constants.py:5-46, client.py:7-13): socket ``./sockets-runtime/hws``; actions
``START`` (with ``dt`` seconds), ``STOP``, ``DUMP`` (with ``dump_name``), ``TICK``; client verbs ``start``,
``stop``, ``dump``, ``tick``; dump format from ``HWSAMPLER_DUMP_FORMAT`` (``npz`` | ``json``).
"""
from __future__ import annotations

import enum
import json
import os
from typing import Any, Dict

SOCKET_PATH = os.path.join(".", "sockets-runtime", "hws")
DEFAULT_DT_S = 0.1
DEFAULT_DUMP_NAME = "hws_dump"


class DumpFormat(str, enum.Enum):
    NPZ = "npz"
    JSON = "json"

    @classmethod
    def from_env(cls) -> "DumpFormat":
        return cls(os.getenv("HWSAMPLER_DUMP_FORMAT", cls.NPZ.value))


class Order(str, enum.Enum):
    """Server-side action names; the lower-case name is the client verb."""

    START = "START"
    STOP = "STOP"
    DUMP = "DUMP"
    TICK = "TICK"

    @property
    def verb(self) -> str:
        return self.name.lower()

    @classmethod
    def from_verb(cls, verb: str) -> "Order":
        try:
            return cls[verb.upper()]
        except KeyError:
            raise RuntimeError(f"[HWS Client] Unknown cmds {verb} as first argument of the executable") from None


def encode(order: Order, dump_name: str = DEFAULT_DUMP_NAME, dt: float = DEFAULT_DT_S) -> bytes:
    """The message the reference client sends for ``order`` (every message carries ``dump_name``)."""
    msg: Dict[str, Any] = {"action": order.value}
    if order is Order.START:
        msg["dt"] = dt
    msg["dump_name"] = dump_name
    return json.dumps(msg).encode("utf8")


def decode(raw: bytes) -> Dict[str, Any]:
    return json.loads(raw.decode("utf8"))


# series recorded per sample: gpu_* are [sample][gpu], cpu_* are [sample]
# (the first four gpu_* and both cpu_* are the reference's npz keys, server.py:77-83)
SERIES_GPU = ("gpu_psu", "gpu_exe_utl", "gpu_mem_utl", "gpu_mem", "gpu_sm_mhz", "gpu_throttle")
SERIES_CPU = ("cpu_exe_utl", "cpu_psu")
