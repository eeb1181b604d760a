"""Protocol constants and hardware specs (reference: /root/reference/src/tcn/hws/constants.py:1-63).

Same socket path, same JSON orders (START{dt} / STOP / DUMP{dump_name} / TICK), same client command
names and environment variables.  Added: a B200 spec entry (the reference knows A100 only, :48-58)
and per-GPU series in the dump.
"""
import os
from typing import Any, Dict

SOCKET_DIRECTORY = "./sockets-runtime"
SOCKET_FILENAME = f"{SOCKET_DIRECTORY}/hws"

HWS_DUMP_NAME = "hws_dump"
HWS_DUMP_NPZ = "npz"
HWS_DUMP_JSON = "json"
HWS_DUMP_FORMAT = os.getenv("HWSAMPLER_DUMP_FORMAT", HWS_DUMP_NPZ)

SERV_ORDER_START = "START"
SERV_ORDER_STOP = "STOP"
SERV_ORDER_DUMP = "DUMP"
SERV_ORDER_TICK = "TICK"

CLIENT_CMD_START = "start"
CLIENT_CMD_STOP = "stop"
CLIENT_CMD_DUMP = "dump"
CLIENT_CMD_TICK = "tick"

DEFAULT_SAMPLERATE_IN_S = 0.1

CLIENT_CMDS: Dict[str, Any] = {
    CLIENT_CMD_START: {"action": SERV_ORDER_START, "dt": DEFAULT_SAMPLERATE_IN_S},
    CLIENT_CMD_STOP: {"action": SERV_ORDER_STOP},
    CLIENT_CMD_DUMP: {"action": SERV_ORDER_DUMP, "dump_name": HWS_DUMP_NAME},
    CLIENT_CMD_TICK: {"action": SERV_ORDER_TICK},
}

LBL_EPYC_7402 = "EPYC 7402"
LBL_EPYC_7763 = "EPYC 7763"
LBL_XEON_GENERIC = "Xeon (generic)"
LBL_A100 = "A100_SX40"
LBL_B200 = "B200_SXM"
HWS_HARDWARE_SPECS = {
    LBL_EPYC_7402: {"PSU_IDLE": 60, "PSU_TDP": 180},
    LBL_EPYC_7763: {"PSU_IDLE": 60, "PSU_TDP": 280},
    LBL_XEON_GENERIC: {"PSU_IDLE": 80, "PSU_TDP": 350},
    LBL_A100: {"PSU_TDP": 400, "MAX_VRAM": 40536},
    # 1000 W limit, 183,359 MiB reported by the driver (B200_PROFILING.md); the server prefers
    # nvmlDeviceGetMemoryInfo().total over this constant
    LBL_B200: {"PSU_TDP": 1000, "MAX_VRAM": 183359},
}

HWS_HW_CPU = os.getenv("HWS_HW_CPU", LBL_XEON_GENERIC)
HWS_HW_GPU = os.getenv("HWS_HW_GPU", LBL_B200)

# series recorded per sample; gpu_* are [sample][gpu], cpu_* are [sample] (reference keys: server.py:77-83)
SERIES_GPU = ("gpu_psu", "gpu_exe_utl", "gpu_mem_utl", "gpu_mem", "gpu_sm_mhz", "gpu_throttle")
SERIES_CPU = ("cpu_exe_utl", "cpu_psu")
