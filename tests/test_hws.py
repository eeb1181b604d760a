"""hws server/client round trip with a fake NVML provider (runs on the GPU-less container).
Behavioural spec: /root/reference/test/hws/test_gpu_psu.py:19-32 (start, work, dump, stop, dump file exists)."""
import multiprocessing as mp
import os
import time

import numpy as np

from b200stencil.hws import FakeNVML, Sampler
from b200stencil.hws import analysis
from b200stencil.hws.client import send_order
from b200stencil.hws.sampler import decode_reasons


def _serve(sock):
    from b200stencil.hws import server

    server.cli(FakeNVML(2), sock)


def test_server_client_roundtrip(tmp_path):
    os.chdir(tmp_path)
    sock = str(tmp_path / "hws.sock")
    p = mp.get_context("spawn").Process(target=_serve, args=(sock,))
    p.start()
    for _ in range(100):
        if os.path.exists(sock):
            break
        time.sleep(0.1)
    send_order("start", socket_path=sock)
    time.sleep(0.5)
    send_order("tick", socket_path=sock)
    time.sleep(0.3)
    send_order("dump", "hws_dump", socket_path=sock)
    time.sleep(0.3)
    send_order("stop", socket_path=sock)
    p.join(timeout=10)
    assert p.exitcode == 0
    path = tmp_path / "hws_dump.npz"
    assert path.exists()
    d = np.load(path)
    # reference keys (server.py:77-83), one column per GPU
    for k in ("gpu_psu", "gpu_exe_utl", "gpu_mem_utl", "gpu_mem"):
        assert d[k].ndim == 2 and d[k].shape[1] == 2 and d[k].shape[0] >= 3
    assert d["cpu_exe_utl"].shape[0] == d["gpu_psu"].shape[0] == d["timestamps"].shape[0]
    assert 0 < int(d["ticks"][0]) <= d["gpu_psu"].shape[0]  # TICK = sample index, not len(dict)
    assert float(d["dt"]) == 0.1
    rep = analysis.energy_report(d["cpu_psu"], d["gpu_psu"], d["timestamps"])
    assert rep.gpu_wh > 0 and rep.duration_s > 0 and "Wh" in rep.describe()


def test_protocol_matches_the_reference_wire_format():
    """START carries dt and dump_name, the others dump_name only (reference client.py:7-13, constants.py:33-46)."""
    import json

    from b200stencil.hws import protocol

    assert json.loads(protocol.encode(protocol.Order.START)) == {"action": "START", "dt": 0.1, "dump_name": "hws_dump"}
    assert json.loads(protocol.encode(protocol.Order.from_verb("dump"), "x")) == {"action": "DUMP", "dump_name": "x"}
    assert protocol.SOCKET_PATH == "./sockets-runtime/hws"
    import pytest

    with pytest.raises(RuntimeError):
        protocol.Order.from_verb("reboot")


def test_sampler_thread_and_clock_summary():
    s = Sampler(FakeNVML(1), dt=0.01)
    with s:
        time.sleep(0.15)
    assert len(s.timestamps) >= 3
    c = s.clocks_summary(0)
    assert c["sm_mhz"] == 1900.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["rejected"] is False
    assert decode_reasons(0x8 | 0x40) == ["hw_slowdown", "hw_thermal_slowdown"]


def test_energy_exact_for_constant_power():
    t = np.arange(11) * 0.5  # 5 s
    rep = analysis.energy_report(np.full(11, 360.0), np.full((11, 2), 720.0), t)
    assert abs(rep.cpu_wh - 360.0 * 5 / 3600) < 1e-9
    assert abs(rep.gpu_wh - 1440.0 * 5 / 3600) < 1e-9
    assert abs(rep.total_kwh - (360.0 + 1440.0) * 5 / 3600 / 1000) < 1e-12
