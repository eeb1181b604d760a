"""hws server/client round trip with a fake NVML provider (runs on the GPU-less container).
Behavioural spec: /root/reference/test/hws/test_gpu_psu.py:19-32 (start, work, dump, stop, dump file exists)."""
import multiprocessing as mp
import os
import time

import numpy as np

from b200stencil.hws import FakeNVML, Sampler
from b200stencil.hws import analysis
from b200stencil.hws.client import client_main
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
    client_main("start", socket_filename=sock)
    time.sleep(0.5)
    client_main("tick", socket_filename=sock)
    time.sleep(0.3)
    client_main("dump", "hws_dump", socket_filename=sock)
    time.sleep(0.3)
    client_main("stop", socket_filename=sock)
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
    rep = analysis.energy_envelop_calculation(d["cpu_psu"], d["gpu_psu"], d["timestamps"], verbose=False)
    assert rep.GPU_envelop_kWh > 0 and rep.duration_s > 0


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
    rep = analysis.energy_envelop_calculation(np.full(11, 360.0), np.full((11, 2), 720.0), t, verbose=False)
    assert abs(rep.CPU_envelop_kWh - 0.36 * 5 / 3600) < 1e-12
    assert abs(rep.GPU_envelop_kWh - 1.44 * 5 / 3600) < 1e-12
