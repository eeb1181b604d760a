"""hws server: asyncio loop on an AF_UNIX socket taking one JSON order per connection.

Reference: /root/reference/src/tcn/hws/server.py:64-151 (same socket, same orders, same npz keys;
per-GPU columns instead of the single-GPU assert at :90, TICK fixed, dt/timestamps in the dump).
"""
from __future__ import annotations

import asyncio
import json
import os
import socket

import numpy as np

from .constants import (HWS_DUMP_FORMAT, HWS_DUMP_JSON, HWS_DUMP_NPZ, SERV_ORDER_DUMP, SERV_ORDER_START,
                        SERV_ORDER_STOP, SERV_ORDER_TICK, SOCKET_DIRECTORY, SOCKET_FILENAME)
from .sampler import NVMLProvider, Sampler


async def _sample_loop(sampler: Sampler, dt: float):
    while True:
        sampler.sample_once()
        await asyncio.sleep(dt)


def dump(sampler: Sampler, name: str, fmt: str = HWS_DUMP_FORMAT) -> str:
    d = sampler.dump_dict()
    if fmt == HWS_DUMP_NPZ:
        path = f"{name}.npz"
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})
    elif fmt == HWS_DUMP_JSON:
        path = f"{name}.json"
        with open(path, "w") as f:
            json.dump(d, f, indent=4)
    else:
        raise RuntimeWarning(f"Can't dump in unknown format {fmt}")
    return path


async def main(provider=None, socket_filename: str = SOCKET_FILENAME):
    print("NVML server up & waiting for connection")
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    os.makedirs(os.path.dirname(socket_filename) or SOCKET_DIRECTORY, exist_ok=True)
    if os.path.exists(socket_filename):
        os.remove(socket_filename)
    server.bind(socket_filename)
    server.listen(1)
    server.setblocking(False)

    sampler = Sampler(provider if provider is not None else NVMLProvider())
    print("[NVML SERVER] Driver Version:", sampler.provider.driver())
    for i, n in zip(sampler.provider.indices, sampler.provider.names()):
        print(f"[NVML SERVER] Device {i}: {n}")

    loop = asyncio.get_running_loop()
    task = None
    while True:
        client, _ = await loop.sock_accept(server)
        request = (await loop.sock_recv(client, 255)).decode("utf8")
        order = json.loads(request)
        action = order.get("action")
        if action == SERV_ORDER_STOP:
            print("[NVML SERVER] Closing...")
            client.close()
            break
        elif action == SERV_ORDER_START:
            sampler.dt = float(order["dt"])
            if task is None:
                task = loop.create_task(_sample_loop(sampler, sampler.dt))
            print(f"[NVML SERVER] Recording every {sampler.dt} seconds")
        elif action == SERV_ORDER_DUMP:
            print(f"[NVML SERVER] Dumped {dump(sampler, order['dump_name'])}")
        elif action == SERV_ORDER_TICK:
            print(f"[NVML SERVER] Recorded tick at {sampler.tick()}")
        else:
            print(f"[NVML SERVER] Received unknown {order}")
        client.close()
    if task is not None:
        task.cancel()
    server.close()
    if os.path.exists(socket_filename):
        os.remove(socket_filename)


def cli(provider=None, socket_filename: str = SOCKET_FILENAME):
    asyncio.run(main(provider, socket_filename))
