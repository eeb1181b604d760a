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

from . import protocol
from .protocol import DumpFormat, Order
from .sampler import NVMLProvider, Sampler


async def _sample_loop(sampler: Sampler, dt: float):
    while True:
        sampler.sample_once()
        await asyncio.sleep(dt)


def dump(sampler: Sampler, name: str, fmt=None) -> str:
    d = sampler.dump_dict()
    fmt = DumpFormat(fmt) if fmt is not None else DumpFormat.from_env()
    if fmt is DumpFormat.NPZ:
        path = f"{name}.npz"
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})
    elif fmt is DumpFormat.JSON:
        path = f"{name}.json"
        with open(path, "w") as f:
            json.dump(d, f, indent=4)
    return path


async def main(provider=None, socket_filename: str = protocol.SOCKET_PATH):
    print("NVML server up & waiting for connection")
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    os.makedirs(os.path.dirname(socket_filename) or ".", exist_ok=True)
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
        order = protocol.decode(await loop.sock_recv(client, 255))
        action = order.get("action")
        if action == Order.STOP.value:
            print("[NVML SERVER] Closing...")
            client.close()
            break
        elif action == Order.START.value:
            sampler.dt = float(order["dt"])
            if task is None:
                task = loop.create_task(_sample_loop(sampler, sampler.dt))
            print(f"[NVML SERVER] Recording every {sampler.dt} seconds")
        elif action == Order.DUMP.value:
            print(f"[NVML SERVER] Dumped {dump(sampler, order['dump_name'])}")
        elif action == Order.TICK.value:
            print(f"[NVML SERVER] Recorded tick at {sampler.tick()}")
        else:
            print(f"[NVML SERVER] Received unknown {order}")
        client.close()
    if task is not None:
        task.cancel()
    server.close()
    if os.path.exists(socket_filename):
        os.remove(socket_filename)


def cli(provider=None, socket_filename: str = protocol.SOCKET_PATH):
    asyncio.run(main(provider, socket_filename))
