"""hws client: connect, send one order, disconnect (the reference client does the same, client.py:7-13)."""
from __future__ import annotations

import socket

from . import protocol


def send_order(verb: str, dump_name: str = protocol.DEFAULT_DUMP_NAME, socket_path: str = protocol.SOCKET_PATH) -> None:
    order = protocol.Order.from_verb(verb)
    with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as conn:
        conn.connect(socket_path)
        conn.sendall(protocol.encode(order, dump_name))


def cli(command: str, dump_name: str = protocol.DEFAULT_DUMP_NAME) -> None:
    send_order(command, dump_name)
