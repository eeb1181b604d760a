"""hws client: one JSON message per connection (reference: /root/reference/src/tcn/hws/client.py:7-22)."""
import json
import socket

from .constants import CLIENT_CMDS, HWS_DUMP_NAME, SOCKET_FILENAME


def client_main(order: str, dump_name: str = HWS_DUMP_NAME, socket_filename: str = SOCKET_FILENAME):
    filtered_order = dict(CLIENT_CMDS[order])
    filtered_order["dump_name"] = dump_name
    data = json.dumps(filtered_order)
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    server.connect(socket_filename)
    server.send(data.encode("utf8"))
    server.close()


def cli(command: str, dump_name: str = HWS_DUMP_NAME):
    if command in CLIENT_CMDS.keys():
        client_main(command, dump_name)
    else:
        raise RuntimeError(f"[HWS Client] Unknown cmds {command} as first argument of the executable")
