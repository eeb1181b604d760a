"""Command line of the hardware sampler:

    python -m b200stencil.hws.cli server
    python -m b200stencil.hws.cli client {start,stop,dump,tick} [--name DUMP]
    python -m b200stencil.hws.cli envelop DUMP.npz

Same sub-commands as the reference's ``tcn-hws`` (/root/reference/src/tcn/hws/cli.py:17-52) minus ``graph``
(plotly is not in this image).  ``envelop`` prints the energy of a dump, integrated over its timestamps.
"""
from __future__ import annotations

import argparse
import sys
from typing import Optional, Sequence

from . import protocol


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="b200stencil-hws", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="what", required=True)
    sub.add_parser("server", help="run the sampling daemon on ./sockets-runtime/hws")
    c = sub.add_parser("client", help="send one order to the daemon")
    c.add_argument("command", choices=[o.verb for o in protocol.Order])
    c.add_argument("--name", default="hws", help="[dump] file name of the dump, without extension")
    e = sub.add_parser("envelop", help="energy of a dump")
    e.add_argument("dump")
    return ap


def main(argv: Optional[Sequence[str]] = None) -> int:
    ns = build_parser().parse_args(argv)
    if ns.what == "server":
        from . import server

        server.cli()
    elif ns.what == "client":
        from . import client

        client.send_order(ns.command, ns.name)
    else:
        from . import analysis

        d = analysis.load(ns.dump)
        print(analysis.energy_report(d["cpu_psu"], d["gpu_psu"], d["timestamps"], float(d["dt"])).describe())
    return 0


if __name__ == "__main__":
    sys.exit(main())
