"""``python -m b200stencil.hws.cli server | client {start,stop,dump,tick} [--name N] | envelop FILE``
(reference: tcn-hws, /root/reference/src/tcn/hws/cli.py:12-52; the plotly ``graph`` command is left out:
plotly is not in this image)."""
import click

from . import analysis, client as hws_client, server as hws_server


@click.group()
def cli():
    pass


@cli.command()
def server():
    hws_server.cli()


@cli.command()
@click.argument("command")
@click.option("--name", default="hws", help="[dump] Filename for the .npz dump")
def client(command: str, name: str):
    hws_client.cli(command, name)


@cli.command()
@click.argument("data_filepath")
def envelop(data_filepath: str):
    d = analysis.load_data(data_filepath)
    analysis.energy_envelop_calculation(d["cpu_psu"], d["gpu_psu"], d["timestamps"], float(d["dt"]))


if __name__ == "__main__":
    cli()
