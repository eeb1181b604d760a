"""Cubed-sphere connectivity and the host side of the halo exchange (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch

from b200stencil.halo.partitioner import (EAST, NORTH, SOUTH, WEST, CubedSpherePartitioner, _face, corner_fill_source,
                                           layout_for, unfold)
from b200stencil.halo.updater import HaloPlan, HaloUpdater, exchange_in_process

from halo_util import batch_field, check_field, cpu_mover


def test_fv3_tile_convention():
    """Odd tiles: east -> n+1, north -> n+2 (rotated), west -> n-2, south -> n-1;
    even tiles: north -> n+1, east -> n+2 (rotated), west -> n-1, south -> n-2 (1-based n)."""
    N = 8
    wrap = lambda n: (n - 1) % 6 + 1  # noqa: E731
    for t in range(6):
        n = t + 1
        w, e = unfold(t, -1, 3, N), unfold(t, N, 3, N)
        s, no = unfold(t, 3, -1, N), unfold(t, 3, N, N)
        if n % 2 == 1:
            assert e == (wrap(n + 1) - 1, 0, 3)  # west edge of n+1, same orientation
            assert no[0] == wrap(n + 2) - 1 and no[1] == 0  # west edge of n+2 (rotated)
            assert w[0] == wrap(n - 2) - 1 and w[2] == N - 1  # north edge of n-2 (rotated)
            assert s == (wrap(n - 1) - 1, 3, N - 1)  # north edge of n-1, same orientation
        else:
            assert no == (wrap(n + 1) - 1, 3, 0)  # south edge of n+1, same orientation
            assert e[0] == wrap(n + 2) - 1 and e[2] == 0  # south edge of n+2 (rotated)
            assert w == (wrap(n - 1) - 1, N - 1, 3)  # east edge of n-1
            assert s[0] == wrap(n - 2) - 1 and s[1] == N - 1  # east edge of n-2 (rotated)


def test_adjacency_is_symmetric():
    """If cell B is the halo neighbour of A across an edge, stepping back from B reaches A."""
    N = 6
    for t in range(6):
        for p in range(N):
            for (gi, gj), back in (((-1, p), None), ((N, p), None), ((p, -1), None), ((p, N), None)):
                t2, i2, j2 = unfold(t, gi, gj, N)
                # the interior cell adjacent to the halo cell
                ai, aj = min(max(gi, 0), N - 1), min(max(gj, 0), N - 1)
                # from (t2,i2,j2), one of its four outward steps must land on (t, ai, aj)
                hits = []
                for di, dj in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                    ni, nj = i2 + di, j2 + dj
                    if 0 <= ni < N and 0 <= nj < N:
                        continue
                    hits.append(unfold(t2, ni, nj, N))
                assert (t, ai, aj) in hits


def test_corner_cells_have_no_owner():
    with pytest.raises(ValueError):
        unfold(0, -1, -1, 8)


@pytest.mark.parametrize("layout", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 2)])
def test_links_cover_every_edge_halo_cell_once(layout):
    N, h = 12, 3
    part = CubedSpherePartitioner(N, layout, h)
    assert part.total_ranks == 6 * layout[0] * layout[1]
    for r in range(part.total_ranks):
        seen = np.zeros((part.nx + 2 * h, part.ny + 2 * h), dtype=int)
        for l in part.links_into(r):
            for d in range(l.nd):
                for p in range(l.np_):
                    seen[l.di0 + d * l.ddi + p * l.dpi + h, l.dj0 + d * l.ddj + p * l.dpj + h] += 1
                    si, sj = l.si0 + d * l.sdi + p * l.spi, l.sj0 + d * l.sdj + p * l.spj
                    assert 0 <= si < part.nx and 0 <= sj < part.ny  # sources are interior cells
        want = np.ones_like(seen)
        want[h:-h, h:-h] = 0
        for ci in (slice(0, h), slice(-h, None)):
            for cj in (slice(0, h), slice(-h, None)):
                want[ci, cj] = 0
        assert np.array_equal(seen, want)


def _centre(t, i, j, N):
    """3-D position of the centre of cell (i, j) of tile t on the cube of side 2N (integer coordinates)."""
    n, ei, ej = _face(t)
    return N * (n - ei - ej) + (2 * i + 1) * ei + (2 * j + 1) * ej


def test_copy_corners_continues_rows_and_columns_around_the_cube_corner():
    """FV3's copy_corners rule, checked against geometry instead of against itself: walking along a halo ROW
    (direction 1) or COLUMN (direction 2) from the edge halo into the corner block, consecutive cells must be
    neighbours on the cube (centres 2 apart on one face, sqrt 2 apart across an edge), i.e. the fill continues the
    line around the corner into the face that really lies there."""
    N, h = 8, 3
    for t in range(6):
        for ci, cj in ((-1, -1), (N, -1), (-1, N), (N, N)):  # first corner cell of SW, SE, NW, NE
            si, sj = (-1 if ci < 0 else 1), (-1 if cj < 0 else 1)
            for depth in range(h):
                # direction 1: the row gj (in the south / north halo), walked in i from inside the tile into the corner
                gj = cj + sj * depth
                chain = [(ci - si, gj)] + [(ci + si * a, gj) for a in range(h)]
                cells = [unfold(t, *chain[0], N)] + [unfold(t, *corner_fill_source(i, j, N, 1), N) for i, j in chain[1:]]
                for a, b in zip(cells[:-1], cells[1:]):
                    assert np.linalg.norm(_centre(*a, N) - _centre(*b, N)) <= 2.0 + 1e-9, (t, ci, cj, depth, "x")
                # direction 2: the column gi (in the west / east halo), walked in j
                gi = ci + si * depth
                chain = [(gi, cj - sj)] + [(gi, cj + sj * a) for a in range(h)]
                cells = [unfold(t, *chain[0], N)] + [unfold(t, *corner_fill_source(i, j, N, 2), N) for i, j in chain[1:]]
                for a, b in zip(cells[:-1], cells[1:]):
                    assert np.linalg.norm(_centre(*a, N) - _centre(*b, N)) <= 2.0 + 1e-9, (t, ci, cj, depth, "y")


@pytest.mark.parametrize("layout", [(1, 1), (1, 2), (2, 2), (3, 2)])
def test_corner_links_cover_every_halo_cell_once(layout):
    N, h = 12, 3
    part = CubedSpherePartitioner(N, layout, h, corners=True)
    for r in range(part.total_ranks):
        seen = np.zeros((part.nx + 2 * h, part.ny + 2 * h), dtype=int)
        for l in part.links_into(r):
            for d in range(l.nd):
                for p in range(l.np_):
                    seen[l.di0 + d * l.ddi + p * l.dpi + h, l.dj0 + d * l.ddj + p * l.dpj + h] += 1
                    si, sj = l.si0 + d * l.sdi + p * l.spi, l.sj0 + d * l.sdj + p * l.spj
                    assert 0 <= si < part.nx and 0 <= sj < part.ny  # corner fills read interiors too: one pass
        want = np.ones_like(seen)
        want[h:-h, h:-h] = 0
        assert np.array_equal(seen, want)
    flags = [part.cube_corner_flags(r) for r in range(part.total_ranks)]
    assert sum(bin(f).count("1") for f in flags) == 24  # 6 tiles x 4 corners, whatever the layout


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
def test_global_id_exchange_with_corners(n_gpus):
    """Corner blocks: the diagonal neighbour's ids inside a tile and across one tile edge, the copy_corners
    (direction 1) ids at the cube corners -- through the same pack/segment/unpack tables."""
    N, nk = 12, 2
    part = CubedSpherePartitioner(N, layout_for(n_gpus), corners=True)
    fields = [batch_field(part, n_gpus, g, nk) for g in range(n_gpus)]
    exchange_in_process(part, n_gpus, fields, mover=cpu_mover)
    for g in range(n_gpus):
        check_field(part, n_gpus, g, fields[g], nk)
        assert (fields[g] >= 0).all()  # no halo cell left unfilled


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
@pytest.mark.parametrize("pad", [0, 2])
def test_global_id_exchange_virtual_gpus(n_gpus, pad):
    """Every halo cell ends up holding the id of its geometric neighbour, for the
    decompositions used at 1/2/4/8 GPUs, through the same pack/segment/unpack tables as NCCL."""
    N, nk = 12, 2
    part = CubedSpherePartitioner(N, layout_for(n_gpus))
    fields = [batch_field(part, n_gpus, g, nk, pad=pad) for g in range(n_gpus)]
    exchange_in_process(part, n_gpus, fields, mover=cpu_mover)
    for g in range(n_gpus):
        check_field(part, n_gpus, g, fields[g], nk)


def test_send_and_receive_segments_agree():
    part = CubedSpherePartitioner(12, (2, 2))
    plans = [HaloPlan(part, 8, g) for g in range(8)]
    for g, pl in enumerate(plans):
        for peer in pl.peers:
            assert pl.segment_elems(pl.send.get(peer, []), 5) == plans[peer].segment_elems(plans[peer].recv.get(g, []), 5)
        assert g not in pl.peers
    # one GPU: everything is local
    assert HaloPlan(CubedSpherePartitioner(12), 1, 0).peers == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, N, nk):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = CubedSpherePartitioner(N, layout_for(world))
        f = batch_field(part, world, rank, nk)
        up = HaloUpdater(part, world, rank, mover=cpu_mover)
        up.start(f)
        up.wait()
        check_field(part, world, rank, f, nk)
        assert up.bytes_sent_per_update > 0
        up.update(f)  # idempotent
        check_field(part, world, rank, f, nk)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_gloo(world):
    """The N>1 path over torch.distributed (gloo, CPU): start()/wait() with real isend/irecv."""
    import torch.multiprocessing as mp

    mp.spawn(_gloo_worker, args=(world, _free_port(), 12, 2), nprocs=world, join=True)


def test_plan_tables_name_the_owner_of_every_strip():
    """halo/device.py build_plan_table (the host table of b2s_halo_plan): 12 words per link, [10] = session rank that
    owns the source strip, [11] = destination sub-domain, offsets relative to the field; the device-side handshake relies on adjacency being symmetric
    (whoever I wait for also waits for me), and every halo cell must be covered exactly once."""
    import torch

    from b200stencil import fields
    from b200stencil.halo import device

    for n_gpus in (1, 2, 4, 8):
        part = CubedSpherePartitioner(24, layout_for(n_gpus), 3)
        nsub = part.subdomains_per_gpu(n_gpus)
        f = fields.empty((part.nx + 6, part.ny + 6, 5), torch.float64, "cpu", batch=nsub)
        ranks = [(3 * g + 1) % n_gpus for g in range(n_gpus)] if n_gpus in (2, 4, 8) else [0]  # a non-identity rank map
        neighbours = {}
        for gpu in range(n_gpus):
            t = device.build_plan_table(part, n_gpus, gpu, f, ranks)
            assert t.shape[1] == device.PLAN_WORDS and t.dtype == np.int64
            assert len(t) >= 4 * nsub  # at least one strip per edge of every sub-domain
            assert set(int(r) for r in t[:, 10]) <= set(ranks) and set(int(b) for b in t[:, 11]) == set(range(nsub))
            neighbours[ranks[gpu]] = {int(r) for r in t[:, 10] if r != ranks[gpu]}
            # destination cells: every edge-halo cell of every local sub-domain exactly once, none in the interior
            hits = np.zeros(f.numel() + 64, dtype=np.int32)
            for L in t:
                d, p = np.meshgrid(np.arange(L[8]), np.arange(L[9]), indexing="ij")
                np.add.at(hits, (L[4] + d * L[5] + p * L[6]).ravel(), 1)
            assert hits.max() == 1 and hits.sum() == nsub * 2 * 3 * (part.nx + part.ny)
        for a_, ns_ in neighbours.items():
            for b_ in ns_:
                assert a_ in neighbours[b_]
