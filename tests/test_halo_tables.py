"""The link tables of the library-owned exchange (halo/device.py build_plan_table), executed in NumPy: no GPU needed.

Every GPU is a flat array (its field followed by its staging area, as ``HaloContext.field(part=...)`` lays them out);
the rows are run the way k_halo_exchange3 runs them -- same-GPU strips pulled, crossing strips pushed by their owner
(in place, or packed into the destination's staging area and unpacked there after the delivery) -- and every halo
cell must then hold the id of its geometric neighbour, for the 2/4/8-GPU decompositions, with and
without corner blocks.  Also: the two ends of every crossing strip agree on its place and layout in the staging area,
and the staged segments of a GPU are disjoint."""
import numpy as np
import pytest
import torch

from b200stencil.halo.device import LINK_OUT, LINK_PUSHED, LINK_STAGED, build_plan_table, staging_elements
from b200stencil.halo.partitioner import CubedSpherePartitioner, expected_halo, global_id_field, layout_for


def _world(n_gpus, corners, staged, N=24, nk=3):
    part = CubedSpherePartitioner(N, layout_for(n_gpus), corners=corners)
    nsub = part.subdomains_per_gpu(n_gpus)
    ni, nj = part.nx + 6, part.ny + 6
    nip = (ni + 1) // 2 * 2
    numel = nsub * nk * nj * nip
    stage = staging_elements(part, n_gpus, nk) if staged else 0
    at = (numel + 15) // 16 * 16
    flats = [torch.full((at + stage,), -5.0, dtype=torch.float64) for _ in range(n_gpus)]
    fields = [fl[:numel].view(nsub, nk, nj, nip).permute(0, 3, 2, 1)[:, :ni] for fl in flats]
    for g in range(n_gpus):
        for b in range(nsub):
            fields[g][b].copy_(torch.from_numpy(global_id_field(part, g * nsub + b, nk)))
    tables = [build_plan_table(part, n_gpus, g, fields[g], list(range(n_gpus)), push=True, staging_offset=at if staged else None)
              for g in range(n_gpus)]  # fmt: skip
    return part, nsub, nk, at, stage, flats, fields, tables


def _copy(src, dst, r, nk):
    d, p, k = np.meshgrid(np.arange(int(r[8])), np.arange(int(r[9])), np.arange(nk), indexing="ij")
    dst[r[4] + d * r[5] + p * r[6] + k * r[7]] = src[r[0] + d * r[1] + p * r[2] + k * r[3]]


@pytest.mark.parametrize("staged", [False, True])
@pytest.mark.parametrize("corners", [False, True])
@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_mixed_exchange_tables_fill_every_halo_cell(n_gpus, corners, staged):
    part, nsub, nk, at, stage, flats, fields, tables = _world(n_gpus, corners, staged)
    mem = [fl.numpy() for fl in flats]
    for g, t in enumerate(tables):  # before the deliveries: same-GPU pulls and pushes
        for r in t:
            mark = int(r[11])
            if mark & LINK_OUT:
                assert int(r[10]) != g
                _copy(mem[g], mem[int(r[10])], r, nk)
            elif not mark & (LINK_PUSHED | LINK_STAGED):
                assert int(r[10]) == g, "an unmarked row of a push table must be a same-GPU strip"
                _copy(mem[g], mem[g], r, nk)
    for g, t in enumerate(tables):  # after the deliveries: unpack
        for r in t:
            if int(r[11]) & LINK_STAGED:
                _copy(mem[g], mem[g], r, nk)
    for g in range(n_gpus):
        for b in range(nsub):
            assert np.array_equal(fields[g][b].numpy(), expected_halo(part, g * nsub + b, nk)), (g, b)


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_both_ends_of_a_staged_strip_agree(n_gpus):
    part, nsub, nk, at, stage, flats, fields, tables = _world(n_gpus, True, True)
    outs, unpacks, pushed = {}, {}, 0
    for g, t in enumerate(tables):
        for r in t:
            mark = int(r[11])
            if mark & LINK_OUT:  # (destination GPU, first staging element) -> owner, packed layout, strip size
                outs[(int(r[10]), int(r[4]))] = (g, tuple(int(x) for x in r[5:10]))
            if mark & LINK_STAGED:
                unpacks[(g, int(r[0]))] = (int(r[10]), tuple(int(x) for x in r[1:4]) + (int(r[8]), int(r[9])))
            pushed += bool(mark & LINK_PUSHED)
    assert outs == unpacks and len(outs) == pushed > 0
    for g, t in enumerate(tables):
        segs = sorted((int(r[0]), int(r[8] * r[9]) * nk) for r in t if int(r[11]) & LINK_STAGED)
        assert segs[0][0] >= at and segs[-1][0] + segs[-1][1] <= at + stage
        for (a, n), (b, _) in zip(segs, segs[1:]):
            assert a + n <= b, "staged segments overlap"
        for r in t:  # what crosses NVLink is contiguous: a packed strip has a unit stride and a dense level stride
            if int(r[11]) & LINK_OUT:
                assert 1 in (abs(int(r[5])), abs(int(r[6]))) and int(r[7]) == int(r[8] * r[9])


def test_pull_only_table_is_unchanged_by_the_push_options():
    part = CubedSpherePartitioner(24, layout_for(8))
    f = torch.zeros(3, 2, part.ny + 6, part.nx + 6).permute(0, 3, 2, 1)
    plain = build_plan_table(part, 8, 3, f, list(range(8)))
    marked = build_plan_table(part, 8, 3, f, list(range(8)), push=True)
    incoming = marked[(marked[:, 11] & (LINK_OUT | LINK_STAGED)) == 0].copy()
    incoming[:, 11] &= 0xFFFF
    assert np.array_equal(plain, incoming)
