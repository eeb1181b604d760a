"""-m gpu: the N>1 data plane on ONE device with virtual peers.

* ``b2s_halo_pull_c`` with per-rank fields living in one process: word [10] of every link holds the base address of
  the field of the (virtual) GPU that owns the source sub-domain;
* the library-owned exchange (``b2s_halo_init / alloc / plan / exchange`` -- csrc/halo_ctx.cu, k_halo_exchange):
  one thread per virtual rank, the ranks meet in the shared-memory rendezvous and handshake through their flag
  arrays exactly as separate processes on separate GPUs do; three epochs exercise the device-resident step counter;
* the gated stencil (``b2s_fv_tp2d_gated_c``): exchange forked onto the context's stream, one gate per sub-domain,
  sub-domain b computed while the halos of b+1.. arrive -- bit-identical to exchange-then-stencil, eagerly and
  under CUDA-graph replay.

Every halo cell is checked against the partitioner's geometric definition (global-id fields), edges and corners,
for the 1/2/4/8-GPU decompositions (SURVEY.md 8e "every halo cell must hold the ID of its geometric neighbour").
"""
import threading
import uuid

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from b200stencil import _abi, fields as F, stencils  # noqa: E402
from b200stencil.halo.device import HaloContext, build_plan_table  # noqa: E402
from b200stencil.halo.partitioner import CubedSpherePartitioner, expected_halo, global_id_field, layout_for  # noqa: E402
from b200stencil.halo.transport import FvTransport  # noqa: E402

from halo_util import batch_field, check_field  # noqa: E402


def _check(part, n_gpus, gpu, field, nk):
    nsub = part.subdomains_per_gpu(n_gpus)
    for b in range(nsub):
        want = expected_halo(part, gpu * nsub + b, nk)
        got = field[b].cpu().numpy()
        assert np.array_equal(got, want), f"gpu {gpu} sub-domain {b}: {np.argwhere(got != want)[:5]}"


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
@pytest.mark.parametrize("corners", [False, True])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_halo_pull_virtual_peers(n_gpus, corners, dtype):
    """b2s_halo_pull[_f32]_c: G fields in one process stand for the G GPUs; every link reads the owner's field by address."""
    N, nk = 24, 3
    part = CubedSpherePartitioner(N, layout_for(n_gpus), corners=corners)
    flds = [batch_field(part, n_gpus, g, nk, device="cuda", dtype=dtype, pad=2) for g in range(n_gpus)]
    for g in range(n_gpus):
        t = build_plan_table(part, n_gpus, g, flds[g], list(range(n_gpus)))
        t11 = t[:, :11].copy()
        t11[:, 10] = [flds[int(o)].data_ptr() for o in t[:, 10]]
        links = torch.from_numpy(t11).cuda()
        stencils.prepare_halo_pull(links, nk, flds[g], int((t11[:, 8] * t11[:, 9]).max()))()
    torch.cuda.synchronize()
    for g in range(n_gpus):
        _check(part, n_gpus, g, flds[g], nk)


def _run_ranks(world, body):
    """One thread per virtual rank on cuda:0; re-raises the first failure.

    The exchange kernels of ALL the virtual ranks must be resident on the one GPU at the same time (they wait for each
    other), so the ranks take several levels per work unit: 8 ranks x (links x levels) blocks would not fit otherwise.
    One rank per GPU -- the product configuration -- has the GPU to itself."""
    errors = []
    if world > 1 and _abi.get_option("halo_levels_per_unit") == 0:
        _abi.set_option("halo_levels_per_unit", 8)
        try:
            return _run_ranks(world, body)
        finally:
            _abi.set_option("halo_levels_per_unit", 0)

    def wrap(rank):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()):  # the legacy default stream would serialise the ranks' kernels
                body(rank)
                torch.cuda.current_stream().synchronize()
        except BaseException as exc:  # noqa: BLE001
            errors.append((rank, exc))

    ts = [threading.Thread(target=wrap, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(180)
    assert not any(t.is_alive() for t in ts), "a virtual rank is stuck"
    if errors:
        raise errors[0][1]


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
@pytest.mark.parametrize("corners", [False, True])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("push", ["staged", "inplace", False])
def test_device_exchange_virtual_ranks(n_gpus, corners, dtype, push, monkeypatch):
    """b2s_halo_init .. exchange .. finalize with one thread per rank: rendezvous, symmetric allocation, handshake,
    epoch counter; every halo cell against geometry after each of three exchanges.  ``push``: the plan carries the
    outgoing strips, so the ungated exchanges (reps 0 and 2) pull the same-GPU strips and PUSH what crosses ranks
    (announce / deliver flags, k_halo_exchange3) -- "staged": packed into the staging area behind the destination's field
    and unpacked there, "inplace": straight into the halo cells; the forked exchange of rep 1 goes the same way.
    False: pull only."""
    monkeypatch.setenv("B2S_RDV_TIMEOUT", "60")
    N, nk = 24, 3
    part = CubedSpherePartitioner(N, layout_for(n_gpus), corners=corners)
    nsub = part.subdomains_per_gpu(n_gpus)
    session = uuid.uuid4().hex

    def body(rank):
        ctx = HaloContext(rank, n_gpus, 0, session)
        try:
            f = ctx.field((part.nx + 6, part.ny + 6, nk), nsub, dtype, part=part if push == "staged" else None)
            ex = ctx.plan(f, part, push=bool(push))
            for rep in range(3):
                for b in range(nsub):
                    f[b].copy_(torch.from_numpy(global_id_field(part, rank * nsub + b, nk)))
                torch.cuda.current_stream().synchronize()
                ctx.barrier()  # every rank's interior is in place before anyone pulls (the test rewrites it on the host side)
                if rep == 1:
                    ex.start(gated=False)
                    ex.wait()
                else:
                    ex.update()
                torch.cuda.current_stream().synchronize()
                _check(part, n_gpus, rank, f, nk)
                ctx.barrier()
            assert ctx.status() == (3, 0)
            if n_gpus > 1:
                assert ex.remote_bytes > 0
        finally:
            ctx.finalize()

    _run_ranks(n_gpus, body)


@pytest.mark.parametrize("n_gpus", [2, 8])
@pytest.mark.parametrize("option", [("halo_variant", 4), ("halo_variant", 1), ("halo_handshake", 1), ("halo_levels_per_unit", 3)])
def test_exchange_variants_virtual_ranks(n_gpus, option):
    """The pull exchange in its other forms (b2s_set_option): the one-block handshake kernel followed by the flat-grid pull,
    the first version of the one-kernel exchange, the per-block handshake wait, several levels per work unit -- same halos."""
    N, nk = 24, 2  # few levels: version 1 takes one block per (link, level), and all ranks' blocks must be co-resident
    part = CubedSpherePartitioner(N, layout_for(n_gpus), corners=True)
    nsub = part.subdomains_per_gpu(n_gpus)
    session = uuid.uuid4().hex
    _abi.set_option(*option)

    def body(rank):
        ctx = HaloContext(rank, n_gpus, 0, session)
        try:
            f = ctx.field((part.nx + 6, part.ny + 6, nk), nsub, torch.float64)
            ex = ctx.plan(f, part)
            for rep in range(3):
                for b in range(nsub):
                    f[b].copy_(torch.from_numpy(global_id_field(part, rank * nsub + b, nk)))
                torch.cuda.current_stream().synchronize()
                ctx.barrier()
                ex.update()
                torch.cuda.current_stream().synchronize()
                _check(part, n_gpus, rank, f, nk)
                ctx.barrier()
            assert ctx.status() == (3, 0)
        finally:
            ctx.finalize()

    try:
        _run_ranks(n_gpus, body)
    finally:
        _abi.set_option(option[0], 0)


def _fv_fields(nsub, ni, nj, nk, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mk = lambda s, lo, hi: F.empty(s, dtype, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    crx, cry = mk((ni + 1, nj, nk), -0.9, 0.9), mk((ni, nj + 1, nk), -0.9, 0.9)
    return dict(crx=crx, xfx=mk((ni + 1, nj, nk), -1, 1), cry=cry, yfx=mk((ni, nj + 1, nk), -1, 1), rarea=mk((ni, nj), 0.9, 1.1),
                core=mk((ni, nj, nk), 0.5, 1.5))  # fmt: skip


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("N,nk", [(192, 3), (24, 2), (200, 2)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gated_step_equals_exchange_then_stencil(variant, N, nk, dtype):
    """world = 1 (six tiles on this GPU): the overlapped step -- exchange forked, fv_tp2d_gated, join -- must give the
    bits of exchange-then-fv_tp2d with either TMA kernel, repeatedly (the gate is lowered and raised every step) and
    when the step is captured into a CUDA graph and replayed."""
    part = CubedSpherePartitioner(N)
    ctx = HaloContext(0, 1, 0)
    _abi.set_option("fv_variant", variant)
    try:
        q = ctx.field((N + 6, N + 6, nk), 6, dtype)
        ex = ctx.plan(q, part)
        d = _fv_fields(6, N, N, nk, dtype, 11)
        q[:, 3:-3, 3:-3] = d["core"]
        ref_q = F.empty((N + 6, N + 6, nk), dtype, batch=6)
        ref_q.copy_(q)
        ref = F.zeros((N, N, nk), dtype, batch=6)
        FvTransport(part, 1, 0).step(ref_q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], ref)  # halo_move + fv_tp2d
        tr = FvTransport(part, 1, 0, exchange="device", halo_exchange=ex, overlap=True)
        for rep in range(3):
            q[:, :3] = -7.0  # scrub the west halo: the exchange has to refill it every step
            out = F.zeros((N, N, nk), dtype, batch=6)
            tr.step(q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
            torch.cuda.synchronize()
            assert torch.equal(q[:, :, 3:-3], ref_q[:, :, 3:-3]) and torch.equal(q[:, 3:-3], ref_q[:, 3:-3]), "halos differ"
            assert torch.equal(out, ref), f"rep {rep}: max |diff| = {(out - ref).abs().max().item():.3e}"
        assert ctx.status() == (3, 0)
        assert ctx.gate[:65].abs().sum().item() == 0, "the gates must be lowered after every gated launch"
        # serial device path
        out = F.zeros((N, N, nk), dtype, batch=6)
        FvTransport(part, 1, 0, exchange="device", halo_exchange=ex, overlap=False).step(q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
        assert torch.equal(out, ref)
        # fused: halo update + stencil in ONE launch (b2s_halo_fv_tp2d), eagerly and from a replayed graph
        fz = FvTransport(part, 1, 0, exchange="device", halo_exchange=ex, fused=True)
        for rep in range(2):
            q[:, -3:] = -9.0  # scrub the east halo
            out = F.zeros((N, N, nk), dtype, batch=6)
            launches = _abi.launch_count()
            fz.step(q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
            assert _abi.launch_count() - launches == 1, "the fused step is one kernel launch"
            torch.cuda.synchronize()
            assert torch.equal(q[:, :, 3:-3], ref_q[:, :, 3:-3]) and torch.equal(out, ref), f"fused rep {rep}"
        capf = torch.cuda.Stream()
        capf.wait_stream(torch.cuda.current_stream())
        gf = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gf, stream=capf):
            fz.step(q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
        for rep in range(2):
            q[:, :3] = -7.0
            out.zero_()
            gf.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, ref), f"fused graph replay {rep}"
        del gf
        assert ctx.gate[:65].abs().sum().item() == 0
        # CUDA graph: capture the forked step once, replay it
        out = F.zeros((N, N, nk), dtype, batch=6)
        tr.step(q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)  # marshal outside the capture
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            tr.step(q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
        for rep in range(3):
            q[:, :3] = -7.0
            out.zero_()
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, ref), f"graph replay {rep}"
        del graph
        epoch, status = ctx.status()
        assert status == 0 and epoch == 12
    finally:
        _abi.set_option("fv_variant", 0)
        ctx.finalize()


@pytest.mark.parametrize("mode", ["overlap", "fused", "serial"])
@pytest.mark.parametrize("n_gpus", [2, 8])
def test_gated_step_virtual_ranks(n_gpus, mode):
    """Handshake + exchange + gate together: every virtual rank runs overlapped (forked pull beside the gated stencil),
    fused (ONE kernel per step) or serial (mixed pull / push exchange, then the plain stencil) transport steps; results
    equal the exchange-in-process reference (CUDA halo_move tables) followed by the plain stencil.  The domains are
    small, so all the ranks' grids are co-resident on the GPU."""
    fused = mode == "fused"
    from b200stencil.halo.updater import exchange_in_process

    N, nk, dtype = 48, 2, torch.float64
    part = CubedSpherePartitioner(N, layout_for(n_gpus))
    nsub, ni, nj = part.subdomains_per_gpu(n_gpus), part.nx, part.ny
    data = [_fv_fields(nsub, ni, nj, nk, dtype, 100 + g) for g in range(n_gpus)]
    ref_q = []
    for g in range(n_gpus):
        qf = F.zeros((ni + 6, nj + 6, nk), dtype, batch=nsub)
        qf[:, 3:-3, 3:-3] = data[g]["core"]
        ref_q.append(qf)
    exchange_in_process(part, n_gpus, ref_q)
    refs = []
    for g in range(n_gpus):
        o = F.zeros((ni, nj, nk), dtype, batch=nsub)
        d = data[g]
        stencils.fv_tp2d(ref_q[g], d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], o)
        refs.append(o)
    torch.cuda.synchronize()
    session = uuid.uuid4().hex

    def body(rank):
        ctx = HaloContext(rank, n_gpus, 0, session)
        try:
            q = ctx.field((ni + 6, nj + 6, nk), nsub, dtype, part=part)
            q[:, 3:-3, 3:-3] = data[rank]["core"]
            ex = ctx.plan(q, part, push=True)  # the serial step pushes (staged) what crosses ranks; gated / fused pull
            tr = FvTransport(part, n_gpus, rank, exchange="device", halo_exchange=ex, overlap=mode != "serial", fused=fused)
            d = data[rank]
            torch.cuda.current_stream().synchronize()
            ctx.barrier()
            for rep in range(3):
                out = F.zeros((ni, nj, nk), dtype, batch=nsub)
                tr.step(q, d["crx"], d["xfx"], d["cry"], d["yfx"], d["rarea"], out)
                torch.cuda.current_stream().synchronize()
                assert torch.equal(out, refs[rank]), f"rank {rank} rep {rep}"
            ctx.check()
            ctx.barrier()
        finally:
            ctx.finalize()

    _run_ranks(n_gpus, body)


def test_halo_context_argument_errors():
    with pytest.raises(_abi.B200StencilError):
        HaloContext(3, 2, 0, "x")  # rank outside the world
    with pytest.raises(_abi.B200StencilError):
        HaloContext(0, 2, 0, "bad/name")
    ctx = HaloContext(0, 1, 0)
    try:
        part = CubedSpherePartitioner(12)
        outside = F.zeros((18, 18, 2), batch=6)  # not a symmetric allocation: fine for world 1 (all links are local)
        ex = ctx.plan(outside, part)
        ex.update()
        with pytest.raises(_abi.B200StencilError):
            ctx.free(12345)
        ex.start()
        with pytest.raises(_abi.B200StencilError):
            ex.start()  # twice without wait
        ex.wait()
    finally:
        ctx.finalize()
    with pytest.raises(_abi.B200StencilError):
        ctx2 = HaloContext(0, 1, 0)
        h = ctx2.handle
        ctx2.finalize()
        _abi.check("b2s_halo_barrier", _abi.load()[1].b2s_halo_barrier(h))  # a finalized handle is refused
