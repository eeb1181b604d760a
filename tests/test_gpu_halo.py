"""-m gpu: the CUDA halo_move kernel through the same link tables the NCCL path uses (virtual GPUs on
one device), and the interior/frame split of the transport step against the unsplit stencil."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from b200stencil.halo.partitioner import CubedSpherePartitioner, layout_for  # noqa: E402
from b200stencil.halo.transport import FvTransport, split_regions  # noqa: E402
from b200stencil.halo.updater import exchange_in_process  # noqa: E402

from halo_util import batch_field, check_field  # noqa: E402


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_global_id_exchange_cuda(n_gpus, dtype):
    N, nk = 24, 3
    part = CubedSpherePartitioner(N, layout_for(n_gpus))
    fields = [batch_field(part, n_gpus, g, nk, device="cuda", dtype=dtype, pad=2) for g in range(n_gpus)]
    exchange_in_process(part, n_gpus, fields)
    torch.cuda.synchronize()
    for g in range(n_gpus):
        check_field(part, n_gpus, g, fields[g], nk)


@pytest.mark.parametrize("shape", [(192, 40), (140, 20), (64, 64), (20, 9)])
def test_region_split_equals_full(shape):
    from b200stencil import fields as F
    from b200stencil import stencils

    ni, nj = shape
    nk, nb = 3, 2
    g = torch.Generator(device="cuda").manual_seed(1)
    mk = lambda s, lo, hi: F.empty(s, torch.float64, batch=nb).uniform_(lo, hi, generator=g)  # noqa: E731
    q = mk((ni + 6, nj + 6, nk), 0.5, 1.5)
    crx, cry = mk((ni + 1, nj, nk), -0.9, 0.9), mk((ni, nj + 1, nk), -0.9, 0.9)
    xfx, yfx = mk((ni + 1, nj, nk), -1, 1), mk((ni, nj + 1, nk), -1, 1)
    rarea = mk((ni, nj), 0.9, 1.1)
    full = F.zeros((ni, nj, nk), batch=nb)
    stencils.fv_tp2d(q, crx, xfx, cry, yfx, rarea, full)
    split = F.zeros((ni, nj, nk), batch=nb)
    interior, frame = split_regions(ni, nj)
    cover = np.zeros((ni, nj), int)
    for r in [interior] + frame:
        if r[1] > r[0] and r[3] > r[2]:
            stencils.fv_tp2d(q, crx, xfx, cry, yfx, rarea, split, region=r)
            cover[r[0] : r[1], r[2] : r[3]] += 1
    assert np.all(cover == 1), "interior + frame must tile the domain exactly once"
    # different tile geometries, explicit-rounding arithmetic (csrc/fv_math.cuh): bit-identical
    assert torch.equal(full, split), f"max |diff| = {(full - split).abs().max().item():.3e}"


def test_transport_single_gpu_fills_halos_then_steps():
    """G = 1: six tiles on one device; the step must equal 'exchange, then stencil'."""
    from b200stencil import fields as F
    from b200stencil import stencils

    N, nk = 24, 2
    part = CubedSpherePartitioner(N)
    tr = FvTransport(part, 1, 0)
    g = torch.Generator(device="cuda").manual_seed(2)
    mk = lambda s, lo, hi: F.empty(s, torch.float64, batch=6).uniform_(lo, hi, generator=g)  # noqa: E731
    q = mk((N + 6, N + 6, nk), 0.5, 1.5)
    crx, cry = mk((N + 1, N, nk), -0.9, 0.9), mk((N, N + 1, nk), -0.9, 0.9)
    xfx, yfx, rarea = mk((N + 1, N, nk), -1, 1), mk((N, N + 1, nk), -1, 1), mk((N, N), 0.9, 1.1)
    q2 = q.clone()
    out1, out2 = F.zeros((N, N, nk), batch=6), F.zeros((N, N, nk), batch=6)
    tr.step(q, crx, xfx, cry, yfx, rarea, out1)
    exchange_in_process(part, 1, [q2])
    stencils.fv_tp2d(q2, crx, xfx, cry, yfx, rarea, out2)
    assert torch.equal(out1, out2)
    # a halo cell now holds its neighbour's interior value: west halo of tile 1 <- north edge of tile 5
    assert torch.equal(q[0, 2, 3:-3, :], q[4, 3:-3, -4, :].flip(0))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_dycore_chain_matches_oracle(dtype):
    """BASELINE config 5 in miniature: halo update + fv_tp2d + pe_prefix + remap on six 12x12 tiles against
    the oracle applied stage by stage (halo filled by the partitioner's geometric definition)."""
    from b200stencil import fields as F
    from b200stencil.halo.partitioner import unfold
    from b200stencil.halo.transport import DycoreChain
    from oracle import numpy_oracle as orc

    N, nk = 12, 6
    part = CubedSpherePartitioner(N)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    rng = np.random.default_rng(9)
    core = rng.uniform(0.5, 1.5, (6, N, N, nk)).astype(npdt)
    crx = rng.uniform(-0.9, 0.9, (6, N + 1, N, nk)).astype(npdt)
    cry = rng.uniform(-0.9, 0.9, (6, N, N + 1, nk)).astype(npdt)
    xfx = (crx * rng.uniform(0.9, 1.1, crx.shape)).astype(npdt)
    yfx = (cry * rng.uniform(0.9, 1.1, cry.shape)).astype(npdt)
    rarea = rng.uniform(0.9, 1.1, (6, N, N)).astype(npdt)
    delp = (rng.uniform(0.5, 1.5, (6, N, N, nk)) * 1e5 / nk).astype(npdt)
    ptop = 1.0
    # oracle: halo by geometry, then the three stencils
    ref = np.zeros((6, N, N, nk), npdt)
    pe2_all = np.zeros((6, N, N, nk + 1), npdt)
    for t in range(6):
        qh = np.zeros((N + 6, N + 6, nk), npdt)
        for li in range(-3, N + 3):
            for lj in range(-3, N + 3):
                if not (0 <= li < N) and not (0 <= lj < N):
                    continue  # corners are not read by the stencil
                t2, i2, j2 = unfold(t, li, lj, N)
                qh[li + 3, lj + 3] = core[t2, i2, j2]
        adv = np.zeros((N, N, nk), npdt)
        orc.fv_tp2d(qh, crx[t], xfx[t], cry[t], yfx[t], rarea[t], adv)
        pe1 = np.zeros((N, N, nk + 1), npdt)
        orc.pe_prefix(delp[t], ptop, pe1)
        sig = (np.arange(nk + 1) / nk).astype(npdt)
        pe2 = (pe1[:, :, :1] + (pe1[:, :, -1:] - pe1[:, :, :1]) * sig).astype(npdt)
        pe2[:, :, -1] = pe1[:, :, -1]
        pe2_all[t] = pe2
        orc.remap(pe1, adv, pe2, ref[t])
    # device
    dev = lambda a, shape: _batch(F, a, shape, dtype)  # noqa: E731
    q = F.zeros((N + 6, N + 6, nk), dtype, batch=6)
    q[:, 3:-3, 3:-3] = torch.from_numpy(core).cuda()
    d_crx, d_xfx = dev(crx, (N + 1, N, nk)), dev(xfx, (N + 1, N, nk))
    d_cry, d_yfx = dev(cry, (N, N + 1, nk)), dev(yfx, (N, N + 1, nk))
    d_rarea, d_delp, d_pe2 = dev(rarea, (N, N)), dev(delp, (N, N, nk)), dev(pe2_all, (N, N, nk + 1))
    q_adv, pe1_d, q_new = F.zeros((N, N, nk), dtype, batch=6), F.zeros((N, N, nk + 1), dtype, batch=6), F.zeros((N, N, nk), dtype, batch=6)
    chain = DycoreChain(FvTransport(part, 1, 0), ptop)
    chain.step(q, d_crx, d_xfx, d_cry, d_yfx, d_rarea, d_delp, d_pe2, q_adv, pe1_d, q_new)
    q_fused = F.zeros((N, N, nk), dtype, batch=6)
    DycoreChain(FvTransport(part, 1, 0), ptop, fused=True).step(q, d_crx, d_xfx, d_cry, d_yfx, d_rarea, d_delp, d_pe2, q_adv, pe1_d, q_fused)
    assert torch.equal(q_fused, q_new), "remap_delp must reproduce pe_prefix + remap bit for bit"
    got = q_new.cpu().numpy()
    rtol = 1e-12 if dtype == torch.float64 else 1e-5
    assert np.all(np.abs(got - ref) <= rtol * np.maximum(np.abs(ref), np.abs(ref).max()))


def _batch(F, a, shape, dtype):
    t = F.empty(shape, dtype, batch=a.shape[0])
    t.copy_(torch.from_numpy(a))
    return t


@pytest.mark.parametrize("n_gpus", [1, 8])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_split_transport_on_the_cubed_sphere(n_gpus, dtype):
    """S5b end to end: corner-including halo update (CUDA link tables, virtual ranks for the 8-GPU layout) followed
    by fv_tp2d_split with the cube-corner flags, against the oracle run per sub-domain on halos filled by the CPU
    interpreter of the same links."""
    from b200stencil import fields as F
    from b200stencil import stencils
    from b200stencil.halo.transport import SplitTransport
    from halo_util import cpu_mover
    from oracle import inputs as gen
    from oracle.c_oracle import COracle

    N, nk = 12, 2
    part = CubedSpherePartitioner(N, layout_for(n_gpus), corners=True)
    nsub, nx, ny = part.subdomains_per_gpu(n_gpus), part.nx, part.ny
    g = torch.Generator(device="cuda").manual_seed(3)
    mk = lambda s, lo, hi: F.empty(s, dtype, batch=nsub).uniform_(lo, hi, generator=g)  # noqa: E731
    sets = []
    for _ in range(n_gpus):
        crx, cry = mk((nx + 1, ny + 6, nk), -0.45, 0.45), mk((nx + 6, ny + 1, nk), -0.45, 0.45)
        sets.append(dict(q=mk((nx + 6, ny + 6, nk), 0.5, 1.5), crx=crx, cry=cry, xfx=mk(crx.shape[1:], 0.9, 1.1).mul_(crx),
                         yfx=mk(cry.shape[1:], 0.9, 1.1).mul_(cry), area=mk((nx + 6, ny + 6), 0.9, 1.1), rarea=mk((nx, ny), 0.9, 1.1)))  # fmt: skip
    # reference halos (edges and corners): the CPU interpreter of the same link tables
    host_q = [F.empty((nx + 6, ny + 6, nk), dtype, device="cpu", batch=nsub).copy_(s["q"]) for s in sets]
    exchange_in_process(part, n_gpus, host_q, mover=cpu_mover)
    # device: the same exchange through the CUDA kernels, then the stencil with the corner flags
    exchange_in_process(part, n_gpus, [s["q"] for s in sets])
    np_dt = np.float64 if dtype == torch.float64 else np.float32
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    corc = COracle()
    for gpu, s in enumerate(sets):
        assert torch.equal(s["q"].cpu(), host_q[gpu])
        tr = SplitTransport(part, n_gpus, gpu)
        out = F.zeros((nx, ny, nk), dtype, batch=nsub)
        stencils.fv_tp2d_split(s["q"], s["crx"], s["xfx"], s["cry"], s["yfx"], s["area"], s["rarea"], out,
                               corner_flags=tr.corner_flags(s["q"].device))
        for b in range(nsub):
            f = {k: gen.as_ifirst(v[b].cpu().numpy().astype(np_dt)) for k, v in s.items()}
            ref = gen.ifirst_empty((nx, ny, nk), np_dt)
            ref[...] = 0
            corc.fv_tp2d_split(f["q"], f["crx"], f["xfx"], f["cry"], f["yfx"], f["area"], f["rarea"], ref,
                               corner_flags=part.cube_corner_flags(gpu * nsub + b))
            assert np.abs(out[b].cpu().numpy() - ref).max() <= tol * np.abs(ref).max(), (gpu, b)
        if n_gpus == 1:  # the step object does exchange + stencil in one call
            out2 = F.zeros((nx, ny, nk), dtype, batch=nsub)
            tr.step(s["q"], s["crx"], s["xfx"], s["cry"], s["yfx"], s["area"], s["rarea"], out2)
            assert torch.equal(out2, out)
