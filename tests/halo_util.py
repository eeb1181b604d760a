"""Test-only helpers for the halo exchange: a CPU interpreter of link tables (the product path
runs them with the CUDA halo_move kernel) and global-id field builders."""
import numpy as np
import torch

from b200stencil.halo.partitioner import CubedSpherePartitioner, expected_halo, global_id_field


def cpu_mover(links: torch.Tensor, nk: int, src: torch.Tensor, dst: torch.Tensor, max_strip: int = 0) -> None:
    """dst[doff + d*dsd + p*dsp + k*dsk] = src[soff + d*ssd + p*ssp + k*ssk] for every link."""
    for L in links.tolist():
        soff, ssd, ssp, ssk, doff, dsd, dsp, dsk, nd, np_ = L
        d, p, k = torch.meshgrid(torch.arange(nd), torch.arange(np_), torch.arange(nk), indexing="ij")
        vals = src[(soff + d * ssd + p * ssp + k * ssk).reshape(-1)].clone()
        dst[(doff + d * dsd + p * dsp + k * dsk).reshape(-1)] = vals


def batch_field(part: CubedSpherePartitioner, n_gpus: int, gpu: int, nk: int, device="cpu", dtype=torch.float64, pad=0):
    """GPU ``gpu``'s halo-padded batch field [b,i,j,k] (i-fastest) of global ids, halos = -1."""
    nsub = part.subdomains_per_gpu(n_gpus)
    h = part.halo
    nip = part.nx + 2 * h + pad
    store = torch.full((nsub, nk, part.ny + 2 * h, nip), -1.0, dtype=dtype, device=device)
    f = store.permute(0, 3, 2, 1)[:, : part.nx + 2 * h]
    for b in range(nsub):
        f[b].copy_(torch.from_numpy(global_id_field(part, gpu * nsub + b, nk)))
    return f


def check_field(part, n_gpus, gpu, field, nk):
    nsub = part.subdomains_per_gpu(n_gpus)
    for b in range(nsub):
        want = expected_halo(part, gpu * nsub + b, nk)
        got = field[b].cpu().numpy()
        assert np.array_equal(got, want), f"gpu {gpu} sub-domain {b}: {np.argwhere(got != want)[:5]}"
