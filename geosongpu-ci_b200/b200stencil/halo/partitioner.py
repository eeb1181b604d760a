"""Cubed-sphere domain decomposition and halo connectivity.

The reference only states the decomposition -- 6 tiles x layout (lx, ly), one rank per sub-domain,
``total_ranks = 6 * layout[0] * layout[1]``
(/root/reference/src/tcn/validation/serialbox/serialbox_dat_to_netcdf.py:91-93; ``NX = lx``,
``NY = 6 * ly`` in /root/reference/src/tcn/benchmark/geos_log_parser.py:48-58; ``layout_1/layout_2``
in /root/reference/src/tcn/py_ftn_interface/example_def_dycore.yaml:41-42) -- the connectivity itself
lives in NDSL, which is not vendored.  It is therefore built here from geometry: six faces of a
cube, each with its own (i, j) axes, a halo cell being the cell reached by unfolding the neighbouring
face across the shared edge.  The faces are oriented so that the FV3 convention holds
(odd tiles: east -> n+1, north -> n+2 rotated; even tiles: north -> n+1, east -> n+2 rotated), which
tests/test_partitioner.py asserts together with adjacency symmetry.

Pure Python/NumPy: no torch, no CUDA.  Sub-domains are decoupled from GPUs: a GPU hosts a
contiguous block of ``6 * lx * ly / G`` sub-domains as a batch axis.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

# face t: (origin corner, e_i, e_j) on the cube [-1, 1]^3; outward normal = e_i x e_j
_X, _Y, _Z = np.array([1, 0, 0]), np.array([0, 1, 0]), np.array([0, 0, 1])
_FACES = [
    (+_X, +_Y, +_Z),  # tile 1: face +x, i -> +y, j -> +z
    (+_Y, -_X, +_Z),  # tile 2: face +y
    (+_Z, -_X, -_Y),  # tile 3: face +z (north pole)
    (-_X, -_Z, -_Y),  # tile 4: face -x
    (-_Y, -_Z, +_X),  # tile 5: face -y
    (-_Z, +_Y, +_X),  # tile 6: face -z (south pole)
]
WEST, EAST, SOUTH, NORTH = 0, 1, 2, 3
SW, SE, NW, NE = 4, 5, 6, 7  # corner blocks (Link.edge codes), only with corners=True
EDGE_NAMES = ("west", "east", "south", "north", "south-west", "south-east", "north-west", "north-east")


def corner_fill_source(gi: int, gj: int, N: int, direction: int) -> Tuple[int, int]:
    """FV3's copy_corners ([recalled] fv_mp_mod / tp_core.F90 copy_corners): the halo cell that a cube-corner
    halo cell (outside the tile in BOTH directions, 0-based tile coordinates) copies, for sweeps in x
    (direction 1: the row is continued around the corner into the west / east neighbour) or in y (direction 2).
    The source lies outside the tile in ONE direction only, so it has an owner."""
    w, s = gi < 0, gj < 0
    e, n = gi >= N, gj >= N
    assert (w or e) and (s or n), "not a corner cell"
    if direction == 1:
        if w and s:
            return gj, -gi - 1  # q(i,j) = q(j, 1-i)
        if e and s:
            return N - 1 - gj, gi - N  # q(i,j) = q(npy-j, i-npx+1)
        if e and n:
            return gj, 2 * N - 1 - gi  # q(i,j) = q(j, 2*npx-1-i)
        return N - 1 - gj, gi + N  # nw: q(i,j) = q(npy-j, i-1+npx)
    if w and s:
        return -gj - 1, gi  # q(i,j) = q(1-j, i)
    if e and s:
        return N + gj, N - 1 - gi  # q(i,j) = q(npy+j-1, npx-i)
    if e and n:
        return 2 * N - 1 - gj, gi  # q(i,j) = q(2*npy-1-j, i)
    return gj - N, N - 1 - gi  # nw: q(i,j) = q(j+1-npx, npy-i)


def _face(t: int):
    n, ei, ej = _FACES[t]
    assert np.array_equal(np.cross(ei, ej), n), "face axes must give an outward normal"
    return n, ei, ej


def unfold(t: int, gi: int, gj: int, N: int) -> Tuple[int, int, int]:
    """Tile-local cell (gi, gj) of tile ``t``, possibly outside [0, N) in ONE direction -> the
    (tile, i, j) that owns it.  Integer arithmetic on a cube of side 2N (cell centres at odd offsets)."""
    if 0 <= gi < N and 0 <= gj < N:
        return t, gi, gj
    if not (0 <= gi < N or 0 <= gj < N):
        raise ValueError("corner halo cells have no unique owner on the cubed sphere")
    n, ei, ej = _face(t)
    u, v = 2 * gi + 1, 2 * gj + 1  # in [0, 2N] inside the face
    origin = N * (n - ei - ej)  # corner (i, j) = (0, 0)
    # clamp to the face, and carry the overshoot down the neighbouring face (direction -n)
    uc, vc = min(max(u, 0), 2 * N), min(max(v, 0), 2 * N)
    over = abs(u - uc) + abs(v - vc)
    p = origin + uc * ei + vc * ej - over * n
    for t2 in range(6):
        n2, ei2, ej2 = _face(t2)
        if t2 != t and int(p @ n2) == N:
            o2 = N * (n2 - ei2 - ej2)
            u2, v2 = int((p - o2) @ ei2), int((p - o2) @ ej2)
            assert u2 % 2 == 1 and v2 % 2 == 1
            return t2, (u2 - 1) // 2, (v2 - 1) // 2
    raise AssertionError("unfolded point is on no face")


@dataclass(frozen=True)
class Link:
    """One affine halo strip copy: halo cell (d, p) of ``dst`` <- interior cell of ``src``.

    ``d`` = depth into the halo (0 nearest the edge), ``p`` = position along the edge.  Cell
    coordinates are sub-domain-local compute coordinates (halo cells are negative or >= n):
        dst cell = (di0 + d*ddi + p*dpi, dj0 + d*ddj + p*dpj)     src cell likewise with s*.
    """

    dst: int
    src: int
    edge: int
    nd: int
    np_: int
    di0: int
    dj0: int
    ddi: int
    ddj: int
    dpi: int
    dpj: int
    si0: int
    sj0: int
    sdi: int
    sdj: int
    spi: int
    spj: int

    @property
    def key(self):
        return (self.dst, self.edge, self.di0, self.dj0)


class CubedSpherePartitioner:
    """6 tiles x (lx, ly) sub-domains of an N x N-cell tile; rank = tile*lx*ly + sy*lx + sx."""

    def __init__(self, N: int, layout: Tuple[int, int] = (1, 1), halo: int = 3, corners: bool = False):
        """``corners=True`` adds the four corner blocks of every halo to the links (needed by stencils that
        sweep twice, S5b fv_tp2d_split): a corner block inside the tile or across ONE tile edge comes from the
        diagonal neighbour; at the eight cube corners, where no such neighbour exists, it is filled by FV3's
        copy_corners rule for x-sweeps (direction 1), folded into the link so that it reads the owner's interior
        directly (one pass, no ordering between edge and corner copies)."""
        lx, ly = layout
        self.corners = corners
        if N % lx or N % ly:
            raise ValueError(f"layout {layout} does not divide C{N}")
        self.N, self.lx, self.ly, self.halo = N, lx, ly, halo
        self.nx, self.ny = N // lx, N // ly
        if min(self.nx, self.ny) < halo:
            raise ValueError("sub-domains thinner than the halo are not supported")
        self.total_ranks = 6 * lx * ly  # serialbox_dat_to_netcdf.py:91-93

    # ---- sub-domain <-> rank -------------------------------------------------------------------
    def rank_of(self, t: int, sx: int, sy: int) -> int:
        return t * self.lx * self.ly + sy * self.lx + sx

    def subdomain(self, rank: int) -> Tuple[int, int, int]:
        t, r = divmod(rank, self.lx * self.ly)
        sy, sx = divmod(r, self.lx)
        return t, sx, sy

    def owner(self, t: int, gi: int, gj: int) -> Tuple[int, int, int]:
        """Owning (rank, local i, local j) of tile cell (gi, gj), unfolding across tile edges."""
        t2, i2, j2 = unfold(t, gi, gj, self.N)
        sx, sy = i2 // self.nx, j2 // self.ny
        return self.rank_of(t2, sx, sy), i2 - sx * self.nx, j2 - sy * self.ny

    def owner_any(self, t: int, gi: int, gj: int, direction: int = 1) -> Tuple[int, int, int]:
        """:meth:`owner` extended to cube-corner halo cells through :func:`corner_fill_source`."""
        N = self.N
        if not (0 <= gi < N or 0 <= gj < N):
            gi, gj = corner_fill_source(gi, gj, N, direction)
        return self.owner(t, gi, gj)

    def cube_corner_flags(self, rank: int) -> int:
        """Bit mask of the halo corners of ``rank`` that sit at a cube corner: 1 SW, 2 SE, 4 NW, 8 NE."""
        _, sx, sy = self.subdomain(rank)
        w, e, s, n = sx == 0, sx == self.lx - 1, sy == 0, sy == self.ly - 1
        return (1 if w and s else 0) | (2 if e and s else 0) | (4 if w and n else 0) | (8 if e and n else 0)

    def _corner_links(self, rank: int) -> List[Link]:
        t, sx, sy = self.subdomain(rank)
        h, nx, ny = self.halo, self.nx, self.ny
        out: List[Link] = []
        # (code, first cell, step in depth (i), step along (j)): d runs over i, p over j
        blocks = [(SW, (-1, -1), (-1, 0), (0, -1)), (SE, (nx, -1), (1, 0), (0, -1)),
                  (NW, (-1, ny), (-1, 0), (0, 1)), (NE, (nx, ny), (1, 0), (0, 1))]  # fmt: skip
        for code, (i0, j0), (ddi, ddj), (dpi, dpj) in blocks:
            src = np.empty((h, h, 3), dtype=np.int64)
            for d in range(h):
                for p in range(h):
                    li, lj = i0 + d * ddi + p * dpi, j0 + d * ddj + p * dpj
                    src[d, p] = self.owner_any(t, sx * nx + li, sy * ny + lj, 1)
            dd, pp = np.meshgrid(np.arange(h), np.arange(h), indexing="ij")
            sd = (src[1, 0, 1:] - src[0, 0, 1:]) if h > 1 else np.zeros(2, dtype=np.int64)
            sp = (src[0, 1, 1:] - src[0, 0, 1:]) if h > 1 else np.zeros(2, dtype=np.int64)
            one_owner = np.all(src[:, :, 0] == src[0, 0, 0])
            affine = one_owner and np.array_equal(src[:, :, 1], src[0, 0, 1] + dd * sd[0] + pp * sp[0]) \
                and np.array_equal(src[:, :, 2], src[0, 0, 2] + dd * sd[1] + pp * sp[1])
            if affine:
                out.append(Link(rank, int(src[0, 0, 0]), code, h, h, i0, j0, ddi, ddj, dpi, dpj,
                                int(src[0, 0, 1]), int(src[0, 0, 2]), int(sd[0]), int(sd[1]), int(sp[0]), int(sp[1])))  # fmt: skip
            else:  # a block that straddles two owners (exotic layouts): one link per cell
                for d in range(h):
                    for p in range(h):
                        out.append(Link(rank, int(src[d, p, 0]), code, 1, 1, i0 + d * ddi + p * dpi, j0 + d * ddj + p * dpj,
                                        0, 0, 0, 0, int(src[d, p, 1]), int(src[d, p, 2]), 0, 0, 0, 0))  # fmt: skip
        return out

    # ---- halo links -------------------------------------------------------------------------------
    def links_into(self, rank: int) -> List[Link]:
        """Every strip copy that fills the edge halo of ``rank`` (and, with ``corners=True``, its corner blocks)."""
        t, sx, sy = self.subdomain(rank)
        h, nx, ny = self.halo, self.nx, self.ny
        out: List[Link] = []
        # (edge, first halo cell, step in depth, step along the edge, edge length)
        strips = [
            (WEST, (-1, 0), (-1, 0), (0, 1), ny),
            (EAST, (nx, 0), (1, 0), (0, 1), ny),
            (SOUTH, (0, -1), (0, -1), (1, 0), nx),
            (NORTH, (0, ny), (0, 1), (1, 0), nx),
        ]
        for edge, (i0, j0), (ddi, ddj), (dpi, dpj), length in strips:
            src = np.empty((h, length, 3), dtype=np.int64)
            for d in range(h):
                for p in range(length):
                    li, lj = i0 + d * ddi + p * dpi, j0 + d * ddj + p * dpj
                    src[d, p] = self.owner(t, sx * nx + li, sy * ny + lj)
            # split the edge into runs owned by one source rank
            p0 = 0
            while p0 < length:
                p1 = p0 + 1
                while p1 < length and src[0, p1, 0] == src[0, p0, 0]:
                    p1 += 1
                run = src[:, p0:p1]
                s_rank = int(run[0, 0, 0])
                assert np.all(run[:, :, 0] == s_rank), "a halo strip run must have one owner at every depth"
                si0, sj0 = int(run[0, 0, 1]), int(run[0, 0, 2])
                sd = (run[1, 0, 1:] - run[0, 0, 1:]) if h > 1 else np.zeros(2, dtype=np.int64)
                sp = (run[0, 1, 1:] - run[0, 0, 1:]) if p1 - p0 > 1 else np.zeros(2, dtype=np.int64)
                dd, pp = np.meshgrid(np.arange(h), np.arange(p1 - p0), indexing="ij")
                assert np.array_equal(run[:, :, 1], si0 + dd * sd[0] + pp * sp[0]), "strip is not affine"
                assert np.array_equal(run[:, :, 2], sj0 + dd * sd[1] + pp * sp[1]), "strip is not affine"
                out.append(
                    Link(rank, s_rank, edge, h, p1 - p0, i0 + p0 * dpi, j0 + p0 * dpj, ddi, ddj, dpi, dpj,
                         si0, sj0, int(sd[0]), int(sd[1]), int(sp[0]), int(sp[1]))
                )  # fmt: skip
                p0 = p1
        if self.corners:
            out += self._corner_links(rank)
        return out

    def all_links(self) -> List[Link]:
        if not hasattr(self, "_links"):
            links: List[Link] = []
            for r in range(self.total_ranks):
                links += self.links_into(r)
            self._links = sorted(links, key=lambda l: l.key)
        return self._links

    # ---- GPUs ---------------------------------------------------------------------------------------
    def subdomains_per_gpu(self, n_gpus: int) -> int:
        if self.total_ranks % n_gpus:
            raise ValueError(f"{self.total_ranks} sub-domains do not spread evenly over {n_gpus} GPUs")
        return self.total_ranks // n_gpus

    def gpu_of(self, rank: int, n_gpus: int) -> int:
        return rank // self.subdomains_per_gpu(n_gpus)

    def local_index(self, rank: int, n_gpus: int) -> int:
        return rank % self.subdomains_per_gpu(n_gpus)


def layout_for(n_gpus: int) -> Tuple[int, int]:
    """Sub-tile layout used at each GPU count: 3 sub-domains per GPU from 2 GPUs up."""
    return {1: (1, 1), 2: (1, 1), 3: (1, 1), 6: (1, 1), 4: (1, 2), 8: (2, 2), 12: (2, 1), 24: (2, 2)}[n_gpus]


def global_id_field(part: CubedSpherePartitioner, rank: int, nk: int = 1) -> np.ndarray:
    """Halo-padded field [i, j, k] of ``rank`` whose interior holds a globally unique cell id
    (tile, gi, gj, k) -> ((tile*N + gj)*N + gi)*nk + k, halo = -1.  The adjacency tests fill the
    halo by exchange and compare with :func:`expected_halo`."""
    t, sx, sy = part.subdomain(rank)
    h, nx, ny, N = part.halo, part.nx, part.ny, part.N
    f = np.full((nx + 2 * h, ny + 2 * h, nk), -1.0)
    gi = sx * nx + np.arange(nx)
    gj = sy * ny + np.arange(ny)
    ids = (t * N + gj[None, :]) * N + gi[:, None]
    f[h : h + nx, h : h + ny, :] = ids[:, :, None] * nk + np.arange(nk)[None, None, :]
    return f


def expected_halo(part: CubedSpherePartitioner, rank: int, nk: int = 1, direction: int = 1) -> np.ndarray:
    """What the halo-padded global-id field of ``rank`` must hold after an exchange.  Without ``part.corners`` the
    corner blocks stay -1; with it they hold the diagonal neighbour's ids, or at a cube corner the ids FV3's
    copy_corners rule for ``direction`` puts there (the exchange fills direction 1)."""
    t, sx, sy = part.subdomain(rank)
    h, nx, ny, N = part.halo, part.nx, part.ny, part.N
    f = global_id_field(part, rank, nk)
    for li in range(-h, nx + h):
        for lj in range(-h, ny + h):
            inside_i, inside_j = 0 <= li < nx, 0 <= lj < ny
            if inside_i and inside_j:
                continue
            if not inside_i and not inside_j and not part.corners:
                continue
            gi, gj = sx * nx + li, sy * ny + lj
            if not (0 <= gi < N or 0 <= gj < N):
                gi, gj = corner_fill_source(gi, gj, N, direction)
            t2, i2, j2 = unfold(t, gi, gj, N)
            f[li + h, lj + h, :] = ((t2 * N + j2) * N + i2) * nk + np.arange(nk)
    return f
