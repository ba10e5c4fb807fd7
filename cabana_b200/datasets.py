"""Seeded synthetic particle sets for the BASELINE.json configs and the reference fixtures.

Host-side numpy only (inputs are generated on the host and uploaded).
All generators use the counter-based Philox bit generator keyed by the seed so the
inputs do not depend on device or thread count.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ParticleSet:
    name: str
    xyz: np.ndarray  # (n,3) float64
    grid_min: tuple
    grid_max: tuple
    radius: float
    cell_ratio: float = 1.0

    @property
    def n(self) -> int:
        return self.xyz.shape[0]


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=int(seed)))


def uniform_box(n: int, seed: int, radius: float = 3.0, cell_ratio: float = 1.0) -> ParticleSet:
    """cfg1/cfg2: iid uniform in [0, 1.3 n^(1/3)]^3.

    Box rule: benchmark/core/Cabana_NeighborVerletPerformance.cpp:71-72.
    """
    hi = 1.3 * float(n) ** (1.0 / 3.0)
    xyz = _rng(seed).random((n, 3)) * hi
    # random() is in [0,1): strictly inside the box as locatePoint requires.
    return ParticleSet(f"uniform_{n}", xyz, (0.0,) * 3, (hi,) * 3, radius, cell_ratio)


FCC_DENSITY = 0.8442


def fcc_lattice(cells: int, radius: float = 2.8, cell_ratio: float = 1.0,
                jitter: float = 0.0, seed: int = 20240103,
                cells_yz: int | None = None) -> ParticleSet:
    """cfg3: perfect FCC at reduced density 0.8442, 4*cells^3 atoms (cells=159 -> 16 078 716).

    Atoms at (i,j,k)*a + basis*a + a/4 so none sits on a cell face; box [0, cells*a]^3,
    non-periodic.  Order: i slowest, then j, k, basis (lattice order, NOT cell order).
    `cells_yz` lets x and y/z extents differ (slab-shaped samples).
    """
    a = (4.0 / FCC_DENSITY) ** (1.0 / 3.0)
    cy = cells if cells_yz is None else cells_yz
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    i = np.arange(cells, dtype=np.float64)
    j = np.arange(cy, dtype=np.float64)
    ii, jj, kk = np.meshgrid(i, j, j, indexing="ij")
    corner = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)  # (c^3,3)
    xyz = (corner[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a + 0.25 * a
    hi = (cells * a, cy * a, cy * a)
    if jitter > 0.0:
        xyz = xyz + _rng(seed).normal(0.0, jitter, xyz.shape)
        for d in range(3):
            np.clip(xyz[:, d], 0.0, np.nextafter(hi[d], 0.0), out=xyz[:, d])
    return ParticleSet(f"fcc_{cells}", np.ascontiguousarray(xyz), (0.0,) * 3, hi, radius, cell_ratio)


def clustered(n: int, seed: int = 20240104, radius: float = 3.0, cell_ratio: float = 1.0,
              n_blobs: int = 64, contrast: float = 10.0) -> ParticleSet:
    """cfg4: 50% uniform background + 50% in Gaussian blobs with ~`contrast`x peak density."""
    hi = 1.3 * float(n) ** (1.0 / 3.0)
    rng = _rng(seed)
    n_bg = n // 2
    n_bl = n - n_bg
    bg = rng.random((n_bg, 3)) * hi
    rho_bg = n_bg / hi**3
    per_blob = n_bl / n_blobs
    # peak density of an isotropic Gaussian: N / ((2 pi)^(3/2) s^3) = (contrast-1) * rho_bg
    s = (per_blob / ((contrast - 1.0) * rho_bg * (2 * np.pi) ** 1.5)) ** (1.0 / 3.0)
    centres = rng.random((n_blobs, 3)) * hi
    which = rng.integers(0, n_blobs, n_bl)
    bl = centres[which] + rng.normal(0.0, s, (n_bl, 3))
    xyz = np.concatenate([bg, bl], axis=0)
    # reflect stragglers back into the box (clipping would stack them on the faces and
    # create coincident points, which half lists drop -- SURVEY.md Appendix B.4)
    xyz = np.abs(xyz)
    xyz = hi - np.abs(hi - xyz)
    np.clip(xyz, 0.0, np.nextafter(hi, 0.0), out=xyz)
    xyz = xyz[rng.permutation(n)]
    return ParticleSet(f"clustered_{n}", np.ascontiguousarray(xyz), (0.0,) * 3, (hi,) * 3, radius, cell_ratio)


# ----------------------------------------------------------------------------- reference fixtures
def fixture_random300(seed: int = 300, n: int = 300) -> ParticleSet:
    """NeighborListTestData<3> (core/unit_test/neighbor_unit_test.hpp:995-1045).

    300 random particles, r = 2.32, box [-5.3 r, 4.7 r]^3, cell_size_ratio 0.5.  The
    reference draws from Kokkos' XorShift64 pool which is not reproducible here, so the
    points are re-drawn with Philox; expectations come from the N^2 oracle as in the
    reference test.
    """
    r = 2.32
    lo, hi = -5.3 * r, 4.7 * r
    xyz = lo + _rng(seed).random((n, 3)) * (hi - lo)
    return ParticleSet("random300", xyz, (lo,) * 3, (hi,) * 3, r, 0.5)


def fixture_ordered(particle_x: int, m: int = 3) -> ParticleSet:
    """NeighborListTestDataOrdered (neighbor_unit_test.hpp:1101-1157).

    n^3 lattice in [0,5]^3, spacing dx = 5/n, particles at dx/2 + dx*i, r = m*dx + 1e-7.
    """
    n = particle_x
    dx = (5.0 - 0.0) / n
    pid = np.arange(n**3)
    i = pid // (n * n)
    j = (pid // n) % n
    k = pid % n
    xyz = np.stack([dx / 2 + dx * i, dx / 2 + dx * j, dx / 2 + dx * k], axis=1).astype(np.float64)
    return ParticleSet(f"ordered_{n}", xyz, (0.0,) * 3, (5.0,) * 3, m * dx + 1e-7, 0.5)


def fixture_lcl_grid(nx: int = 10) -> ParticleSet:
    """LCLTestData<3> (core/unit_test/tstLinkedCellList.hpp:30-105).

    One particle at the centre of each unit cell of a 10^3 grid, created x-fastest
    (particle_id = i + j*nx + k*nx*nx) -- the reverse of the binned order.
    """
    pid = np.arange(nx**3)
    i = pid % nx
    j = (pid // nx) % nx
    k = pid // (nx * nx)
    xyz = np.stack([0.0 + (i + 0.5) * 1.0, 0.0 + (j + 0.5) * 1.0, 0.0 + (k + 0.5) * 1.0], axis=1)
    return ParticleSet("lcl_grid", xyz.astype(np.float64), (0.0,) * 3, (float(nx),) * 3, 1.0, 1.0)


def fixture_tutorial81() -> ParticleSet:
    """Tutorial set: 81 particles, 3 coincident per cell centre of a 3^3 unit grid, r = 0.25.

    example/core_tutorial/10_neighbor_parallel_for/neighbor_parallel_for_example.cpp
    ("two neighbors each", :163).
    """
    pts = []
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for _ in range(3):
                    pts.append((0.5 + i, 0.5 + j, 0.5 + k))
    xyz = np.array(pts, dtype=np.float64)
    return ParticleSet("tutorial81", xyz, (0.0,) * 3, (3.0,) * 3, 0.25, 1.0)


def near_cutoff_adversarial(seed: int = 77, n_pairs: int = 400, radius: float = 1.0,
                            box: float = 40.0) -> ParticleSet:
    """pairs at |x_i - x_j| within a few ulp of r (closed
    cutoff and cell-prune rounding), scattered over a box so they straddle cell faces."""
    rng = _rng(seed)
    a = rng.random((n_pairs, 3)) * (box - 4 * radius) + 2 * radius
    dirs = rng.normal(size=(n_pairs, 3))
    # a third axis-aligned so that distances along one axis hit r exactly
    dirs[: n_pairs // 3] = 0.0
    dirs[np.arange(n_pairs // 3), rng.integers(0, 3, n_pairs // 3)] = 1.0
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    scale = np.full(n_pairs, radius)
    ulps = rng.integers(-4, 5, n_pairs)
    for _ in range(4):
        up = ulps > 0
        dn = ulps < 0
        scale[up] = np.nextafter(scale[up], np.inf)
        scale[dn] = np.nextafter(scale[dn], -np.inf)
        ulps = ulps - np.sign(ulps)
    b = a + dirs * scale[:, None]
    xyz = np.concatenate([a, b], axis=0)
    np.clip(xyz, 0.0, np.nextafter(box, 0.0), out=xyz)
    return ParticleSet("near_cutoff", np.ascontiguousarray(xyz), (0.0,) * 3, (box,) * 3, radius, 1.0)
