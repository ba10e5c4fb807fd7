"""Edge cases of the Verlet build on the GPU against the oracle: empty and tiny inputs,
points on the grid's upper faces, one-cell grids, every refinement / stencil regime
(cell_size_ratio 0.2 .. 3), very dense cells (the general-kernel fallback with rows that
span several candidate windows), empty ranges, and handle reuse across regimes."""
import numpy as np
import pytest
import torch

from cabana_b200 import datasets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    assert torch.cuda.is_available()
    from cabana_b200 import core

    return core


def _check(orc, cb, xyz, r, ratio, gmin, gmax, begin=None, end=None, max_neigh=0,
           algos=("full", "half"), layouts=("csr", "2d")):
    n = xyz.shape[0]
    b = 0 if begin is None else begin
    e = n if end is None else end
    for algo in algos:
        for layout in layouts:
            a = cb.FULL if algo == "full" else cb.HALF
            lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
            x = cb.view_from_array(xyz) if n else cb.Slice(torch.zeros(3, dtype=torch.float64, device="cuda"), 0, 3, 1, 1, 3)
            lst = cb.VerletList(x, b, e, r, ratio, gmin, gmax, max_neigh, algorithm=a, layout=lay)
            ox = orc.view_from_xyz(xyz) if n else orc.PositionsView(np.zeros(3), 0, 3, 1, 1)
            ref = orc.verlet_build(ox, b, e, r, ratio, gmin, gmax, max_neigh=max_neigh, algo=a, layout=lay)
            counts = lst._data.counts.cpu().numpy()
            assert np.array_equal(counts, ref.counts), (algo, layout)
            offsets = lst._data.offsets.cpu().numpy() if lay == cb.CSR else None
            nb = lst._data.neighbors.cpu().numpy()
            got, _ = orc.sorted_rows_flat(lay, counts, offsets, nb, lst.width)
            assert np.array_equal(got, ref.sorted_rows_flat()[0]), (algo, layout)
            assert lst.total == ref.total and lst._data.max_n == ref.max_n
            if lay == cb.LAYOUT_2D:
                assert lst.width == ref.width and lst.refilled == ref.refilled


def test_empty_and_tiny_inputs(orc, cb):
    box = ((0.0,) * 3, (4.0,) * 3)
    _check(orc, cb, np.zeros((0, 3)), 1.0, 1.0, *box)
    _check(orc, cb, np.array([[1.0, 1.0, 1.0]]), 1.0, 1.0, *box)
    _check(orc, cb, np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 1.0]]), 1.0, 1.0, *box)  # coincident
    _check(orc, cb, np.array([[1.0, 1.0, 1.0], [1.5, 1.0, 1.0], [3.9, 3.9, 3.9]]), 1.0, 1.0, *box)
    # empty range and single-row range
    ps = datasets.fixture_random300()
    _check(orc, cb, ps.xyz, ps.radius, 0.5, ps.grid_min, ps.grid_max, begin=100, end=100)
    _check(orc, cb, ps.xyz, ps.radius, 0.5, ps.grid_min, ps.grid_max, begin=299, end=300)


def test_points_on_upper_faces_and_single_cell_grid(orc, cb):
    rng = np.random.default_rng(2)
    xyz = rng.random((500, 3)) * 4.0
    xyz[:40] = 4.0 * (rng.random((40, 3)) > 0.5)  # corners: 0 or exactly grid_max
    xyz[40:80, 0] = 4.0                              # upper x face (exact edge clamp, :180)
    _check(orc, cb, xyz, 1.0, 1.0, (0.0,) * 3, (4.0,) * 3)
    # radius larger than half the box: 1-cell grid (nx = 1)
    _check(orc, cb, xyz, 3.0, 1.0, (0.0,) * 3, (4.0,) * 3)
    # anisotropic box, negative coordinates
    xyz2 = rng.random((2000, 3)) * np.array([9.0, 2.5, 5.0]) + np.array([-4.0, -1.0, 10.0])
    _check(orc, cb, xyz2, 0.8, 1.0, (-4.0, -1.0, 10.0), (5.0, 1.5, 15.0))


@pytest.mark.parametrize("ratio", [0.2, 0.25, 0.34, 0.5, 0.75, 1.0, 1.5, 2.0, 3.0])
def test_every_cell_size_ratio(orc, cb, ratio):
    ps = datasets.uniform_box(6000, 17, radius=2.1)
    _check(orc, cb, ps.xyz, 2.1, ratio, ps.grid_min, ps.grid_max, layouts=("csr",))


def test_exact_multiple_box(orc, cb):
    # box = k * r exactly: the reach needs three refined cells
    rng = np.random.default_rng(8)
    xyz = rng.random((8000, 3)) * 12.0
    xyz[:500] = np.round(xyz[:500])  # lattice points at exact multiples of r: distances == r
    xyz[:500] = np.clip(xyz[:500], 0.0, 12.0)
    xyz = np.unique(xyz, axis=0)
    _check(orc, cb, xyz, 1.0, 1.0, (0.0,) * 3, (12.0,) * 3)
    _check(orc, cb, xyz, 2.0, 0.5, (0.0,) * 3, (12.0,) * 3, layouts=("csr",))


def test_very_dense_cells_take_the_general_kernel(orc, cb):
    # 6000 particles inside one cutoff sphere + sparse background: candidate lists of
    # several thousand entries (> the column kernel's limits) and rows of thousands of ids
    rng = np.random.default_rng(4)
    dense = 10.0 + rng.normal(0.0, 0.4, (6000, 3))
    bg = rng.random((4000, 3)) * 20.0
    xyz = np.concatenate([dense, bg])
    np.clip(xyz, 0.0, np.nextafter(20.0, 0.0), out=xyz)
    _check(orc, cb, xyz, 1.5, 1.0, (0.0,) * 3, (20.0,) * 3, algos=("full",), layouts=("csr",))
    _check(orc, cb, xyz, 1.5, 1.0, (0.0,) * 3, (20.0,) * 3, algos=("half",), layouts=("2d",))
    _check(orc, cb, xyz, 1.5, 1.0, (0.0,) * 3, (20.0,) * 3, max_neigh=64, algos=("full",), layouts=("2d",))


def test_handle_reuse_across_regimes(orc, cb):
    lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
    for seed, n, r, ratio in ((1, 5000, 3.0, 1.0), (2, 300, 3.0, 0.2), (3, 20000, 2.0, 2.0), (4, 50, 3.0, 1.0)):
        ps = datasets.uniform_box(n, seed, radius=r)
        x = cb.slice_from_array(ps.xyz, vlen=16, extra=3)
        lst.build(x, 0, n, r, ratio, ps.grid_min, ps.grid_max)
        ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, n, r, ratio, ps.grid_min, ps.grid_max)
        counts = lst._data.counts.cpu().numpy()
        assert np.array_equal(counts, ref.counts)
        got, _ = orc.sorted_rows_flat(orc.CSR, counts, lst._data.offsets.cpu().numpy(),
                                      lst._data.neighbors.cpu().numpy(), 0)
        assert np.array_equal(got, ref.sorted_rows_flat()[0])


def test_speculative_fill_paths(orc, cb, monkeypatch):
    """Rebuilds on one handle launch the CSR fill pass before the host has seen the list size: the
    list that grows past the array of the previous build (guard fails, pass re-launched after the
    allocation), the one that fits (speculation holds), the one that shrinks, and the switch
    CB_VERLET_SPECULATE=0 must all give the oracle's list."""
    ps = datasets.uniform_box(12_000, 20240811, radius=3.0)
    x = cb.slice_from_array(ps.xyz, vlen=32)
    ox = orc.view_from_xyz(ps.xyz)
    lst = cb.VerletList(algorithm=cb.HALF, layout=cb.CSR)
    for r, spec in ((1.5, "1"), (3.0, "1"), (3.0, "1"), (2.2, "1"), (3.0, "0"), (2.9, "1")):
        monkeypatch.setenv("CB_VERLET_SPECULATE", spec)
        lst.build(x, 0, ps.n, r, 1.0, ps.grid_min, ps.grid_max)
        ref = orc.verlet_build(ox, 0, ps.n, r, 1.0, ps.grid_min, ps.grid_max, algo=orc.HALF)
        counts = lst._data.counts.cpu().numpy()
        assert lst.total == ref.total and np.array_equal(counts, ref.counts), r
        got, _ = orc.sorted_rows_flat(orc.CSR, counts, lst._data.offsets.cpu().numpy(),
                                      lst._data.neighbors.cpu().numpy(), 0)
        assert np.array_equal(got, ref.sorted_rows_flat()[0]), r


# ---------------------------------------------------------------- kernel generations
@pytest.fixture
def verlet_impl():
    """Select the kernel generation for one test (the library reads CB_VERLET_IMPL and
    CB_TILE_STAGING per build): v0, v1, v2 (default: TMA bulk-copy staging) or v2-async (v2 with
    the per-lane cp.async gather)."""
    import os

    old = {k: os.environ.get(k) for k in ("CB_VERLET_IMPL", "CB_TILE_STAGING")}

    def select(name):
        os.environ.pop("CB_VERLET_IMPL", None)
        os.environ.pop("CB_TILE_STAGING", None)
        if name == "v2-async":
            os.environ["CB_TILE_STAGING"] = "async"
        elif name not in (None, "v2"):
            os.environ["CB_VERLET_IMPL"] = name

    yield select
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("impl", ["v0", "v1", "v2", "v2-async"])
def test_kernel_generations_match_oracle(orc, cb, verlet_impl, impl):
    """v2 (tile kernels) is the default; v1 (refined grid + FP32 SIMT filter) and v0
    (reference-shaped exact FP64) stay selectable with CB_VERLET_IMPL and must give the same
    lists."""
    verlet_impl(impl)
    ps = datasets.fixture_random300()
    _check(orc, cb, ps.xyz, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max)
    _check(orc, cb, ps.xyz, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max, begin=75, end=225)
    ps = datasets.near_cutoff_adversarial()
    _check(orc, cb, ps.xyz, ps.radius, 1.0, ps.grid_min, ps.grid_max, layouts=("csr",))
    ps = datasets.fcc_lattice(8, jitter=0.05)
    _check(orc, cb, ps.xyz, ps.radius, 1.0, ps.grid_min, ps.grid_max)
    ps = datasets.clustered(20_000)
    _check(orc, cb, ps.xyz, ps.radius, 0.5, ps.grid_min, ps.grid_max, layouts=("csr",))


@pytest.mark.parametrize("impl", ["v0", "v1", "v2"])
def test_cell_size_ratio_one_tenth(orc, cb, verlet_impl, impl):
    # ratio 0.1: stencil range 10 cells; v1 hands this regime to the v0 kernels
    verlet_impl(impl)
    ps = datasets.uniform_box(3000, 23, radius=2.1)
    _check(orc, cb, ps.xyz, 2.1, 0.1, ps.grid_min, ps.grid_max, layouts=("csr",))
    _check(orc, cb, ps.xyz, 2.1, 0.125, ps.grid_min, ps.grid_max, algos=("half",), layouts=("2d",))


# ---------------------------------------------------------------- large extents
def _near_cutoff_far_from_origin(origin, extent, radius, n_pairs, n_bg, seed):
    """Pairs within +-4 ulp of the cutoff (a third axis-aligned) plus a uniform background,
    in the box [origin, origin + extent]: coordinates ~1e4 make ulp(x) ~2e-12, so the filter's
    error bound (which grows with the extent) and the exact tier's band carry real load."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    origin = np.asarray(origin, dtype=np.float64)
    extent = np.asarray(extent, dtype=np.float64)
    a = origin + 2 * radius + rng.random((n_pairs, 3)) * (extent - 4 * radius)
    dirs = rng.normal(size=(n_pairs, 3))
    k = n_pairs // 3
    dirs[:k] = 0.0
    dirs[np.arange(k), rng.integers(0, 3, k)] = 1.0
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    scale = np.full(n_pairs, radius)
    ulps = rng.integers(-4, 5, n_pairs)
    for _ in range(4):
        scale[ulps > 0] = np.nextafter(scale[ulps > 0], np.inf)
        scale[ulps < 0] = np.nextafter(scale[ulps < 0], -np.inf)
        ulps = ulps - np.sign(ulps)
    b = a + dirs * scale[:, None]
    bg = origin + rng.random((n_bg, 3)) * extent
    xyz = np.concatenate([a, b, bg])
    hi = np.nextafter(origin + extent, -np.inf)
    return np.ascontiguousarray(np.minimum(np.maximum(xyz, origin), hi))


@pytest.mark.parametrize("impl", ["v1", "v2"])
def test_large_extent_far_from_origin(orc, cb, verlet_impl, impl):
    """grid_min = 1e4, box 2000 wide (the 8-GPU cfg5 box is 2080 wide): the filters' error
    bounds scale with the extent, the prune band with |coordinate|."""
    verlet_impl(impl)
    origin, extent, r = (1.0e4, 1.0e4, -2.0e4), (2000.0, 60.0, 45.0), 2.8
    xyz = _near_cutoff_far_from_origin(origin, extent, r, n_pairs=3000, n_bg=400_000, seed=91)
    gmin = origin
    gmax = tuple(o + e for o, e in zip(origin, extent))
    _check(orc, cb, xyz, r, 1.0, gmin, gmax, layouts=("csr",))
    # the z extent as the long one (the tile kernels' pencils run along z)
    xyz2 = np.ascontiguousarray(xyz[:, ::-1])
    _check(orc, cb, xyz2, r, 1.0, gmin[::-1], gmax[::-1], algos=("full",), layouts=("csr",))
    _check(orc, cb, xyz2, r, 0.5, gmin[::-1], gmax[::-1], algos=("half",), layouts=("2d",))


@pytest.mark.parametrize("case", ["z_long", "x_long", "fcc", "dense"])
def test_filter_selftest(cb, case):
    """The tensor-core filter's observed error stays inside the proven bound (every tested pair is
    compared with the exact FP64 value), also when the box is 2000 wide and far from the origin,
    and no value outside the exact tier's band ever has the wrong sign."""
    r = 2.8
    if case in ("z_long", "x_long"):
        origin, extent = (1.0e4, 1.0e4, -2.0e4), (45.0, 60.0, 2000.0)
        xyz = _near_cutoff_far_from_origin(origin, extent, r, n_pairs=2000, n_bg=300_000, seed=92)
        gmin = origin
        gmax = tuple(o + e for o, e in zip(origin, extent))
        if case == "x_long":
            xyz, gmin, gmax = np.ascontiguousarray(xyz[:, ::-1]), gmin[::-1], gmax[::-1]
    elif case == "fcc":
        ps = datasets.fcc_lattice(20, jitter=0.05)
        xyz, gmin, gmax = ps.xyz, ps.grid_min, ps.grid_max
    else:
        ps = datasets.clustered(60_000)
        xyz, gmin, gmax, r = ps.xyz, ps.grid_min, ps.grid_max, ps.radius
    x = cb.view_from_array(xyz)
    for algo in (cb.FULL, cb.HALF):
        lst = cb.VerletList(algorithm=algo, layout=cb.CSR)
        seen, bound, misses, flips = lst.filter_selftest(x, r, gmin, gmax)
        assert misses == 0 and flips == 0, (seen, bound, misses, flips)
        assert 0.0 < seen <= bound
