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
