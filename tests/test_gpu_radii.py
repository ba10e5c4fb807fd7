"""Per-particle cutoff radius build (core/src/Cabana_VerletList.hpp:181-203,
:244-305, :989-1017) on the CUDA path against the oracle and the reference's literal answer
(testNonUniformRadius, core/unit_test/tstNeighborList.hpp:210-253: 6 neighbours for the two
large-radius particles, 4 for the others)."""
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


def _ordered(px):
    dx = 5.0 / px
    return np.array([[dx / 2 + dx * i, dx / 2 + dx * j, dx / 2 + dx * k]
                     for i in range(px) for j in range(px) for k in range(px)])


def _compare(orc, cb, xyz, radii, begin, end, bg, ratio, gmin, gmax, algo, layout, max_neigh=0, kind="slice"):
    n = xyz.shape[0]
    x = cb.slice_from_array(xyz, vlen=32) if kind == "slice" else cb.view_from_array(xyz)
    r = (cb.slice_from_array(radii.reshape(-1, 1), vlen=32) if kind == "slice"
         else cb.view_from_array(radii.reshape(-1, 1)))
    lst = cb.VerletList(algorithm=algo, layout=layout)
    lst.build_radii(x, begin, end, bg, r, ratio, gmin, gmax, max_neigh)
    ref, _ = orc.verlet_build_radii(orc.view_from_xyz(xyz), radii, begin, end, bg, ratio, gmin, gmax,
                                    max_neigh=max_neigh, algo=algo, layout=layout)
    counts = lst._data.counts.cpu().numpy()
    assert np.array_equal(counts, ref.counts)
    offsets = lst._data.offsets.cpu().numpy() if layout == cb.CSR else None
    nb = lst._data.neighbors.cpu().numpy()
    got, _ = orc.sorted_rows_flat(layout, counts, offsets, nb, lst.width)
    assert np.array_equal(got, ref.sorted_rows_flat()[0])
    assert lst.total == ref.total and lst._data.max_n == ref.max_n
    if layout == cb.CSR:
        assert np.array_equal(offsets, ref.offsets)
    else:
        assert lst.width == ref.width and lst.refilled == ref.refilled
    return counts


@pytest.mark.parametrize("layout", ["csr", "2d"])
def test_non_uniform_radius_literal_answer(orc, cb, layout):
    xyz = _ordered(2)
    radii = np.full(8, 3.32)
    radii[0] = radii[7] = 4.05
    lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
    counts = _compare(orc, cb, xyz, radii, 0, 8, 3.32, 0.5, (0.0,) * 3, (5.0,) * 3, cb.FULL, lay)
    assert list(counts) == [6, 4, 4, 4, 4, 4, 4, 6]
    # the reference's count pass books the symmetric extra on the finder's row (Appendix B.6):
    # totals agree, rows do not -- the library follows the fill pass
    _, cp = orc.verlet_build_radii(orc.view_from_xyz(xyz), radii, 0, 8, 3.32, 0.5, (0.0,) * 3, (5.0,) * 3)
    assert list(cp) == [9, 3, 3, 3, 3, 3, 3, 9] and cp.sum() == counts.sum()


@pytest.mark.parametrize("algo", ["full", "half"])
@pytest.mark.parametrize("layout", ["csr", "2d"])
def test_random_radii_match_oracle(orc, cb, algo, layout):
    ps = datasets.uniform_box(8000, 31, radius=3.0)
    rng = np.random.default_rng(5)
    radii = rng.uniform(1.0, 3.0, ps.n)       # every radius <= the background radius
    a = cb.FULL if algo == "full" else cb.HALF
    lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
    _compare(orc, cb, ps.xyz, radii, 0, ps.n, 3.0, 1.0, ps.grid_min, ps.grid_max, a, lay)
    _compare(orc, cb, ps.xyz, radii, 1000, 6000, 3.0, 0.5, ps.grid_min, ps.grid_max, a, lay, kind="view")


def test_uniform_radii_equal_fixed_radius_list(orc, cb):
    # all radii equal: no pair is "not found from the other side" except at exact equality
    ps = datasets.fcc_lattice(8, jitter=0.05)
    radii = np.full(ps.n, ps.radius)
    counts = _compare(orc, cb, ps.xyz, radii, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max,
                      cb.FULL, cb.CSR)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max)
    assert np.array_equal(counts, ref.counts)


@pytest.mark.parametrize("max_neigh,refill", [(200, False), (3, True)])
def test_radii_2d_max_neigh(orc, cb, max_neigh, refill):
    ps = datasets.fixture_random300()
    radii = np.random.default_rng(9).uniform(0.8, ps.radius, ps.n)
    x = cb.slice_from_array(ps.xyz)
    lst = cb.VerletList(algorithm=cb.FULL, layout=cb.LAYOUT_2D)
    lst.build_radii(x, 0, ps.n, ps.radius, cb.view_from_array(radii.reshape(-1, 1)), ps.cell_ratio,
                    ps.grid_min, ps.grid_max, max_neigh)
    assert lst.refilled == refill
    _compare(orc, cb, ps.xyz, radii, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
             cb.FULL, cb.LAYOUT_2D, max_neigh=max_neigh)


def test_radii_argument_checks(cb):
    ps = datasets.fixture_random300()
    x = cb.slice_from_array(ps.xyz)
    lst = cb.VerletList()
    from cabana_b200.capi import CabanaB200Error

    with pytest.raises(CabanaB200Error):   # size( positions ) == size( radius ) (:194)
        lst.build_radii(x, 0, ps.n, 1.0, cb.view_from_array(np.ones((ps.n - 1, 1))), 1.0,
                        ps.grid_min, ps.grid_max)
    with pytest.raises(CabanaB200Error):
        lst.build_radii(x, 0, ps.n, 1.0, cb.view_from_array(np.ones((ps.n, 1), dtype=np.float32)), 1.0,
                        ps.grid_min, ps.grid_max)
