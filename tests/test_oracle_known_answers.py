"""Pin the CPU oracle against every literal known answer the reference's own unit tests
hold for the hot path.  CPU only.

Each test cites the reference test (relative to /root/reference/) it replays.
"""
import numpy as np
import pytest

from cabana_b200 import datasets


# --------------------------------------------------------------------------- CartesianGrid
def test_cartesian_grid_known_answer(orc):
    # core/unit_test/tstCartesianGrid.cpp:22-59
    g = orc.Grid((-1.0, -0.5, -0.6), (2.5, 1.5, 1.9), (0.5, 0.125, 0.25))
    assert g.nx == (7, 16, 10)
    assert g.total_cells == 7 * 16 * 10
    c = g.locate((-0.9, 1.4, 0.1))
    assert c == (0, 15, 2)
    assert g.min_distance((-0.9, 1.4, 0.1), c) == 0.0
    # upper-edge clamp
    assert g.locate((2.5, 1.5, 1.9)) == (6, 15, 9)


def test_cardinal_roundtrip(orc):
    g = orc.Grid((0, 0, 0), (7, 5, 3), (1, 1, 1))
    assert g.nx == (7, 5, 3)
    for c in range(g.total_cells):
        i, j, k = g.ijk(c)
        assert g.cardinal(i, j, k) == c
        assert c == (i * 5 + j) * 3 + k  # x slowest, z fastest


# --------------------------------------------------------------------------- stencil
def test_linked_cell_stencil_known_answer(orc):
    # core/unit_test/tstLinkedCellList.hpp:445-518: 10^3 unit grid, r = 1, ratio = 1
    st = orc.Stencil(1.0, 1.0, (0, 0, 0), (10, 10, 10))
    assert st.cell_range == 1
    assert st.nx == (10, 10, 10)
    g = orc.Grid((0, 0, 0), (10, 10, 10), (1, 1, 1))
    # interior point (4.5, 5.5, 3.5) -> [3,6) x [4,7) x [2,5)
    c = g.cardinal(*g.locate((4.5, 5.5, 3.5)))
    mn, mx = st.cells(c)
    assert mn == (3, 4, 2) and mx == (6, 7, 5)
    # lower corner clips to [0,2)
    mn, mx = st.cells(g.cardinal(*g.locate((0.5, 0.5, 0.5))))
    assert mn == (0, 0, 0) and mx == (2, 2, 2)
    # upper corner clips to [8,10)
    mn, mx = st.cells(g.cardinal(*g.locate((9.5, 9.5, 9.5))))
    assert mn == (8, 8, 8) and mx == (10, 10, 10)


def test_stencil_cell_range_ratio(orc):
    # Cabana_LinkedCellList.hpp:63  cell_range = ceil(1/ratio)
    assert orc.Stencil(1.0, 0.5, (0, 0, 0), (10, 10, 10)).cell_range == 2
    assert orc.Stencil(1.0, 0.25, (0, 0, 0), (10, 10, 10)).cell_range == 4
    assert orc.Stencil(1.0, 3.0, (0, 0, 0), (30, 30, 30)).cell_range == 1


# --------------------------------------------------------------------------- LinkedCellList
def _check_linked_cell(ps, res, begin, end, nx=10):
    # core/unit_test/tstLinkedCellList.hpp:281-365 (checkLinkedCell)
    particle_id = 0
    g = res.grid
    for i in range(nx):
        for j in range(nx):
            for k in range(nx):
                original = i + j * nx + k * nx * nx
                c = g.cardinal(i, j, k)
                if begin <= original < end:
                    assert res.counts[c] == 1
                    assert res.offsets[c] == particle_id
                    assert res.permute[particle_id] == original
                    particle_id += 1
                else:
                    assert res.counts[c] == 0
    assert particle_id == end - begin


def test_linked_cell_list_full_range(orc):
    # testLinkedList (tstLinkedCellList.hpp:584-620)
    ps = datasets.fixture_lcl_grid()
    x = orc.slice_from_xyz(ps.xyz, vlen=16, extra_doubles=2)
    res = orc.lcl_build(x, 0, ps.n, (1, 1, 1), ps.grid_min, ps.grid_max)
    _check_linked_cell(ps, res, 0, ps.n)
    # particle_bins: cell of every particle
    for p in (0, 1, 17, 999):
        i, j, k = p % 10, (p // 10) % 10, p // 100
        assert res.particle_bins[p] == res.grid.cardinal(i, j, k)
    # rebuild is idempotent
    res2 = orc.lcl_build(x, 0, ps.n, (1, 1, 1), ps.grid_min, ps.grid_max)
    assert np.array_equal(res.permute, res2.permute)


def test_linked_cell_list_partial_range(orc):
    # testLinkedListRange (tstLinkedCellList.hpp:622-660): [250,750)
    ps = datasets.fixture_lcl_grid()
    x = orc.view_from_xyz(ps.xyz)
    res = orc.lcl_build(x, 250, 750, (1, 1, 1), ps.grid_min, ps.grid_max)
    _check_linked_cell(ps, res, 250, 750)


def test_linked_cell_permute_slice(orc):
    # permute(lcl, slice): Cabana_Sort.hpp:600-656; after it positions are i-slowest/k-fastest
    ps = datasets.fixture_lcl_grid()
    x = orc.slice_from_xyz(ps.xyz, vlen=16)
    res = orc.lcl_build(x, 0, ps.n, (1, 1, 1), ps.grid_min, ps.grid_max)
    orc.permute_slice(x, 3, 0, ps.n, res.permute)
    xyz = x.to_xyz()
    sid = 0
    for i in range(10):
        for j in range(10):
            for k in range(10):
                assert tuple(xyz[sid]) == (i + 0.5, j + 0.5, k + 0.5)
                sid += 1
    # partial range leaves the rest untouched
    x2 = orc.slice_from_xyz(ps.xyz, vlen=16)
    res2 = orc.lcl_build(x2, 250, 750, (1, 1, 1), ps.grid_min, ps.grid_max)
    orc.permute_slice(x2, 3, 250, 750, res2.permute)
    xyz2 = x2.to_xyz()
    assert np.array_equal(xyz2[:250], ps.xyz[:250])
    assert np.array_equal(xyz2[750:], ps.xyz[750:])
    assert sorted(map(tuple, xyz2[250:750])) == sorted(map(tuple, ps.xyz[250:750]))


# --------------------------------------------------------------------------- Verlet vs N^2
def _assert_same_sets(a, b, begin=None, end=None):
    fa, sa = a.sorted_rows_flat()
    fb, sb = b.sorted_rows_flat()
    assert np.array_equal(a.counts, b.counts)
    assert np.array_equal(fa, fb)


@pytest.mark.parametrize("layout", ["csr", "2d"])
@pytest.mark.parametrize("positions", ["slice", "view"])
def test_verlet_full_matches_brute_force(orc, layout, positions):
    # testVerletListFull (tstNeighborList.hpp:27-79) + checkFullNeighborList
    ps = datasets.fixture_random300()
    x = orc.slice_from_xyz(ps.xyz) if positions == "slice" else orc.view_from_xyz(ps.xyz)
    lay = orc.CSR if layout == "csr" else orc.LAYOUT_2D
    n2 = orc.brute_force(x, ps.radius)
    vl = orc.verlet_build(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                          algo=orc.FULL, layout=lay)
    _assert_same_sets(vl, n2)
    assert vl.total == n2.total
    assert vl.max_n == n2.max_n


@pytest.mark.parametrize("max_neigh,expect_refill", [(100, False), (2, True)])
def test_verlet_2d_max_neigh_paths(orc, max_neigh, expect_refill):
    # tstNeighborList.hpp:58-77: max_neigh = 100 (no recount) and = 2 (realloc + refill)
    ps = datasets.fixture_random300()
    x = orc.slice_from_xyz(ps.xyz)
    n2 = orc.brute_force(x, ps.radius)
    vl = orc.verlet_build(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                          max_neigh=max_neigh, algo=orc.FULL, layout=orc.LAYOUT_2D)
    assert vl.refilled == expect_refill
    assert vl.width == (n2.max_n if expect_refill else max_neigh)
    assert vl.max_n == n2.max_n
    _assert_same_sets(vl, n2)


@pytest.mark.parametrize("layout", ["csr", "2d"])
def test_verlet_half_properties(orc, layout):
    # checkHalfNeighborList (neighbor_unit_test.hpp:201-243)
    ps = datasets.fixture_random300()
    x = orc.slice_from_xyz(ps.xyz)
    lay = orc.CSR if layout == "csr" else orc.LAYOUT_2D
    n2 = orc.brute_force(x, ps.radius)
    half = orc.verlet_build(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                            algo=orc.HALF, layout=lay)
    assert 2 * half.total == n2.total
    assert half.max_n <= n2.max_n
    assert half.total == n2.total // 2
    pairs = set()
    for i in range(ps.n):
        full_i = set(int(v) for v in n2.row(i))
        for j in half.row(i):
            j = int(j)
            assert j in full_i
            assert (j, i) not in pairs
            pairs.add((i, j))
    # every full pair appears in exactly one direction
    assert len(pairs) == n2.total // 2


@pytest.mark.parametrize("layout", ["csr", "2d"])
def test_verlet_full_partial_range(orc, layout):
    # checkFullNeighborListPartialRange (neighbor_unit_test.hpp:246-288): [75,225)
    ps = datasets.fixture_random300()
    x = orc.slice_from_xyz(ps.xyz)
    lay = orc.CSR if layout == "csr" else orc.LAYOUT_2D
    n2 = orc.brute_force(x, ps.radius)
    vl = orc.verlet_build(x, 75, 225, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                          algo=orc.FULL, layout=lay)
    for i in range(ps.n):
        if 75 <= i < 225:
            assert sorted(vl.row(i)) == sorted(n2.row(i))
        else:
            assert vl.counts[i] == 0


# --------------------------------------------------------------------------- literal known answers
def _kokkos_binop1d_hist(keys, nbin):
    """Kokkos::BinOp1D as used by Cabana::binByKey (Cabana_Sort.hpp:220-235; Kokkos 4.x
    Kokkos_BinOpsPublicAPI.hpp: mul = nbin/(max-min); bin = int(mul*(key-min)); nbin+1 bins)."""
    kmin, kmax = float(keys.min()), float(keys.max())
    mul = float(nbin) / (kmax - kmin)
    bins = (mul * (keys.astype(np.float64) - kmin)).astype(np.int64)
    return np.bincount(bins, minlength=nbin + 1)


def test_neighbor_histogram_known_answer(orc):
    # testNeighborHistogram (tstNeighborList.hpp:328-381): 10^3 lattice, r = 3dx + 1e-7
    ps = datasets.fixture_ordered(10)
    x = orc.slice_from_xyz(ps.xyz, extra_doubles=1)
    vl = orc.verlet_build(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                          algo=orc.FULL, layout=orc.CSR)
    assert vl.max_n == 122  # full spherical shell with cutoff/dx = 3
    for nbin, edges, expect in (
        (10, [12, 24, 36, 48, 61, 73, 85, 97, 109, 122], [32, 72, 24, 152, 120, 168, 0, 216, 0, 152]),
        (5, [24, 48, 73, 97, 122], [104, 176, 288, 216, 152]),
    ):
        width = vl.max_n / nbin  # Cabana_NeighborList.hpp:317-323
        assert [int((b + 1) * width) for b in range(nbin)] == edges
        hist = _kokkos_binop1d_hist(vl.counts, nbin)
        assert list(hist[:nbin]) == expect
    # and it is the N^2 list
    n2 = orc.brute_force(x, ps.radius)
    _assert_same_sets(vl, n2)


def test_tutorial_two_neighbors_each(orc):
    # example/core_tutorial/10_neighbor_parallel_for/neighbor_parallel_for_example.cpp:163
    ps = datasets.fixture_tutorial81()
    x = orc.view_from_xyz(ps.xyz)
    full = orc.verlet_build(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algo=orc.FULL)
    assert np.all(full.counts == 2)
    # coincident points are half-neighbours of neither
    half = orc.verlet_build(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algo=orc.HALF)
    assert np.all(half.counts == 0)


def test_closed_cutoff(orc):
    # dist_sqr <= rsqr (Cabana_VerletList.hpp:254)
    xyz = np.array([[1.0, 1.0, 1.0], [2.0, 1.0, 1.0], [1.0, 1.0, np.nextafter(2.0, 3.0)]])
    x = orc.view_from_xyz(xyz)
    vl = orc.verlet_build(x, 0, 3, 1.0, 1.0, (0, 0, 0), (4, 4, 4), algo=orc.FULL)
    assert list(vl.counts) == [1, 1, 0]


def test_neighbor_parallel_for_id_sum(orc):
    # checkFirstNeighborParallelFor (neighbor_unit_test.hpp:291-348): result[i] = sum of nbr ids
    ps = datasets.fixture_random300()
    x = orc.slice_from_xyz(ps.xyz)
    n2 = orc.brute_force(x, ps.radius)
    for lay in (orc.CSR, orc.LAYOUT_2D):
        vl = orc.verlet_build(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                              algo=orc.FULL, layout=lay)
        got = orc.neighbor_id_sum(lay, vl.counts, vl.offsets, vl.neighbors, vl.width, 0, ps.n)
        ref = np.array([int(n2.row(i).sum()) for i in range(ps.n)])
        assert np.array_equal(got, ref)


def test_lj_forces_newton_third_law(orc):
    # LJ consumer (north_star): full-list forces == half-list forces with f_j -= f
    ps = datasets.fixture_random300()
    x = orc.view_from_xyz(ps.xyz)
    full = orc.verlet_build(x, 0, ps.n, ps.radius, 0.5, ps.grid_min, ps.grid_max, algo=orc.FULL)
    half = orc.verlet_build(x, 0, ps.n, ps.radius, 0.5, ps.grid_min, ps.grid_max, algo=orc.HALF)
    f_full, fabs = orc.lj_forces(x, orc.CSR, full.counts, full.offsets, full.neighbors, 0, 0, ps.n,
                                 1.0, 1.0, 2.0)
    f_half, _ = orc.lj_forces(x, orc.CSR, half.counts, half.offsets, half.neighbors, 0, 0, ps.n,
                              1.0, 1.0, 2.0, newton=True)
    assert np.all(np.abs(f_full - f_half) <= 1e-12 * np.maximum(fabs, 1e-300))
    assert np.all(np.abs(f_full.sum(axis=0)) <= 1e-9 * fabs.sum())
    e_full = orc.lj_energy(x, orc.CSR, full.counts, full.offsets, full.neighbors, 0, 0, ps.n, 1, 1, 2.0, 0.5)
    e_half = orc.lj_energy(x, orc.CSR, half.counts, half.offsets, half.neighbors, 0, 0, ps.n, 1, 1, 2.0, 1.0)
    assert abs(e_full - e_half) <= 1e-12 * abs(e_full)


def test_fcc_interior_has_78_neighbors(orc):
    # r = 2.8 sigma sits between FCC shells 5 and 6 -> 12+6+24+12+24
    ps = datasets.fcc_lattice(8)
    x = orc.view_from_xyz(ps.xyz)
    vl = orc.verlet_build(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algo=orc.FULL)
    assert vl.max_n == 78
    n2 = orc.brute_force(x, ps.radius)
    _assert_same_sets(vl, n2)


def test_near_cutoff_adversarial_matches_structure(orc):
    # the oracle INCLUDES the reference's cell prune, so it may
    # differ from N^2 only by dropping pairs -- never by adding any.
    ps = datasets.near_cutoff_adversarial()
    x = orc.view_from_xyz(ps.xyz)
    n2 = orc.brute_force(x, ps.radius)
    vl = orc.verlet_build(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algo=orc.FULL)
    for i in range(ps.n):
        assert set(vl.row(i)) <= set(n2.row(i))
    # the set is actually adversarial: a good share of pairs sit within 4 ulp of r
    assert n2.total > 100


def test_non_uniform_radius_literal_answer(orc):
    """testNonUniformRadius (core/unit_test/tstNeighborList.hpp:210-253): 2^3 ordered particles,
    the first and last with radius 4.05 (reach everything but the body diagonal), the others
    3.32 (nearest neighbours only): 6 / 4 neighbours after the fill pass.  The count pass of the
    reference books the symmetric extra on the finder's row: 9 / 3,
    same total."""
    px, dx = 2, 2.5
    xyz = np.array([[dx / 2 + dx * i, dx / 2 + dx * j, dx / 2 + dx * k]
                    for i in range(px) for j in range(px) for k in range(px)])
    radii = np.full(8, 3.32)
    radii[0] = radii[7] = 4.05
    for layout in (orc.CSR, orc.LAYOUT_2D):
        res, cp = orc.verlet_build_radii(orc.view_from_xyz(xyz), radii, 0, 8, 3.32, 0.5,
                                         (0.0,) * 3, (5.0,) * 3, layout=layout)
        assert list(res.counts) == [6, 4, 4, 4, 4, 4, 4, 6]
        assert list(cp) == [9, 3, 3, 3, 3, 3, 3, 9]
        assert res.total == 36 and res.max_n == 6
        # particle 1 = (0,0,1): its three edge neighbours plus particle 7 (face diagonal, found
        # from 7's side only)
        assert sorted(int(v) for v in res.row(1)) == [0, 3, 5, 7]
    # equal radii reduce to the fixed-radius list
    res, _ = orc.verlet_build_radii(orc.view_from_xyz(xyz), np.full(8, 3.32), 0, 8, 3.32, 0.5,
                                    (0.0,) * 3, (5.0,) * 3)
    ref = orc.verlet_build(orc.view_from_xyz(xyz), 0, 8, 3.32, 0.5, (0.0,) * 3, (5.0,) * 3)
    assert np.array_equal(res.counts, ref.counts)
