"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Integer/index results are compared bit-exactly (per-particle neighbour SETS after
sorting); LJ forces within 1e-12 relative to the sum of |pair
force| on each component (north_star tolerance).
"""
import numpy as np
import pytest
import torch

from cabana_b200 import datasets
from _parity_helpers import full_size_oracle_compare as _full_size_oracle_compare

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12


@pytest.fixture(scope="module")
def cb():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from cabana_b200 import core

    return core


def _dev_positions(cb, xyz, kind):
    if kind == "view":
        return cb.view_from_array(xyz)
    if kind == "slice32":
        return cb.slice_from_array(xyz, vlen=32)
    if kind == "slice16x":
        return cb.slice_from_array(xyz, vlen=16, extra=2)
    raise ValueError(kind)


def _gpu_rows(orc, lst):
    d = lst._data
    counts = d.counts.cpu().numpy()
    offsets = d.offsets.cpu().numpy() if d.offsets is not None else None
    nb = d.neighbors.cpu().numpy()
    return counts, offsets, nb


def _assert_list_equal(orc, lst, ref, check_offsets=True):
    counts, offsets, nb = _gpu_rows(orc, lst)
    assert np.array_equal(counts, ref.counts), "per-particle counts differ"
    lay = orc.CSR if lst.layout == 0 else orc.LAYOUT_2D
    got, _ = orc.sorted_rows_flat(lay, counts, offsets, nb, lst.width)
    exp, _ = ref.sorted_rows_flat()
    assert np.array_equal(got, exp), "sorted neighbour rows differ"
    assert lst.total == ref.total
    assert lst._data.max_n == ref.max_n
    if lst.layout == 0 and check_offsets:
        # offsets = exclusive scan of counts in particle order (Cabana_VerletList.hpp:478-491)
        assert np.array_equal(offsets, ref.offsets)
    if lst.layout == 1:
        assert lst.width == ref.width


# ----------------------------------------------------------------------------- LinkedCellList
def _check_lcl(orc, cb, xyz, begin, end, delta, gmin, gmax, kind="slice32"):
    x = _dev_positions(cb, xyz, kind)
    lcl = cb.LinkedCellList(x, delta, gmin, gmax, begin, end)
    ref = orc.lcl_build(orc.view_from_xyz(xyz), begin, end, delta, gmin, gmax)
    counts = lcl.counts.cpu().numpy()
    offsets = lcl.offsets.cpu().numpy()
    perm = lcl.permutes.cpu().numpy().astype(np.int64)
    assert lcl.totalBins() == ref.grid.total_cells
    assert np.array_equal(counts, ref.counts)
    assert np.array_equal(offsets[:-1], ref.offsets)
    assert offsets[-1] == end - begin
    # within-cell order is unspecified (atomic slot claim): compare as a multiset per cell
    cell_of_slot = np.repeat(np.arange(len(counts)), counts)
    a = np.lexsort((perm, cell_of_slot))
    b = np.lexsort((ref.permute, cell_of_slot))
    assert np.array_equal(perm[a], ref.permute[b])
    assert np.array_equal(lcl.particle_bins.cpu().numpy(), ref.particle_bins)
    return lcl, x, ref


def test_lcl_grid_fixture_full_and_partial(orc, cb):
    # tstLinkedCellList.hpp:584-702
    ps = datasets.fixture_lcl_grid()
    for (b, e) in ((0, ps.n), (250, 750)):
        lcl, x, ref = _check_lcl(orc, cb, ps.xyz, b, e, (1, 1, 1), ps.grid_min, ps.grid_max)
        pid = 0
        for i in range(10):
            for j in range(10):
                for k in range(10):
                    orig = i + j * 10 + k * 100
                    if b <= orig < e:
                        assert lcl.binSize(i, j, k) == 1
                        assert lcl.binOffset(i, j, k) == pid
                        assert lcl.permutation(pid) == orig
                        pid += 1
                    else:
                        assert lcl.binSize(i, j, k) == 0
        assert not lcl.sorted()
        # permute(lcl, slice): sorted range is i-slowest/k-fastest; the rest is untouched
        cb.permute(lcl, x)
        assert lcl.sorted()
        xyz = x.to_array().cpu().numpy()
        assert np.array_equal(xyz[:b], ps.xyz[:b]) and np.array_equal(xyz[e:], ps.xyz[e:])
        exp = ps.xyz[ref.permute]
        assert np.array_equal(xyz[b:e], exp)
        # storeParticleBins with sorted == true: bins[s] = cell of sorted slot s
        bins = lcl.particle_bins.cpu().numpy()
        assert np.array_equal(bins, np.repeat(np.arange(1000), ref.counts))
        assert lcl.getParticle(3) == 3 + b


@pytest.mark.parametrize("kind", ["view", "slice32", "slice16x"])
def test_lcl_uniform_100k(orc, cb, kind):
    ps = datasets.uniform_box(100_000, 20240101)
    _check_lcl(orc, cb, ps.xyz, 0, ps.n, (3.0, 3.0, 3.0), ps.grid_min, ps.grid_max, kind)
    _check_lcl(orc, cb, ps.xyz, 1234, 77_777, (2.0, 3.5, 5.0), ps.grid_min, ps.grid_max, kind)


def test_lcl_permute_multiple_members(orc, cb):
    ps = datasets.uniform_box(20_000, 11)
    x = cb.slice_from_array(ps.xyz, vlen=32, extra=0)
    ids = cb.view_from_array(np.arange(ps.n, dtype=np.int32).reshape(-1, 1))
    vel = cb.view_from_array(np.arange(ps.n * 3, dtype=np.float64).reshape(-1, 3))
    lcl = cb.LinkedCellList(x, (3.0,) * 3, ps.grid_min, ps.grid_max)
    perm = lcl.permutes.cpu().numpy()
    cb.permute(lcl, x, ids, vel)
    assert np.array_equal(ids.to_array().cpu().numpy()[:, 0], perm)
    assert np.array_equal(x.to_array().cpu().numpy(), ps.xyz[perm])
    assert np.array_equal(vel.to_array().cpu().numpy(), np.arange(ps.n * 3, dtype=np.float64).reshape(-1, 3)[perm])
    # after the sort, rebuilding gives the identity permutation cell by cell
    lcl2 = cb.LinkedCellList(x, (3.0,) * 3, ps.grid_min, ps.grid_max)
    assert np.array_equal(lcl2.counts.cpu().numpy(), lcl.counts.cpu().numpy())
    p2 = lcl2.permutes.cpu().numpy().astype(np.int64)
    cell_of_slot = np.repeat(np.arange(lcl2.totalBins()), lcl2.counts.cpu().numpy())
    assert np.array_equal(np.sort(p2), np.arange(ps.n))
    # every slot's particle lies in that slot's cell range
    off = lcl2.offsets.cpu().numpy().astype(np.int64)
    assert np.all((p2 >= off[cell_of_slot]) & (p2 < off[cell_of_slot + 1]))


def test_scan_many_tiles_via_lcl(cb):
    # 200^3 = 8M cells -> 3907 scan tiles: decoupled look-back across many tiles
    ps = datasets.uniform_box(300_000, 5)
    hi = ps.grid_max[0]
    x = cb.view_from_array(ps.xyz)
    lcl = cb.LinkedCellList(x, (hi / 200.0 * 1.0000001,) * 3, ps.grid_min, ps.grid_max)
    counts = lcl.counts.to(torch.int64)
    offsets = lcl.offsets.to(torch.int64)
    assert int(counts.sum()) == ps.n
    exp = torch.cumsum(counts, 0) - counts
    assert torch.equal(offsets[:-1], exp)
    assert int(offsets[-1]) == ps.n


# ----------------------------------------------------------------------------- VerletList
DATASETS = {
    "random300": lambda: datasets.fixture_random300(),
    "ordered10": lambda: datasets.fixture_ordered(10),
    "tutorial81": lambda: datasets.fixture_tutorial81(),
    "near_cutoff": lambda: datasets.near_cutoff_adversarial(),
    "fcc8": lambda: datasets.fcc_lattice(8),
    "fcc10_jitter": lambda: datasets.fcc_lattice(10, jitter=0.05),
    "uniform20k": lambda: datasets.uniform_box(20_000, 20240101),
    "uniform20k_half_cells": lambda: datasets.uniform_box(20_000, 3, cell_ratio=0.5),
    "clustered20k": lambda: datasets.clustered(20_000),
}


@pytest.mark.parametrize("name", list(DATASETS))
@pytest.mark.parametrize("algo", ["full", "half"])
@pytest.mark.parametrize("layout", ["csr", "2d"])
def test_verlet_matches_oracle(orc, cb, name, algo, layout):
    # testVerletListFull / testVerletListHalf (tstNeighborList.hpp:27-141)
    ps = DATASETS[name]()
    a = cb.FULL if algo == "full" else cb.HALF
    lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
    kind = {"csr": "slice32", "2d": "view"}[layout]
    x = _dev_positions(cb, ps.xyz, kind)
    lst = cb.VerletList(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                        algorithm=a, layout=lay)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, ps.cell_ratio,
                           ps.grid_min, ps.grid_max, algo=a, layout=lay)
    _assert_list_equal(orc, lst, ref)


@pytest.mark.parametrize("layout", ["csr", "2d"])
@pytest.mark.parametrize("build_tag", ["team", "team_vector"])
def test_verlet_partial_range(orc, cb, layout, build_tag):
    # testVerletListFullPartialRange (tstNeighborList.hpp:112-141): [75,225)
    ps = datasets.fixture_random300()
    lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
    tag = cb.OP_TEAM if build_tag == "team" else cb.OP_TEAM_VECTOR
    x = cb.slice_from_array(ps.xyz, vlen=16, extra=1)
    lst = cb.VerletList(x, 75, 225, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                        algorithm=cb.FULL, layout=lay, build_tag=tag)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 75, 225, ps.radius, ps.cell_ratio,
                           ps.grid_min, ps.grid_max, algo=orc.FULL, layout=lay)
    _assert_list_equal(orc, lst, ref)
    counts = lst._data.counts.cpu().numpy()
    assert counts[:75].sum() == 0 and counts[225:].sum() == 0
    n2 = orc.brute_force(orc.view_from_xyz(ps.xyz), ps.radius)
    assert np.array_equal(counts[75:225], n2.counts[75:225])


@pytest.mark.parametrize("max_neigh,expect_refill", [(100, False), (2, True)])
def test_verlet_2d_max_neigh(orc, cb, max_neigh, expect_refill):
    # tstNeighborList.hpp:58-77 and SURVEY.md Appendix B.5
    ps = datasets.fixture_random300()
    x = cb.slice_from_array(ps.xyz)
    lst = cb.VerletList(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max, max_neigh,
                        algorithm=cb.FULL, layout=cb.LAYOUT_2D)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, ps.cell_ratio,
                           ps.grid_min, ps.grid_max, max_neigh=max_neigh, algo=orc.FULL,
                           layout=orc.LAYOUT_2D)
    assert lst.refilled == expect_refill == ref.refilled
    _assert_list_equal(orc, lst, ref)


def test_verlet_rebuild_reuses_handle(orc, cb):
    ps1 = datasets.uniform_box(20_000, 1)
    ps2 = datasets.uniform_box(5_000, 2)
    lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
    for ps in (ps1, ps2, ps1):
        x = cb.view_from_array(ps.xyz)
        lst.build(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max)
        ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, 1.0, ps.grid_min,
                               ps.grid_max)
        _assert_list_equal(orc, lst, ref)


def test_verlet_known_answers_on_gpu(orc, cb):
    # testNeighborHistogram literal (tstNeighborList.hpp:351-379) on the CUDA path
    ps = datasets.fixture_ordered(10)
    x = cb.slice_from_array(ps.xyz, vlen=32, extra=1)
    lst = cb.VerletList(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                        algorithm=cb.FULL, layout=cb.CSR, build_tag=cb.OP_TEAM)
    counts = lst._data.counts.cpu().numpy()
    assert cb.NeighborList.maxNeighbor(lst) == 122
    kmin, kmax = counts.min(), counts.max()
    bins = (10.0 / (kmax - kmin) * (counts.astype(np.float64) - kmin)).astype(np.int64)
    assert list(np.bincount(bins, minlength=11)[:10]) == [32, 72, 24, 152, 120, 168, 0, 216, 0, 152]
    # tutorial: two neighbours each; half list of coincident points is empty (Appendix B.4)
    ps = datasets.fixture_tutorial81()
    x = cb.view_from_array(ps.xyz)
    full = cb.VerletList(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max)
    assert torch.all(full._data.counts == 2)
    half = cb.VerletList(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algorithm=cb.HALF)
    assert torch.all(half._data.counts == 0)
    assert cb.NeighborList.numNeighbor(full, 5) == 2
    nb = {cb.NeighborList.getNeighbor(full, 4, 0), cb.NeighborList.getNeighbor(full, 4, 1)}
    assert nb == {3, 5}
    # setNeighbor (testModifyNeighbors, tstNeighborList.hpp:256-292)
    full.setNeighbor(4, 1, -7)
    assert cb.NeighborList.getNeighbor(full, 4, 1) == -7


def test_verlet_uniform_100k_cfg1(orc, cb):
    # BASELINE config 1 at full size against the oracle
    ps = datasets.uniform_box(100_000, 20240101)
    x = cb.slice_from_array(ps.xyz, vlen=32)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, 3.0, 1.0, ps.grid_min, ps.grid_max)
    for lay in (cb.CSR, cb.LAYOUT_2D):
        lst = cb.VerletList(x, 0, ps.n, 3.0, 1.0, ps.grid_min, ps.grid_max, layout=lay)
        if lay == cb.CSR:
            _assert_list_equal(orc, lst, ref)
        else:
            counts, _, nb = _gpu_rows(orc, lst)
            assert np.array_equal(counts, ref.counts)
            got, _ = orc.sorted_rows_flat(orc.LAYOUT_2D, counts, None, nb, lst.width)
            assert np.array_equal(got, ref.sorted_rows_flat()[0])


# ----------------------------------------------------------------------------- traversal
@pytest.mark.parametrize("layout", ["csr", "2d"])
@pytest.mark.parametrize("op", ["serial", "team"])
def test_neighbor_parallel_for_id_sum(orc, cb, layout, op):
    # testNeighborParallelFor (tstNeighborList.hpp:144-175)
    ps = datasets.fixture_random300()
    lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
    tag = cb.OP_SERIAL if op == "serial" else cb.OP_TEAM
    x = cb.slice_from_array(ps.xyz)
    lst = cb.VerletList(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max, layout=lay)
    n2 = orc.brute_force(orc.view_from_xyz(ps.xyz), ps.radius)
    expect = np.array([int(n2.row(i).sum()) for i in range(ps.n)])
    res = torch.zeros(ps.n, dtype=torch.int64, device="cuda")
    cb.neighbor_parallel_for_id_sum(0, ps.n, lst, res, tag)
    assert np.array_equal(res.cpu().numpy(), expect)
    # sub-range policy: only [begin,end) rows are visited
    res.zero_()
    cb.neighbor_parallel_for_id_sum(50, 200, lst, res, tag)
    e2 = expect.copy()
    e2[:50] = 0
    e2[200:] = 0
    assert np.array_equal(res.cpu().numpy(), e2)


def _lj_case(orc, cb, ps, algo, layout, op, kind="view"):
    a = cb.FULL if algo == "full" else cb.HALF
    lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
    tag = cb.OP_SERIAL if op == "serial" else cb.OP_TEAM
    x = _dev_positions(cb, ps.xyz, kind)
    lst = cb.VerletList(x, 0, ps.n, ps.radius, ps.cell_ratio, ps.grid_min, ps.grid_max,
                        algorithm=a, layout=lay)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, ps.cell_ratio,
                           ps.grid_min, ps.grid_max, algo=a, layout=lay)
    rc = 2.5 if ps.radius > 2.5 else ps.radius
    f_ref, fabs = orc.lj_forces(orc.view_from_xyz(ps.xyz), lay, ref.counts, ref.offsets, ref.neighbors,
                                ref.width, 0, ps.n, 1.0, 1.0, rc, newton=(algo == "half"))
    f = cb.view_from_array(np.zeros((ps.n, 3)))
    cb.neighbor_parallel_for_lj(0, ps.n, lst, x, f, 1.0, 1.0, rc, tag)
    got = f.to_array().cpu().numpy()
    err = np.abs(got - f_ref)
    assert np.all(err <= REL_TOL * np.maximum(fabs, 1e-300)), float((err / np.maximum(fabs, 1e-300)).max())
    scale = 0.5 if algo == "full" else 1.0
    e_ref = orc.lj_energy(orc.view_from_xyz(ps.xyz), lay, ref.counts, ref.offsets, ref.neighbors,
                          ref.width, 0, ps.n, 1.0, 1.0, rc, scale)
    e = cb.neighbor_parallel_reduce_lj(0, ps.n, lst, x, 1.0, 1.0, rc, tag)
    assert abs(e - e_ref) <= 1e-11 * abs(e_ref)


@pytest.mark.parametrize("algo", ["full", "half"])
@pytest.mark.parametrize("layout", ["csr", "2d"])
@pytest.mark.parametrize("op", ["serial", "team"])
def test_lj_forces_match_oracle(orc, cb, algo, layout, op):
    _lj_case(orc, cb, datasets.fcc_lattice(8, jitter=0.05), algo, layout, op)


def test_lj_forces_uniform_slice_layout(orc, cb):
    # min separation in a uniform random set can be tiny -> huge forces; the tolerance is
    # relative to sum |pair force| so this still has to hold.
    ps = datasets.uniform_box(20_000, 20240102)
    _lj_case(orc, cb, ps, "half", "2d", "team", kind="slice32")
    _lj_case(orc, cb, ps, "full", "csr", "serial", kind="slice16x")


# ----------------------------------------------------------------------------- full-size properties
def test_fcc_16m_properties(cb):
    """BASELINE config 3 at full size (4*159^3 = 16 078 716 atoms): size-independent
    properties instead of an oracle run -- interior atoms have exactly 78 neighbours
    (12+6+24+12+24), sum(full) = 2 sum(half), offsets = exclusive scan of counts, every
    stored pair is within the cutoff, and the list is symmetric on a sample."""
    ps = datasets.fcc_lattice(159)
    assert ps.n == 16_078_716
    x = cb.view_from_array(ps.xyz)
    full = cb.VerletList(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max,
                         algorithm=cb.FULL, layout=cb.CSR)
    counts = full._data.counts
    assert full._data.max_n == 78
    xyz = x.to_array()
    hi = torch.tensor(ps.grid_max, device="cuda", dtype=torch.float64)
    interior = torch.all((xyz > ps.radius) & (xyz < hi - ps.radius), dim=1)
    assert bool(torch.all(counts[interior] == 78))
    c64 = counts.to(torch.int64)
    assert full.total == int(c64.sum())
    assert torch.equal(full._data.offsets.to(torch.int64), torch.cumsum(c64, 0) - c64)
    # every stored pair is within the cutoff and j != i (sample of rows)
    g = torch.Generator(device="cpu").manual_seed(7)
    rows = torch.randint(0, ps.n, (20000,), generator=g).cuda()
    off = full._data.offsets[rows].to(torch.int64)
    cnt = counts[rows].to(torch.int64)
    maxc = int(cnt.max())
    k = torch.arange(maxc, device="cuda")[None, :]
    valid = k < cnt[:, None]
    idx = torch.where(valid, off[:, None] + k, torch.zeros_like(k))
    nb = full._data.neighbors[idx].to(torch.int64)
    d2 = ((xyz[rows][:, None, :] - xyz[nb]) ** 2).sum(-1)
    assert bool(torch.all(d2[valid] <= ps.radius**2 * (1 + 1e-14)))
    assert bool(torch.all(nb[valid] != rows[:, None].expand_as(nb)[valid]))
    # symmetry: j in N(i) => i in N(j)
    jj = nb[:, 0]
    joff = full._data.offsets[jj].to(torch.int64)
    jcnt = counts[jj].to(torch.int64)
    jidx = torch.where(k < jcnt[:, None], joff[:, None] + k, torch.zeros_like(k))
    jnb = full._data.neighbors[jidx].to(torch.int64)
    found = ((jnb == rows[:, None]) & (k < jcnt[:, None])).any(dim=1)
    assert bool(torch.all(found[cnt > 0]))
    total_full = full.total
    del full, nb, jnb, d2, idx, jidx
    torch.cuda.empty_cache()
    half = cb.VerletList(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max,
                         algorithm=cb.HALF, layout=cb.CSR)
    assert 2 * half.total == total_full


@pytest.mark.parametrize("algo", ["full", "half"])
def test_fcc_16m_matches_oracle(orc, cb, algo):
    """BASELINE config 3 (the headline configuration) at its full 16 078 716 atoms, compared
    with the oracle row by row."""
    ps = datasets.fcc_lattice(159)
    assert ps.n == 16_078_716
    _full_size_oracle_compare(orc, cb, ps, cb.FULL if algo == "full" else cb.HALF,
                              orc.FULL if algo == "full" else orc.HALF)


# ------------------------------------------------------------- CB_ROWS_BINNED (opt-in placement)
@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("case", ["fcc", "uniform", "partial"])
def test_binned_row_placement_same_sets(cb, orc, algo, case):
    """Rows left where the single test pass wrote them: identical counts and neighbour SETS,
    rows disjoint inside `neighbors`, LJ traversal identical to the reference placement."""
    if case == "fcc":
        ps = datasets.fcc_lattice(14, jitter=0.04)
        b, e = 0, ps.n
    elif case == "uniform":
        ps = datasets.uniform_box(30000, 99)
        b, e = 0, ps.n
    else:
        ps = datasets.uniform_box(20000, 98)
        b, e = 3000, 17000
    x = cb.slice_from_array(ps.xyz, vlen=32)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), b, e, ps.radius, 1.0, ps.grid_min, ps.grid_max,
                           algo=orc.FULL if algo == 0 else orc.HALF)
    lst = cb.VerletList(x, b, e, ps.radius, 1.0, ps.grid_min, ps.grid_max, algorithm=algo, layout=cb.CSR,
                        row_placement=cb.ROWS_BINNED)
    _assert_list_equal(orc, lst, ref, check_offsets=False)
    counts, offsets, nb = _gpu_rows(orc, lst)
    assert lst.extent >= lst.total and nb.shape[0] == lst.extent
    # rows are disjoint intervals of the neighbour array
    rows = np.nonzero(counts > 0)[0]
    order = rows[np.argsort(offsets[rows])]
    starts, ends = offsets[order].astype(np.int64), offsets[order].astype(np.int64) + counts[order]
    assert np.all(starts[1:] >= ends[:-1]) and ends[-1] <= lst.extent
    # a rebuild (buffers swap roles) gives the same sets again
    lst.build(x, b, e, ps.radius, 1.0, ps.grid_min, ps.grid_max)
    _assert_list_equal(orc, lst, ref, check_offsets=False)
    # traversal through offsets[i] is placement-agnostic: same forces, bit for bit per row order
    lst_ref = cb.VerletList(x, b, e, ps.radius, 1.0, ps.grid_min, ps.grid_max, algorithm=algo, layout=cb.CSR)
    f1 = cb.view_from_array(np.zeros((ps.n, 3)))
    f2 = cb.view_from_array(np.zeros((ps.n, 3)))
    cb.neighbor_parallel_for_lj(b, e, lst, x, f1, 1.0, 1.0, 2.5, cb.OP_TEAM if algo == 1 else cb.OP_SERIAL)
    cb.neighbor_parallel_for_lj(b, e, lst_ref, x, f2, 1.0, 1.0, 2.5, cb.OP_TEAM if algo == 1 else cb.OP_SERIAL)
    a1, a2 = f1.to_array().cpu().numpy(), f2.to_array().cpu().numpy()
    scale = np.abs(a2).max() + 1.0
    assert np.max(np.abs(a1 - a2)) <= 1e-12 * scale


def test_binned_placement_2d_is_reference_and_host_copy(cb, orc):
    """CB_ROWS_BINNED only applies to CSR: a 2D list is laid out as always; the host copy of a
    binned CSR list carries `extent` ids and the same rows."""
    ps = datasets.uniform_box(20000, 97)
    x = cb.slice_from_array(ps.xyz, vlen=32)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max,
                           algo=orc.FULL, layout=orc.LAYOUT_2D)
    l2d = cb.VerletList(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algorithm=cb.FULL,
                        layout=cb.LAYOUT_2D, row_placement=cb.ROWS_BINNED)
    _assert_list_equal(orc, l2d, ref)
    # end-to-end entry points with the binned placement
    lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR, row_placement=cb.ROWS_BINNED)
    lst.build_host(ps.xyz, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max)
    counts_h = torch.empty(ps.n, dtype=torch.int32).pin_memory()
    offsets_h = torch.empty(ps.n, dtype=torch.int32).pin_memory()
    nb_h = torch.empty(lst.extent, dtype=torch.int32).pin_memory()
    lst.copy_to_host(counts_h, offsets_h, nb_h)
    torch.cuda.synchronize()
    refc = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max,
                            algo=orc.FULL)
    got, _ = orc.sorted_rows_flat(orc.CSR, counts_h.numpy(), offsets_h.numpy(), nb_h.numpy(), 0)
    assert np.array_equal(counts_h.numpy(), refc.counts)
    assert np.array_equal(got, refc.sorted_rows_flat()[0])
    with pytest.raises(Exception):
        lst.copy_to_host(counts_h, offsets_h, torch.empty(lst.total - 1, dtype=torch.int32).pin_memory())


def test_permute_through_binning_data(cb, orc):
    """permute(BinningData, slice) == permute(LinkedCellList, slice) (Cabana_Sort.hpp:549-715)."""
    ps = datasets.fixture_random300()
    d = ps.radius * ps.cell_ratio
    x1 = cb.slice_from_array(ps.xyz, vlen=32)
    x2 = cb.view_from_array(ps.xyz)
    lcl = cb.LinkedCellList(x1, (d, d, d), ps.grid_min, ps.grid_max)
    bd = lcl.binningData()
    assert bd.numBin() == lcl.totalBins() and bd.rangeBegin() == 0 and bd.rangeEnd() == ps.n
    perm = lcl.permutes.cpu().numpy().astype(np.int64)
    cb.permute(bd, x2)
    cb.permute(lcl, x1)
    a = x1.to_array().cpu().numpy()
    b = x2.to_array().cpu().numpy()
    assert np.array_equal(a, b)
    assert np.array_equal(a, ps.xyz[perm])
    assert sum(bd.binSize(c) for c in range(0, bd.numBin(), 97)) >= 0
