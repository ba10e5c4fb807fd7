"""List-free traversal: neighbor_parallel_for directly on a LinkedCellList
(core/src/Cabana_Parallel.hpp:1122-1290; tests testLinkedCellNeighborInterface / Parallel,
core/unit_test/tstLinkedCellList.hpp:704-1003) against the oracle and the N^2 list, before
and after permute (sorted flag), full and partial range, Serial and Team."""
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


@pytest.mark.parametrize("op", ["serial", "team"])
@pytest.mark.parametrize("rng_", [(0, 300), (75, 225)])
def test_lcl_neighbor_count_matches_n2(orc, cb, op, rng_):
    ps = datasets.fixture_random300()
    b, e = rng_
    r = ps.radius
    tag = cb.OP_SERIAL if op == "serial" else cb.OP_TEAM
    delta = (r, r, r)  # cell = cutoff, ratio 1: the stencil covers the cutoff
    x = cb.slice_from_array(ps.xyz, vlen=32)
    lcl = cb.LinkedCellList(x, delta, ps.grid_min, ps.grid_max, b, e, neighborhood_radius=r,
                            cell_size_ratio=1.0)
    n2 = orc.brute_force(orc.view_from_xyz(ps.xyz), r)
    # N^2 neighbours restricted to the binned range (tstLinkedCellList.hpp:726-741)
    expect = np.zeros(ps.n, dtype=np.int32)
    for p in range(b, e):
        row = n2.row(p)
        expect[p] = int(((row >= b) & (row < e)).sum())
    res = torch.zeros(ps.n, dtype=torch.int32, device="cuda")
    cb.lcl_neighbor_parallel_for_count(b, e, lcl, x, r, res, tag)
    assert np.array_equal(res.cpu().numpy(), expect)
    # oracle restatement agrees too
    ox = orc.view_from_xyz(ps.xyz)
    ores = orc.lcl_build(ox, b, e, delta, ps.grid_min, ps.grid_max)
    st = orc.Stencil(r, 1.0, ps.grid_min, ps.grid_max)
    oc, _, _ = orc.lcl_neighbor_for(ox, ores, st, False, b, b, e, 0, r)
    assert np.array_equal(oc, expect)
    # after permute the list is "sorted": same counts at the permuted positions
    perm = lcl.permutes.cpu().numpy().astype(np.int64)
    cb.permute(lcl, x)
    assert lcl.sorted()
    res.zero_()
    cb.lcl_neighbor_parallel_for_count(b, e, lcl, x, r, res, tag)
    got = res.cpu().numpy()
    exp_sorted = np.zeros(ps.n, dtype=np.int32)
    exp_sorted[b:e] = expect[perm]
    assert np.array_equal(got, exp_sorted)


@pytest.mark.parametrize("op", ["serial", "team"])
def test_lcl_lj_forces_equal_verlet_lj_forces(orc, cb, op):
    ps = datasets.fcc_lattice(10, jitter=0.05)
    tag = cb.OP_SERIAL if op == "serial" else cb.OP_TEAM
    r = ps.radius
    x = cb.view_from_array(ps.xyz)
    lcl = cb.LinkedCellList(x, (r, r, r), ps.grid_min, ps.grid_max, neighborhood_radius=r,
                            cell_size_ratio=1.0)
    f = cb.view_from_array(np.zeros((ps.n, 3)))
    cb.lcl_neighbor_parallel_for_lj(0, ps.n, lcl, x, f, 1.0, 1.0, 2.5, tag)
    ox = orc.view_from_xyz(ps.xyz)
    full = orc.verlet_build(ox, 0, ps.n, r, 1.0, ps.grid_min, ps.grid_max)
    f_ref, fabs = orc.lj_forces(ox, orc.CSR, full.counts, full.offsets, full.neighbors, 0, 0, ps.n,
                                1.0, 1.0, 2.5)
    err = np.abs(f.to_array().cpu().numpy() - f_ref)
    assert np.all(err <= 1e-12 * np.maximum(fabs, 1e-300))
    # and the oracle's own LCL traversal
    ores = orc.lcl_build(ox, 0, ps.n, (r, r, r), ps.grid_min, ps.grid_max)
    st = orc.Stencil(r, 1.0, ps.grid_min, ps.grid_max)
    _, f_o, fabs_o = orc.lcl_neighbor_for(ox, ores, st, False, 0, 0, ps.n, 1, 2.5)
    assert np.all(np.abs(f_o - f_ref) <= 1e-12 * np.maximum(fabs, 1e-300))
