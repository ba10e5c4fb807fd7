"""The BASELINE.json configurations that are not the bench line, as parity cases:
cfg2 (1M uniform, LCL-sorted, HalfNeighborTag + VerletLayout2D + LJ Serial/Team) and
cfg4 (clustered, 10x density contrast, Full CSR at cell_size_ratio 1.0 and 0.5) against the
oracle at 1M particles, and cfg4 at its full 8M size through size-independent properties."""
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


def _rows_equal(orc, lst, ref):
    counts = lst._data.counts.cpu().numpy()
    assert np.array_equal(counts, ref.counts)
    offsets = lst._data.offsets.cpu().numpy() if lst._data.offsets is not None else None
    nb = lst._data.neighbors.cpu().numpy()
    got, _ = orc.sorted_rows_flat(orc.CSR if lst.layout == 0 else orc.LAYOUT_2D, counts, offsets, nb, lst.width)
    assert np.array_equal(got, ref.sorted_rows_flat()[0])


def test_cfg2_1m_uniform_half_2d_lj(orc, cb):
    ps = datasets.uniform_box(1_000_000, 20240102)
    # the reference benchmark sorts first with an LCL of cell = cutoff
    # (benchmark/core/Cabana_NeighborVerletPerformance.cpp:80-91)
    x = cb.slice_from_array(ps.xyz, vlen=32)
    lcl = cb.LinkedCellList(x, (3.0,) * 3, ps.grid_min, ps.grid_max)
    perm = lcl.permutes.cpu().numpy().astype(np.int64)
    cb.permute(lcl, x)
    xyz_sorted = ps.xyz[perm]
    assert np.array_equal(x.to_array().cpu().numpy(), xyz_sorted)
    lst = cb.VerletList(x, 0, ps.n, 3.0, 1.0, ps.grid_min, ps.grid_max, algorithm=cb.HALF,
                        layout=cb.LAYOUT_2D)
    ox = orc.view_from_xyz(xyz_sorted)
    ref = orc.verlet_build(ox, 0, ps.n, 3.0, 1.0, ps.grid_min, ps.grid_max, algo=orc.HALF,
                           layout=orc.LAYOUT_2D)
    _rows_equal(orc, lst, ref)
    assert lst._data.max_n == ref.max_n and lst.width == ref.width
    f_ref, fabs = orc.lj_forces(ox, orc.LAYOUT_2D, ref.counts, None, ref.neighbors, ref.width, 0, ps.n,
                                1.0, 1.0, 2.5, newton=True)
    for op in (cb.OP_SERIAL, cb.OP_TEAM):
        f = cb.view_from_array(np.zeros((ps.n, 3)))
        cb.neighbor_parallel_for_lj(0, ps.n, lst, x, f, 1.0, 1.0, 2.5, op)
        err = np.abs(f.to_array().cpu().numpy() - f_ref)
        assert np.all(err <= 1e-12 * np.maximum(fabs, 1e-300))


@pytest.mark.parametrize("ratio", [1.0, 0.5])
def test_cfg4_1m_clustered_full_csr(orc, cb, ratio):
    ps = datasets.clustered(1_000_000, cell_ratio=ratio)
    x = cb.view_from_array(ps.xyz)
    lst = cb.VerletList(x, 0, ps.n, ps.radius, ratio, ps.grid_min, ps.grid_max)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, ratio, ps.grid_min, ps.grid_max)
    _rows_equal(orc, lst, ref)
    assert ref.max_n > 3 * ref.total / ps.n  # the set really is clustered


def test_cfg4_8m_clustered_properties(cb):
    ps = datasets.clustered(8_000_000)
    x = cb.view_from_array(ps.xyz)
    full = cb.VerletList(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max)
    counts = full._data.counts
    c64 = counts.to(torch.int64)
    assert full.total == int(c64.sum())
    assert torch.equal(full._data.offsets.to(torch.int64), torch.cumsum(c64, 0) - c64)
    xyz = x.to_array()
    g = torch.Generator(device="cpu").manual_seed(11)
    rows = torch.randint(0, ps.n, (4000,), generator=g).cuda()
    off = full._data.offsets[rows].to(torch.int64)
    cnt = counts[rows].to(torch.int64)
    k = torch.arange(int(cnt.max()), device="cuda")[None, :]
    valid = k < cnt[:, None]
    nb = full._data.neighbors[torch.where(valid, off[:, None] + k, torch.zeros_like(k))].to(torch.int64)
    d2 = ((xyz[rows][:, None, :] - xyz[nb]) ** 2).sum(-1)
    assert bool(torch.all(d2[valid] <= ps.radius**2 * (1 + 1e-14)))
    # exact brute-force counts for the sampled rows against ALL particles
    r2 = ps.radius**2
    for i in rows[:200].tolist():
        d = xyz - xyz[i]
        s = d[:, 0] * d[:, 0]
        s = s + d[:, 1] * d[:, 1]
        s = s + d[:, 2] * d[:, 2]
        assert int((s <= r2).sum()) - 1 == int(counts[i])
    total_full = full.total
    del full, nb, d2
    torch.cuda.empty_cache()
    half = cb.VerletList(x, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algorithm=cb.HALF)
    assert 2 * half.total == total_full


def test_cfg4_8m_clustered_matches_oracle(orc, cb):
    """cfg4 at its full 8 M particles against the oracle: counts, offsets and a hash of every
    row."""
    from _parity_helpers import full_size_oracle_compare as _full_size_oracle_compare

    ps = datasets.clustered(8_000_000)
    _full_size_oracle_compare(orc, cb, ps, cb.FULL, orc.FULL)
