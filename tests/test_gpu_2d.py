"""NumSpaceDim = 2 (VerletList<..., 2>, Cabana_VerletList.hpp:377-392 / :626-639): the CUDA path
(3-D kernels on [x, y, 0] with one shared z cell) against the oracle's genuine 2-D restatement
and an N^2 list."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    assert torch.cuda.is_available()
    from cabana_b200 import core

    return core


def _brute_2d(xy, r, half):
    d = xy[:, None, :] - xy[None, :, :]
    s = d[..., 0] * d[..., 0]
    s = s + d[..., 1] * d[..., 1]
    hit = s <= r * r
    np.fill_diagonal(hit, False)
    if half:
        xn, xp = xy[None, :, :], xy[:, None, :]
        hit &= (xn[..., 0] > xp[..., 0]) | ((xn[..., 0] == xp[..., 0]) & (xn[..., 1] > xp[..., 1]))
    return [np.nonzero(row)[0] for row in hit]


@pytest.mark.parametrize("algo", ["full", "half"])
@pytest.mark.parametrize("layout", ["csr", "2d"])
@pytest.mark.parametrize("ratio", [1.0, 0.5])
def test_verlet_2d_matches_oracle_and_brute_force(orc, cb, algo, layout, ratio):
    rng = np.random.Generator(np.random.Philox(key=404))
    n, r, lo, hi = 3000, 2.32, -12.296, 10.904
    xy = lo + rng.random((n, 2)) * (hi - lo)
    xy[:20, 0] = xy[20:40, 0]                 # equal x: the y tie-break of the half criterion
    xy[40:60] = xy[60:80]                     # coincident points (dropped by half lists)
    xy[80] = (hi, hi)                         # a point on the upper corner
    a = cb.FULL if algo == "full" else cb.HALF
    lay = cb.CSR if layout == "csr" else cb.LAYOUT_2D
    x2 = cb.view_from_array(xy)
    lst = cb.VerletList(algorithm=a, layout=lay)
    lst.build_2d(x2, 0, n, r, ratio, (lo, lo), (hi, hi))
    counts = lst._data.counts.cpu().numpy()
    offsets = lst._data.offsets.cpu().numpy() if lay == cb.CSR else None
    nb = lst._data.neighbors.cpu().numpy()
    got, _ = orc.sorted_rows_flat(lay, counts, offsets, nb, lst.width)
    ref = orc.verlet_build_2d(xy, 0, n, r, ratio, (lo, lo), (hi, hi), algo=orc.FULL if algo == "full" else orc.HALF)
    assert np.array_equal(counts, ref.counts)
    assert np.array_equal(got, ref.sorted_rows_flat()[0])
    rows = _brute_2d(xy, r, algo == "half")
    assert np.array_equal(counts, np.array([len(q) for q in rows]))
    assert np.array_equal(got, np.concatenate(rows))


def test_verlet_2d_partial_range_and_slice_layout(orc, cb):
    rng = np.random.Generator(np.random.Philox(key=405))
    n, r = 20_000, 1.7
    hi = 120.0
    xy = rng.random((n, 2)) * hi
    x2 = cb.slice_from_array(xy, vlen=32)      # AoSoA member double[2]
    lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
    lst.build_2d(x2, 5000, 15000, r, 1.0, (0.0, 0.0), (hi, hi))
    ref = orc.verlet_build_2d(xy, 5000, 15000, r, 1.0, (0.0, 0.0), (hi, hi))
    counts = lst._data.counts.cpu().numpy()
    assert np.array_equal(counts, ref.counts) and counts[:5000].sum() == 0 and counts[15000:].sum() == 0
    got, _ = orc.sorted_rows_flat(orc.CSR, counts, lst._data.offsets.cpu().numpy(),
                                  lst._data.neighbors.cpu().numpy(), 0)
    assert np.array_equal(got, ref.sorted_rows_flat()[0])
