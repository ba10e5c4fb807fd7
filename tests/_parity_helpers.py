"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch


def full_size_oracle_compare(orc, cb, ps, algo, oalgo, ratio=1.0):
    """counts, offsets and a 64-bit order-independent hash of every row, GPU list vs oracle list,
    at the full configuration size (the oracle uses every host core)."""
    orc.use_all_cores()
    x = cb.view_from_array(ps.xyz)
    lst = cb.VerletList(x, 0, ps.n, ps.radius, ratio, ps.grid_min, ps.grid_max, algorithm=algo,
                        layout=cb.CSR)
    counts = lst._data.counts.cpu().numpy()
    offsets = lst._data.offsets.cpu().numpy()
    nb = lst._data.neighbors[: lst.total].cpu().numpy()
    total, max_n = lst.total, lst._data.max_n
    del lst, x
    torch.cuda.empty_cache()
    h_gpu = orc.row_hashes(orc.CSR, counts, offsets, nb)
    del nb
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, ratio, ps.grid_min,
                           ps.grid_max, algo=oalgo, layout=orc.CSR)
    assert total == ref.total and max_n == ref.max_n
    assert np.array_equal(counts, ref.counts), "per-particle counts differ from the oracle"
    assert np.array_equal(offsets, ref.offsets), "offsets differ from the oracle"
    h_ref = orc.row_hashes(orc.CSR, ref.counts, ref.offsets, ref.neighbors)
    bad = np.nonzero(h_gpu != h_ref)[0]
    assert bad.size == 0, f"{bad.size} rows differ from the oracle, first {bad[:5]}"
