"""Time the fused slab-halo plan (k_halo_compact + 16-byte read-back) on one GPU for slabs of the
cfg3 lattice as 2 and 8 ranks own them."""
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402
from cabana_b200 import comm, core as cb  # noqa: E402

k = comm.CudaCommKernels()
for world in (2, 8):
    xyz, bounds, gmax = bench._fcc_slab(1, world)
    x = cb.slice_from_array(xyz, vlen=32)
    n = xyz.shape[0]
    hw = bench.RADIUS * (1.0 + 2.0**-40)
    for _ in range(3):
        steer, n_lo, n_hi = k.slab_halo_plan(x, n, bounds[1] + hw, bounds[2] - hw, 0, 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        k.slab_halo_plan(x, n, bounds[1] + hw, bounds[2] - hw, 0, 2)
    e1.record()
    torch.cuda.synchronize()
    print("world %d: %d owned, ghosts %d + %d, plan %.1f us per call (incl. read-back)"
          % (world, n, n_lo, n_hi, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
