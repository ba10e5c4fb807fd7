"""One rank's share of the 8-GPU cfg3 step on ONE GPU: the owned slab of rank 3 of 8 plus its two
ghost layers (taken from the neighbouring lattice cells), built with begin = 0, end = num_local.
Prints the averaged phase times; run under `ncu --metrics gpu__time_duration.sum` for a launch list."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cabana_b200 import capi, core as cb, datasets  # noqa: E402

world, rank = 8, 3
a = (4.0 / datasets.FCC_DENSITY) ** (1.0 / 3.0)
cuts = [round(bench.FCC_CELLS * g / world) for g in range(world + 1)]
c0, c1 = cuts[rank], cuts[rank + 1]
pad = int(np.ceil(bench.RADIUS / a)) + 1
ps = datasets.fcc_lattice(c1 - c0 + 2 * pad, radius=bench.RADIUS, cells_yz=bench.FCC_CELLS)
xyz = ps.xyz
xyz[:, 0] += (c0 - pad) * a
lo, hi = c0 * a, c1 * a
hw = bench.RADIUS * (1.0 + 2.0**-40)
own = (xyz[:, 0] >= lo) & (xyz[:, 0] < hi)
gho = ~own & (xyz[:, 0] >= lo - hw) & (xyz[:, 0] < hi + hw)
allx = np.concatenate([xyz[own], xyz[gho]])
nl, n = int(own.sum()), allx.shape[0]
x = cb.slice_from_array(allx, vlen=32)
gmax = bench.FCC_CELLS * a
lmin, lmax = (lo - hw, 0.0, 0.0), (hi + hw, gmax, gmax)
L = capi.lib()
lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
for _ in range(3):
    lst.build(x, 0, nl, bench.RADIUS, 1.0, lmin, lmax)
torch.cuda.synchronize()
capi.check(L.cb_verlet_set_profiling(lst._h, 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 20
for _ in range(reps):
    lst.build(x, 0, nl, bench.RADIUS, 1.0, lmin, lmax)
e1.record()
torch.cuda.synchronize()
ph = (C.c_double * 6)()
capi.check(L.cb_verlet_get_phase_times(lst._h, ph))
print("owned %d ghosts %d total %d : bin %.3f plan %.3f count %.3f scan %.3f fill %.3f | build %.3f ms | loop %.3f ms/build"
      % ((nl, n - nl, lst.total) + tuple(ph) + (e0.elapsed_time(e1) / reps,)), flush=True)
