"""Per-particle phase times of the cfg3 build for slabs of decreasing thickness (no ghosts): separates
what is fixed per launch from what scales with the particles."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402
from cabana_b200 import capi, core as cb, datasets  # noqa: E402

L = capi.lib()
a = (4.0 / datasets.FCC_DENSITY) ** (1.0 / 3.0)
for cells_x in (159, 80, 40, 20):
    ps = datasets.fcc_lattice(cells_x, radius=bench.RADIUS, cells_yz=bench.FCC_CELLS)
    x = cb.slice_from_array(ps.xyz, vlen=32)
    n = ps.xyz.shape[0]
    gmax = (cells_x * a, bench.FCC_CELLS * a, bench.FCC_CELLS * a)
    lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
    for _ in range(3):
        lst.build(x, 0, n, bench.RADIUS, 1.0, (0.0, 0.0, 0.0), gmax)
    torch.cuda.synchronize()
    capi.check(L.cb_verlet_set_profiling(lst._h, 1))
    for _ in range(8):
        lst.build(x, 0, n, bench.RADIUS, 1.0, (0.0, 0.0, 0.0), gmax)
    torch.cuda.synchronize()
    ph = (C.c_double * 6)()
    capi.check(L.cb_verlet_get_phase_times(lst._h, ph))
    print("cells_x %3d n %9d K %.2f : count %.4f fill %.4f ms | per particle count %.4f fill %.4f ns | per neighbour count %.4f fill %.4f ps"
          % (cells_x, n, lst.total / n, ph[2], ph[4], ph[2] * 1e6 / n, ph[4] * 1e6 / n,
             ph[2] * 1e9 / lst.total, ph[4] * 1e9 / lst.total), flush=True)
    del lst, x
