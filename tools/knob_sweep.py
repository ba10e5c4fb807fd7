"""Sweep the pencil-grid knobs of the v2 build (CB_TILE_ZDIV, CB_TILE_ZB) on cfg3, one GPU.

Prints per-phase device times (ms: bin, plan, count, scan, fill, total) for every combination.
Usage (GPU box): python tools/knob_sweep.py [zdiv,zdiv,...] [zb,zb,...]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cabana_b200 import capi, core as cb  # noqa: E402

zdivs = [float(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["4", "6", "8"])]
zbs = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["16"])]
L = capi.lib()
xyz, bounds, gmax = bench._fcc_slab(0, 1)
x = cb.slice_from_array(xyz, vlen=32)
n = xyz.shape[0]
lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
capi.check(L.cb_verlet_set_profiling(lst._h, 1))
for zdiv in zdivs:
    for zb in zbs:
        os.environ["CB_TILE_ZDIV"] = str(zdiv)
        os.environ["CB_TILE_ZB"] = str(zb)
        reps = 6
        for it in range(2):
            lst.build(x, 0, n, bench.RADIUS, 1.0, (0.0, 0.0, 0.0), gmax)
        torch.cuda.synchronize()
        capi.check(L.cb_verlet_set_profiling(lst._h, 1))   # the averages start here
        for it in range(reps):
            lst.build(x, 0, n, bench.RADIUS, 1.0, (0.0, 0.0, 0.0), gmax)
        torch.cuda.synchronize()
        ph = (C.c_double * 6)()
        capi.check(L.cb_verlet_get_phase_times(lst._h, ph))
        print("zdiv %4.1f zb %3d total %d : bin %.3f plan %.3f count %.3f scan %.3f fill %.3f | step %.3f ms"
              % ((zdiv, zb, lst.total) + tuple(ph)), flush=True)
