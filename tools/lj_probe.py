"""Time the LJ neighbor_parallel_for over the cfg3 list (one GPU): Serial staged / Serial direct / Team,
and check the two Serial kernels give bit-identical forces.  Usage (GPU box): python tools/lj_probe.py"""
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cabana_b200 import core as cb  # noqa: E402

xyz, bounds, gmax = bench._fcc_slab(0, 1)
x = cb.slice_from_array(xyz, vlen=32)
n = xyz.shape[0]
lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
lst.build(x, 0, n, bench.RADIUS, 1.0, (0.0, 0.0, 0.0), gmax)
torch.cuda.synchronize()
res = {}
for name, env, op in (("serial_staged", "staged", cb.OP_SERIAL), ("serial_direct", "direct", cb.OP_SERIAL),
                      ("team", "staged", cb.OP_TEAM)):
    os.environ["CB_LJ_SERIAL"] = env
    f = cb.view_from_array(np.zeros((n, 3)))
    cb.neighbor_parallel_for_lj(0, n, lst, x, f, 1.0, 1.0, 2.5, op)
    torch.cuda.synchronize()
    res[name] = f.to_array().clone()
    for _ in range(2):
        cb.neighbor_parallel_for_lj(0, n, lst, x, f, 1.0, 1.0, 2.5, op)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        cb.neighbor_parallel_for_lj(0, n, lst, x, f, 1.0, 1.0, 2.5, op)
    e1.record()
    torch.cuda.synchronize()
    print("%-14s %.3f ms  (%.3e pairs/s)" % (name, e0.elapsed_time(e1) / 5, lst.total / (e0.elapsed_time(e1) / 5 * 1e-3)), flush=True)
    del f
print("serial staged == direct bitwise:", bool(torch.equal(res["serial_staged"], res["serial_direct"])))
print("max |team - serial| / max|f|:", float((res["team"] - res["serial_direct"]).abs().max() / res["serial_direct"].abs().max()))
