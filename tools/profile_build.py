"""Three cfg3 builds on one GPU, nothing else: the target of the ncu captures committed under
profiles/ (13 kernels per build; `--launch-skip 26 --launch-count 13` profiles the third, warm one).

    ncu --set full --clock-control none --import-source on --launch-skip 26 --launch-count 13 \
        -o gpurun_out/prof python tools/profile_build.py
"""
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import bench  # noqa: E402
from cabana_b200 import core as cb  # noqa: E402

xyz, bounds, gmax = bench._fcc_slab(0, 1)
x = cb.slice_from_array(xyz, vlen=32)
n = xyz.shape[0]
lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
for _ in range(3):
    lst.build(x, 0, n, bench.RADIUS, 1.0, (0.0, 0.0, 0.0), gmax)
    torch.cuda.synchronize()
print("total", lst.total)
