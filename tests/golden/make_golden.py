"""Generate tests/golden/verlet_golden.npz.

The reference itself cannot be built here (every header needs Kokkos), so the
golden vectors are produced by the oracle -- AFTER it has been pinned to the reference's literal
known answers (tests/test_oracle_known_answers.py) -- and cross-checked against the independent
N^2 construction of core/unit_test/neighbor_unit_test.hpp:86-158 before being written.  They
freeze, for every fixture below, the per-particle counts and the sorted neighbour rows, so that
(a) a later change of the oracle cannot silently move the goal posts and (b) the CUDA path is
checked against committed data.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from cabana_b200 import datasets  # noqa: E402


def cases():
    """name -> (ParticleSet, begin, end, cell_ratio, algo 0=full/1=half)"""
    r300 = datasets.fixture_random300()
    out = {}
    for algo in (0, 1):
        tag = "full" if algo == 0 else "half"
        out[f"random300_{tag}"] = (r300, 0, r300.n, r300.cell_ratio, algo)
        out[f"random300_partial_{tag}"] = (r300, 75, 225, r300.cell_ratio, algo)
        out[f"random300_ratio1_{tag}"] = (r300, 0, r300.n, 1.0, algo)
        out[f"ordered6_{tag}"] = (datasets.fixture_ordered(6), 0, 216, 0.5, algo)
        out[f"tutorial81_{tag}"] = (datasets.fixture_tutorial81(), 0, 81, 1.0, algo)
        out[f"near_cutoff_{tag}"] = (datasets.near_cutoff_adversarial(), 0, 800, 1.0, algo)
        out[f"fcc6_jitter_{tag}"] = (datasets.fcc_lattice(6, jitter=0.05), 0, 864, 1.0, algo)
        out[f"uniform2000_ratio03_{tag}"] = (datasets.uniform_box(2000, 7), 0, 2000, 0.3, algo)
    return out


def main():
    oracle.build()
    blob = {}
    for name, (ps, b, e, ratio, algo) in cases().items():
        x = oracle.view_from_xyz(ps.xyz)
        vl = oracle.verlet_build(x, b, e, ps.radius, ratio, ps.grid_min, ps.grid_max,
                                 algo=oracle.FULL if algo == 0 else oracle.HALF)
        flat, starts = vl.sorted_rows_flat()
        if algo == 0:
            # independent N^2 cross-check (neighbor_unit_test.hpp:86-158) restricted to [b,e)
            n2 = oracle.brute_force(x, ps.radius)
            f2, s2 = n2.sorted_rows_flat()
            for i in range(ps.n):
                row = flat[starts[i]:starts[i + 1]]
                ref = f2[s2[i]:s2[i + 1]] if b <= i < e else f2[0:0]
                assert np.array_equal(row, ref), (name, i)
        blob[name + "__counts"] = np.asarray(vl.counts, dtype=np.int32)
        blob[name + "__rows"] = np.asarray(flat, dtype=np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "verlet_golden.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), "bytes,", len(blob) // 2, "cases")


if __name__ == "__main__":
    main()
