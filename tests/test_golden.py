"""Committed golden vectors (tests/golden/verlet_golden.npz, made by tests/golden/make_golden.py).

CPU: the oracle still reproduces them (the goal posts cannot move silently).
GPU: the CUDA path through the C ABI reproduces them bit for bit -- counts and sorted rows,
CSR and 2D layouts.
"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "verlet_golden.npz")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.cases()


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(orc, golden, name):
    ps, b, e, ratio, algo = CASES[name]
    vl = orc.verlet_build(orc.view_from_xyz(ps.xyz), b, e, ps.radius, ratio, ps.grid_min, ps.grid_max,
                          algo=orc.FULL if algo == 0 else orc.HALF)
    flat, _ = vl.sorted_rows_flat()
    assert np.array_equal(np.asarray(vl.counts, dtype=np.int32), golden[name + "__counts"])
    assert np.array_equal(np.asarray(flat, dtype=np.int32), golden[name + "__rows"])


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_golden(orc, golden, name, layout):
    import torch

    assert torch.cuda.is_available()
    from cabana_b200 import core as cb

    ps, b, e, ratio, algo = CASES[name]
    x = cb.slice_from_array(ps.xyz, vlen=32)
    lst = cb.VerletList(x, b, e, ps.radius, ratio, ps.grid_min, ps.grid_max, algorithm=algo, layout=layout)
    counts = lst._data.counts.cpu().numpy()
    offsets = lst._data.offsets.cpu().numpy() if layout == 0 else None
    nb = lst._data.neighbors.cpu().numpy()
    flat, _ = orc.sorted_rows_flat(orc.CSR if layout == 0 else orc.LAYOUT_2D, counts, offsets, nb, lst.width)
    assert np.array_equal(counts.astype(np.int32), golden[name + "__counts"])
    assert np.array_equal(np.asarray(flat, dtype=np.int32), golden[name + "__rows"])
