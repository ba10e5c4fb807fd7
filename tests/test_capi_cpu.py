"""CPU-only checks of the C-ABI library: it loads, exports every symbol the header
declares, its host-side grid helpers agree with the oracle, and compute entry points fail
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C

import numpy as np
import pytest

from cabana_b200 import capi


@pytest.fixture(scope="module")
def L():
    from cabana_b200 import build

    build.build()
    return capi.lib()


def test_library_exports_every_declared_symbol(L):
    names = capi.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/cabana_b200.h but not exported: {missing}"
    assert L.cb_version() == 1


def test_host_grid_helpers_match_oracle(L, orc):
    # tstCartesianGrid.cpp:22-59 through the C ABI
    g = capi.Grid()
    capi.check(L.cb_grid_init(C.byref(g), capi.d3((-1.0, -0.5, -0.6)), capi.d3((2.5, 1.5, 1.9)),
                              capi.d3((0.5, 0.125, 0.25))))
    assert tuple(g.nx) == (7, 16, 10)
    ijk = (C.c_int32 * 3)()
    L.cb_grid_locate_point(C.byref(g), capi.d3((-0.9, 1.4, 0.1)), ijk)
    assert tuple(ijk) == (0, 15, 2)
    assert L.cb_grid_min_distance_to_point(C.byref(g), capi.d3((-0.9, 1.4, 0.1)), ijk) == 0.0
    L.cb_grid_locate_point(C.byref(g), capi.d3((2.5, 1.5, 1.9)), ijk)
    assert tuple(ijk) == (6, 15, 9)

    # random points: bit-identical to the oracle's restatement
    rng = np.random.default_rng(5)
    og = orc.Grid((-3.3, 0.1, -7.0), (11.9, 12.7, 5.5), (0.37, 0.41, 0.29))
    capi.check(L.cb_grid_init(C.byref(g), capi.d3((-3.3, 0.1, -7.0)), capi.d3((11.9, 12.7, 5.5)),
                              capi.d3((0.37, 0.41, 0.29))))
    assert tuple(g.nx) == og.nx and tuple(g.dx) == og.dx and tuple(g.rdx) == og.rdx
    for _ in range(300):
        p = rng.random(3) * np.array([15.2, 12.6, 12.5]) + np.array([-3.3, 0.1, -7.0])
        L.cb_grid_locate_point(C.byref(g), capi.d3(p), ijk)
        assert tuple(ijk) == og.locate(p)
        c = rng.integers(0, 10, 3)
        ci = (C.c_int32 * 3)(*[int(v) for v in c])
        assert L.cb_grid_min_distance_to_point(C.byref(g), capi.d3(p), ci) == og.min_distance(p, c)


def test_stencil_helpers(L, orc):
    # tstLinkedCellList.hpp:445-518
    g = capi.Grid()
    capi.check(L.cb_grid_init(C.byref(g), capi.d3((0, 0, 0)), capi.d3((10, 10, 10)), capi.d3((1, 1, 1))))
    assert L.cb_stencil_cell_range(C.c_double(1.0)) == 1
    assert L.cb_stencil_cell_range(C.c_double(0.5)) == 2
    mn = (C.c_int32 * 3)()
    mx = (C.c_int32 * 3)()
    cell = L.cb_grid_cardinal_cell_index(C.byref(g), 4, 5, 3)
    L.cb_stencil_get_cells(C.byref(g), 1, cell, mn, mx)
    assert tuple(mn) == (3, 4, 2) and tuple(mx) == (6, 7, 5)
    L.cb_stencil_get_cells(C.byref(g), 1, L.cb_grid_cardinal_cell_index(C.byref(g), 9, 9, 9), mn, mx)
    assert tuple(mn) == (8, 8, 8) and tuple(mx) == (10, 10, 10)


def test_argument_validation_mirrors_reference_asserts(L):
    h = C.c_void_p()
    capi.check(L.cb_verlet_create(C.byref(h)))
    x = capi.Positions(0, 10, 3, 1, 1)
    # end > size(x): assert( end <= size( x ) ) Cabana_VerletList.hpp:1382
    rc = L.cb_verlet_build(h, C.byref(x), C.c_int64(0), C.c_int64(11), C.c_double(1.0), C.c_double(1.0),
                           capi.d3((0, 0, 0)), capi.d3((1, 1, 1)), C.c_int64(0), 0, 0, 2, None)
    assert rc == capi.CB_ERR_INVALID
    assert b"range" in L.cb_last_error_string()
    rc = L.cb_verlet_build(h, C.byref(x), C.c_int64(0), C.c_int64(10), C.c_double(1.0), C.c_double(1.0),
                           capi.d3((0, 0, 0)), capi.d3((1, 1, 1)), C.c_int64(0), 7, 0, 2, None)
    assert rc == capi.CB_ERR_INVALID
    v = capi.VerletView()
    assert L.cb_verlet_get(h, C.byref(v)) == capi.CB_ERR_INVALID  # not built
    L.cb_verlet_destroy(h)


def test_no_cpu_fallback_without_device(L):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert L.cb_device_count() == 0
    # A well-formed build request must fail with CB_ERR_CUDA, not silently run on the CPU.
    xyz = np.random.default_rng(0).random((64, 3))
    h = C.c_void_p()
    capi.check(L.cb_verlet_create(C.byref(h)))
    x = capi.Positions(xyz.ctypes.data, 64, 3, 1, 1)
    rc = L.cb_verlet_build(h, C.byref(x), C.c_int64(0), C.c_int64(64), C.c_double(0.3), C.c_double(1.0),
                           capi.d3((0, 0, 0)), capi.d3((1, 1, 1)), C.c_int64(0), 0, 0, 2, None)
    assert rc == capi.CB_ERR_CUDA
    L.cb_verlet_destroy(h)
