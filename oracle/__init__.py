"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``cabana_b200`` never does.

The functions mirror ``oracle/cabana_oracle.cpp`` one to one; see that file for
the reference file:line each one restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_cabana.so")

FULL, HALF = 0, 1
CSR, LAYOUT_2D = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, seconds)."""
    src = os.path.join(_HERE, "cabana_oracle.cpp")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle_cabana.so"], check=True, capture_output=True)
    return _LIB_PATH


_NATIVE_PATH = os.path.join(_HERE, "liboracle_cabana_native.so")
_NATIVE_STAMP = _NATIVE_PATH + ".host"


def _host_tag() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def use_native_build() -> bool:
    """Switch this process to the -march=native copy (bench.py's timed CPU legs).  It is compiled
    on the host that runs it (the instruction set of the build container and of the GPU box's
    host may differ); returns False and keeps the portable copy when that fails."""
    global _lib
    src = os.path.join(_HERE, "cabana_oracle.cpp")
    tag = _host_tag()
    try:
        stale = (not os.path.exists(_NATIVE_PATH) or os.path.getmtime(_NATIVE_PATH) < os.path.getmtime(src)
                 or not os.path.exists(_NATIVE_STAMP) or open(_NATIVE_STAMP).read() != tag)
        if stale:
            subprocess.run(["make", "-C", _HERE, "-B", "liboracle_cabana_native.so"], check=True,
                           capture_output=True)
            with open(_NATIVE_STAMP, "w") as f:
                f.write(tag)
        _lib = None
        lib(_NATIVE_PATH)
        return True
    except Exception:
        _lib = None
        return False


class _Positions(C.Structure):
    _fields_ = [
        ("base", C.c_void_p),
        ("n", C.c_int64),
        ("outer_stride", C.c_int64),
        ("vlen", C.c_int32),
        ("comp_stride", C.c_int64),
    ]


class _Grid(C.Structure):
    _fields_ = [
        ("min", C.c_double * 3),
        ("max", C.c_double * 3),
        ("dx", C.c_double * 3),
        ("rdx", C.c_double * 3),
        ("nx", C.c_int * 3),
    ]


class _Stencil(C.Structure):
    _fields_ = [
        ("grid", _Grid),
        ("max_cells_dir", C.c_int),
        ("max_cells", C.c_int),
        ("cell_range", C.c_int),
    ]


class _VerletInfo(C.Structure):
    _fields_ = [
        ("total", C.c_int64),
        ("max_n", C.c_int64),
        ("width", C.c_int64),
        ("refilled", C.c_int32),
    ]


_lib = None


def lib(path: str | None = None):
    global _lib
    if _lib is None:
        if path is None:
            build()
            path = _LIB_PATH
        _lib = C.CDLL(path)
        _lib.orc_grid_min_distance.restype = C.c_double
        _lib.orc_lj_energy.restype = C.c_double
        _lib.orc_grid_cardinal.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _d3(v):
    return (C.c_double * 3)(*[float(a) for a in v])


def _i3(v):
    return (C.c_int * 3)(*[int(a) for a in v])


@dataclass
class PositionsView:
    """A host position field in Cabana slice or rank-2 view layout."""

    data: np.ndarray  # flat float64 storage
    n: int
    outer_stride: int
    vlen: int
    comp_stride: int

    def desc(self) -> _Positions:
        assert self.data.dtype == np.float64 and self.data.flags.c_contiguous
        return _Positions(
            self.data.ctypes.data, self.n, self.outer_stride, self.vlen, self.comp_stride
        )

    def to_xyz(self) -> np.ndarray:
        i = np.arange(self.n)
        base = self.outer_stride * (i // self.vlen) + (i % self.vlen)
        return np.stack([self.data[base + self.comp_stride * d] for d in range(3)], axis=1)


def view_from_xyz(xyz: np.ndarray) -> PositionsView:
    """Rank-2 (n,3) row-major view: vlen 1, strides (3,1)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    return PositionsView(xyz.reshape(-1), xyz.shape[0], 3, 1, 1)


def slice_from_xyz(xyz: np.ndarray, vlen: int = 16, extra_doubles: int = 0) -> PositionsView:
    """AoSoA<MemberTypes<double[3], double[extra]...>> position slice.

    Stride = (3 + extra) * vlen doubles per SoA (core/src/Cabana_AoSoA.hpp:192-196).
    """
    xyz = np.asarray(xyz, dtype=np.float64)
    n = xyz.shape[0]
    stride = (3 + extra_doubles) * vlen
    nsoa = (n + vlen - 1) // vlen
    data = np.full(max(nsoa, 1) * stride, np.nan, dtype=np.float64)
    i = np.arange(n)
    base = stride * (i // vlen) + (i % vlen)
    for d in range(3):
        data[base + vlen * d] = xyz[:, d]
    return PositionsView(data, n, stride, vlen, vlen)


# ----------------------------------------------------------------------------- grid
class Grid:
    def __init__(self, gmin, gmax, delta):
        self.c = _Grid()
        lib().orc_grid_init(C.byref(self.c), _d3(gmin), _d3(gmax), _d3(delta))

    @property
    def nx(self):
        return tuple(self.c.nx)

    @property
    def dx(self):
        return tuple(self.c.dx)

    @property
    def rdx(self):
        return tuple(self.c.rdx)

    @property
    def total_cells(self):
        return self.c.nx[0] * self.c.nx[1] * self.c.nx[2]

    def locate(self, p):
        out = (C.c_int * 3)()
        lib().orc_grid_locate(C.byref(self.c), _d3(p), out)
        return tuple(out)

    def min_distance(self, x, c):
        return lib().orc_grid_min_distance(C.byref(self.c), _d3(x), _i3(c))

    def cardinal(self, i, j, k):
        return lib().orc_grid_cardinal(C.byref(self.c), i, j, k)

    def ijk(self, c):
        out = (C.c_int * 3)()
        lib().orc_grid_ijk(C.byref(self.c), int(c), out)
        return tuple(out)


class Stencil:
    def __init__(self, radius, ratio, gmin, gmax):
        self.c = _Stencil()
        lib().orc_stencil_init(C.byref(self.c), C.c_double(radius), C.c_double(ratio), _d3(gmin), _d3(gmax))

    @property
    def cell_range(self):
        return self.c.cell_range

    @property
    def nx(self):
        return tuple(self.c.grid.nx)

    def cells(self, cell):
        mn = (C.c_int * 3)()
        mx = (C.c_int * 3)()
        lib().orc_stencil_cells(C.byref(self.c), int(cell), mn, mx)
        return tuple(mn), tuple(mx)


# ----------------------------------------------------------------------------- LCL
@dataclass
class LinkedCellResult:
    counts: np.ndarray
    offsets: np.ndarray
    permute: np.ndarray
    particle_bins: np.ndarray
    grid: Grid


def lcl_build(x: PositionsView, begin, end, delta, gmin, gmax, parallel=False) -> LinkedCellResult:
    g = Grid(gmin, gmax, delta)
    nc = g.total_cells
    counts = np.zeros(nc, dtype=np.int32)
    offsets = np.zeros(nc, dtype=np.int64)
    permute = np.zeros(max(end - begin, 0), dtype=np.int64)
    bins = np.full(max(end - begin, 0), -1, dtype=np.int32)
    d = x.desc()
    lib().orc_lcl_build(
        C.byref(g.c),
        C.byref(d),
        C.c_int64(begin),
        C.c_int64(end),
        counts.ctypes.data_as(C.c_void_p),
        offsets.ctypes.data_as(C.c_void_p),
        permute.ctypes.data_as(C.c_void_p),
        bins.ctypes.data_as(C.c_void_p),
        C.c_int(1 if parallel else 0),
    )
    return LinkedCellResult(counts, offsets, permute, bins, g)


def permute_slice(x: PositionsView, num_comp, begin, end, permute: np.ndarray):
    """In-place Cabana::permute(BinningData, slice) on a host slice."""
    permute = np.ascontiguousarray(permute, dtype=np.int64)
    lib().orc_permute_slice(
        x.data.ctypes.data_as(C.c_void_p),
        C.c_int64(x.outer_stride),
        C.c_int32(x.vlen),
        C.c_int64(x.comp_stride),
        C.c_int(num_comp),
        C.c_int64(begin),
        C.c_int64(end),
        permute.ctypes.data_as(C.c_void_p),
    )


# ----------------------------------------------------------------------------- Verlet
@dataclass
class VerletResult:
    layout: int
    counts: np.ndarray
    offsets: np.ndarray | None  # CSR only
    neighbors: np.ndarray  # CSR: [total]; 2D: [n, width] row-major
    total: int
    max_n: int
    width: int
    refilled: bool

    def row(self, i: int) -> np.ndarray:
        c = int(self.counts[i])
        if self.layout == CSR:
            o = int(self.offsets[i])
            return self.neighbors[o : o + c]
        return self.neighbors[i, : min(c, self.width)]

    def sorted_rows_flat(self):
        """(row-sorted neighbour ids concatenated in particle order, row starts)."""
        return sorted_rows_flat(self.layout, self.counts, self.offsets, self.neighbors, self.width)


def sorted_rows_flat(layout, counts, offsets, neighbors, width):
    """Canonical form for set comparison: each row sorted, rows concatenated."""
    counts = np.asarray(counts, dtype=np.int64)
    n = counts.shape[0]
    starts = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=starts[1:])
    total = int(starts[-1])
    if total == 0:
        return np.zeros(0, dtype=np.int64), starts
    row_id = np.repeat(np.arange(n, dtype=np.int64), counts)
    within = np.arange(total, dtype=np.int64) - starts[row_id]
    if layout == CSR:
        vals = np.asarray(neighbors)[np.asarray(offsets, dtype=np.int64)[row_id] + within]
    else:
        nb = np.asarray(neighbors).reshape(n, -1)
        assert counts.max() <= nb.shape[1], "2D row overflow: counts exceed allocated width"
        vals = nb[row_id, within]
    vals = vals.astype(np.int64)
    order = np.lexsort((vals, row_id))
    return vals[order], starts


def verlet_build(
    x: PositionsView, begin, end, radius, ratio, gmin, gmax, max_neigh=0, algo=FULL, layout=CSR
) -> VerletResult:
    n = x.n
    counts = np.zeros(n, dtype=np.int32)
    offsets = np.zeros(n, dtype=np.int32)
    nb_ptr = C.POINTER(C.c_int)()
    info = _VerletInfo()
    d = x.desc()
    rc = lib().orc_verlet_build(
        C.byref(d),
        C.c_int64(begin),
        C.c_int64(end),
        C.c_double(radius),
        C.c_double(ratio),
        _d3(gmin),
        _d3(gmax),
        C.c_int64(max_neigh),
        C.c_int(algo),
        C.c_int(layout),
        counts.ctypes.data_as(C.c_void_p),
        offsets.ctypes.data_as(C.c_void_p),
        C.byref(nb_ptr),
        C.byref(info),
    )
    assert rc == 0
    if layout == CSR:
        size = int(info.total)
    else:
        size = n * int(info.width)
    if size > 0:
        nb = np.ctypeslib.as_array(nb_ptr, shape=(size,)).copy()
    else:
        nb = np.zeros(0, dtype=np.int32)
    lib().orc_free(nb_ptr)
    if layout == LAYOUT_2D:
        nb = nb.reshape(n, int(info.width))
    return VerletResult(
        layout,
        counts,
        offsets if layout == CSR else None,
        nb,
        int(info.total),
        int(info.max_n),
        int(info.width),
        bool(info.refilled),
    )


def verlet_build_radii(x: PositionsView, radii, begin, end, background_radius, ratio, gmin, gmax,
                       max_neigh=0, algo=FULL, layout=CSR):
    """Per-particle cutoff radius build (Cabana_VerletList.hpp:181-203, :244-305).  Returns
    (VerletResult with the FILL semantics, literal count-pass counts) -- see the note on the
    reference's count/fill asymmetry in cabana_oracle.cpp."""
    n = x.n
    radii = np.ascontiguousarray(radii, dtype=np.float64)
    assert radii.shape == (n,)
    counts = np.zeros(n, dtype=np.int32)
    offsets = np.zeros(n, dtype=np.int32)
    cp_counts = np.zeros(n, dtype=np.int32)
    nb_ptr = C.POINTER(C.c_int)()
    info = _VerletInfo()
    d = x.desc()
    rc = lib().orc_verlet_build_radii(
        C.byref(d), radii.ctypes.data_as(C.c_void_p), C.c_int64(begin), C.c_int64(end),
        C.c_double(background_radius), C.c_double(ratio), _d3(gmin), _d3(gmax), C.c_int64(max_neigh),
        C.c_int(algo), C.c_int(layout), counts.ctypes.data_as(C.c_void_p),
        offsets.ctypes.data_as(C.c_void_p), C.byref(nb_ptr), cp_counts.ctypes.data_as(C.c_void_p),
        C.byref(info))
    assert rc == 0
    size = int(info.total) if layout == CSR else n * int(info.width)
    nb = np.ctypeslib.as_array(nb_ptr, shape=(size,)).copy() if size > 0 else np.zeros(0, dtype=np.int32)
    lib().orc_free(nb_ptr)
    if layout == LAYOUT_2D:
        nb = nb.reshape(n, int(info.width))
    res = VerletResult(layout, counts, offsets if layout == CSR else None, nb, int(info.total),
                       int(info.max_n), int(info.width), bool(info.refilled))
    return res, cp_counts


def verlet_build_2d(xy: np.ndarray, begin, end, radius, ratio, gmin, gmax, algo=FULL) -> VerletResult:
    """NumSpaceDim = 2 build (CSR): xy is an (n,2) array."""
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    n = xy.shape[0]
    x = PositionsView(xy.reshape(-1), n, 2, 1, 1)
    counts = np.zeros(n, dtype=np.int32)
    offsets = np.zeros(n, dtype=np.int32)
    nb_ptr = C.POINTER(C.c_int)()
    info = _VerletInfo()
    d = x.desc()
    g2 = (C.c_double * 2)
    rc = lib().orc_verlet_build_2d(
        C.byref(d), C.c_int64(begin), C.c_int64(end), C.c_double(radius), C.c_double(ratio),
        g2(*[float(v) for v in gmin]), g2(*[float(v) for v in gmax]), C.c_int(algo),
        counts.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), C.byref(nb_ptr),
        C.byref(info))
    assert rc == 0
    size = int(info.total)
    nb = np.ctypeslib.as_array(nb_ptr, shape=(size,)).copy() if size > 0 else np.zeros(0, dtype=np.int32)
    lib().orc_free(nb_ptr)
    return VerletResult(CSR, counts, offsets, nb, int(info.total), int(info.max_n), 0, False)


def brute_force(x: PositionsView, radius, with_neighbors=True) -> VerletResult:
    """N^2 list (core/unit_test/neighbor_unit_test.hpp:86-158), 2D row-major."""
    n = x.n
    counts = np.zeros(n, dtype=np.int32)
    d = x.desc()
    lib().orc_brute_force(
        C.byref(d), C.c_double(radius), counts.ctypes.data_as(C.c_void_p), None, C.c_int64(0)
    )
    width = int(counts.max()) if n else 0
    nb = np.zeros((n, max(width, 1)), dtype=np.int32)
    if with_neighbors and n:
        lib().orc_brute_force(
            C.byref(d),
            C.c_double(radius),
            counts.ctypes.data_as(C.c_void_p),
            nb.ctypes.data_as(C.c_void_p),
            C.c_int64(nb.shape[1]),
        )
    return VerletResult(
        LAYOUT_2D, counts, None, nb, int(counts.sum()), width, nb.shape[1], False
    )


# ----------------------------------------------------------------------------- traversal
def _list_args(layout, counts, offsets, neighbors, width):
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    neighbors = np.ascontiguousarray(neighbors, dtype=np.int32)
    if offsets is None:
        offsets = np.zeros(1, dtype=np.int32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    keep = (counts, offsets, neighbors)
    return keep, (
        C.c_int(layout),
        counts.ctypes.data_as(C.c_void_p),
        offsets.ctypes.data_as(C.c_void_p),
        neighbors.ctypes.data_as(C.c_void_p),
        C.c_int64(width),
    )


def lj_forces(x: PositionsView, layout, counts, offsets, neighbors, width, begin, end,
              eps, sigma, rc, newton=False):
    """Returns (f[n,3], fabs[n,3]) -- fabs is the per-component sum of |pair force|."""
    n = x.n
    f = np.zeros((n, 3), dtype=np.float64)
    fabs = np.zeros((n, 3), dtype=np.float64)
    keep, la = _list_args(layout, counts, offsets, neighbors, width)
    d = x.desc()
    lib().orc_lj_forces(
        C.byref(d), *la, C.c_int64(begin), C.c_int64(end), C.c_double(eps), C.c_double(sigma),
        C.c_double(rc), C.c_int(1 if newton else 0),
        f.ctypes.data_as(C.c_void_p), fabs.ctypes.data_as(C.c_void_p),
    )
    return f, fabs


def lj_energy(x: PositionsView, layout, counts, offsets, neighbors, width, begin, end,
              eps, sigma, rc, scale):
    keep, la = _list_args(layout, counts, offsets, neighbors, width)
    d = x.desc()
    return lib().orc_lj_energy(
        C.byref(d), *la, C.c_int64(begin), C.c_int64(end), C.c_double(eps), C.c_double(sigma),
        C.c_double(rc), C.c_double(scale),
    )


def neighbor_id_sum(layout, counts, offsets, neighbors, width, begin, end):
    n = len(counts)
    out = np.zeros(n, dtype=np.int64)
    keep, la = _list_args(layout, counts, offsets, neighbors, width)
    lib().orc_neighbor_id_sum(*la, C.c_int64(begin), C.c_int64(end), out.ctypes.data_as(C.c_void_p))
    return out


def lcl_neighbor_for(x: PositionsView, res: LinkedCellResult, stencil: Stencil, sorted_, lcl_begin,
                     pbegin, pend, mode, cutoff, eps=1.0, sigma=1.0):
    """LinkedCellParallelFor with the count (mode 0) or LJ (mode 1) functor."""
    n = x.n
    counts_out = np.zeros(n, dtype=np.int32)
    f = np.zeros((n, 3), dtype=np.float64)
    fabs = np.zeros((n, 3), dtype=np.float64)
    d = x.desc()
    offsets = np.ascontiguousarray(res.offsets, dtype=np.int64)
    permute = np.ascontiguousarray(res.permute, dtype=np.int64)
    lib().orc_lcl_neighbor_for(
        C.byref(d), C.byref(res.grid.c), C.byref(stencil.c),
        res.counts.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p),
        permute.ctypes.data_as(C.c_void_p), res.particle_bins.ctypes.data_as(C.c_void_p),
        C.c_int(1 if sorted_ else 0), C.c_int64(lcl_begin), C.c_int64(pbegin), C.c_int64(pend),
        C.c_int(mode), C.c_double(cutoff), C.c_double(eps), C.c_double(sigma),
        counts_out.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p),
        fabs.ctypes.data_as(C.c_void_p))
    return counts_out, f, fabs


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


def use_all_cores() -> int:
    """torch.distributed.run exports OMP_NUM_THREADS=1: the timed CPU legs must not inherit it."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    set_num_threads(max(1, n))
    return num_threads()


def row_hashes(layout, counts, offsets, neighbors, width=0) -> np.ndarray:
    """Order-independent 64-bit hash of every row (equal <=> equal sorted rows)."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    n = counts.shape[0]
    nb = np.ascontiguousarray(neighbors, dtype=np.int32).reshape(-1)
    off = np.ascontiguousarray(offsets, dtype=np.int32) if offsets is not None else None
    out = np.empty(n, dtype=np.uint64)
    lib().orc_row_hashes(
        C.c_int(layout), C.c_int64(n), counts.ctypes.data_as(C.c_void_p),
        off.ctypes.data_as(C.c_void_p) if off is not None else None,
        nb.ctypes.data_as(C.c_void_p), C.c_int64(width), out.ctypes.data_as(C.c_void_p))
    return out
