"""Host-side Python mirror of the reference interface for the hot path.

Same names and argument meaning as Cabana's C++ API (LinkedCellList, VerletList,
NeighborList traits, neighbor_parallel_for / neighbor_parallel_reduce), every call going
straight through the C ABI (include/cabana_b200.h) to the sm_100a kernels.  PyTorch is
used only for device memory and streams.  The C++ twin of this file is
include/Cabana_B200.hpp.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import capi
from .capi import (CSR, FULL, HALF, LAYOUT_2D, OP_SERIAL, OP_TEAM, OP_TEAM_VECTOR,  # noqa: F401
                   ROWS_BINNED, ROWS_REFERENCE)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _DevArray:
    """Zero-copy view of library-owned device memory for torch.as_tensor."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape),
            "typestr": typestr,
            "data": (int(ptr), False),
            "version": 2,
            "strides": None,
        }
        self._owner = owner


def _as_tensor(ptr, shape, typestr, torch_dtype, owner) -> torch.Tensor:
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return torch.empty(tuple(shape), dtype=torch_dtype, device="cuda")
    return torch.as_tensor(_DevArray(ptr, shape, typestr, owner), device="cuda")


# --------------------------------------------------------------------------------- positions
@dataclass
class Slice:
    """A device particle field in Cabana slice / rank-2 view layout.

    element(i,d) = data[outer_stride*(i // vlen) + (i % vlen) + comp_stride*d]
    (core/src/Cabana_Slice.hpp:134-140).  `data` is a flat torch CUDA tensor.
    """

    data: torch.Tensor
    n: int
    outer_stride: int
    vlen: int
    comp_stride: int
    num_comp: int = 3

    def size(self) -> int:
        return self.n

    def positions_desc(self) -> capi.Positions:
        assert self.data.dtype == torch.float64 and self.data.is_cuda
        return capi.Positions(self.data.data_ptr(), self.n, self.outer_stride, self.vlen, self.comp_stride)

    def field_desc(self) -> capi.Field:
        assert self.data.is_cuda
        return capi.Field(
            self.data.data_ptr(), self.n, self.outer_stride, self.vlen, self.comp_stride,
            self.num_comp, self.data.element_size(),
        )

    def _index(self):
        i = torch.arange(self.n, device=self.data.device)
        return self.outer_stride * (i // self.vlen) + (i % self.vlen)

    def to_array(self) -> torch.Tensor:
        """(n, num_comp) dense copy."""
        base = self._index()
        return torch.stack([self.data[base + self.comp_stride * d] for d in range(self.num_comp)], dim=1)


def view_from_array(a, device="cuda") -> Slice:
    """Rank-2 row-major (n,k) Kokkos::View analogue: vlen 1, strides (k,1)."""
    t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(device).contiguous()
    n, k = t.shape
    return Slice(t.reshape(-1), n, k, 1, 1, k)


def slice_from_array(a, vlen: int = 32, extra: int = 0, device="cuda") -> Slice:
    """Slice of member 0 of AoSoA<MemberTypes<T[k], T[extra]>>: Stride = (k+extra)*vlen."""
    t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a).to(device)
    n, k = t.shape
    stride = (k + extra) * vlen
    nsoa = max((n + vlen - 1) // vlen, 1)
    data = torch.full((nsoa * stride,), float("nan") if t.dtype.is_floating_point else -1,
                      dtype=t.dtype, device=device)
    i = torch.arange(n, device=device)
    base = stride * (i // vlen) + (i % vlen)
    for d in range(k):
        data[base + vlen * d] = t[:, d]
    return Slice(data, n, stride, vlen, vlen, k)


# --------------------------------------------------------------------------------- LinkedCellList
class LinkedCellList:
    """Cabana::LinkedCellList<MemorySpace, double, 3> (core/src/Cabana_LinkedCellList.hpp:128-909)."""

    def __init__(self, positions: Slice, grid_delta, grid_min, grid_max, begin=None, end=None,
                 neighborhood_radius=None, cell_size_ratio=1.0):
        L = capi.lib()
        self._h = C.c_void_p()
        radius = -1.0 if neighborhood_radius is None else float(neighborhood_radius)
        capi.check(L.cb_lcl_create(C.byref(self._h), capi.d3(grid_delta), capi.d3(grid_min),
                                   capi.d3(grid_max), C.c_double(radius), C.c_double(cell_size_ratio)))
        b = 0 if begin is None else int(begin)
        e = positions.size() if end is None else int(end)
        self.build(positions, b, e)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                capi.lib().cb_lcl_destroy(h)
            except Exception:
                pass

    def build(self, positions: Slice, begin=None, end=None):
        b = 0 if begin is None else int(begin)
        e = positions.size() if end is None else int(end)
        d = positions.positions_desc()
        capi.check(capi.lib().cb_lcl_build(self._h, C.byref(d), C.c_int64(b), C.c_int64(e), _stream()))
        self._refresh()

    def _refresh(self):
        v = capi.LclView()
        capi.check(capi.lib().cb_lcl_get(self._h, C.byref(v)))
        self._v = v
        np_ = v.end - v.begin
        self.counts = _as_tensor(v.counts, (v.num_cells,), "<i4", torch.int32, self)
        self.offsets = _as_tensor(v.offsets, (v.num_cells + 1,), "<i4", torch.int32, self)
        self.permutes = _as_tensor(v.permute, (np_,), "<i4", torch.int32, self)
        self.particle_bins = _as_tensor(v.particle_bins, (np_,), "<i4", torch.int32, self)

    # -- reference accessors (:462-872)
    def numParticles(self):
        return self._v.end - self._v.begin

    def getParticleBegin(self):
        return self._v.begin

    def getParticleEnd(self):
        return self._v.end

    rangeBegin = getParticleBegin
    rangeEnd = getParticleEnd

    def totalBins(self):
        return self._v.num_cells

    def numBin(self, dim):
        return self._v.grid.nx[dim]

    def cardinalBinIndex(self, i, j, k):
        return capi.lib().cb_grid_cardinal_cell_index(C.byref(self._v.grid), i, j, k)

    def ijkBinIndex(self, cardinal):
        out = (C.c_int32 * 3)()
        capi.lib().cb_grid_ijk_bin_index(C.byref(self._v.grid), int(cardinal), out)
        return tuple(out)

    def binSize(self, i, j, k):
        return int(self.counts[self.cardinalBinIndex(i, j, k)])

    def binOffset(self, i, j, k):
        return int(self.offsets[self.cardinalBinIndex(i, j, k)])

    def permutation(self, particle_id):
        return int(self.permutes[particle_id])

    def getParticleBins(self):
        return self.particle_bins

    def getParticleBin(self, particle_index):
        return int(self.particle_bins[particle_index - self._v.begin])

    def sorted(self):
        return bool(self._v.sorted)

    def update(self, sorted_):
        capi.check(capi.lib().cb_lcl_update(self._h, 1 if sorted_ else 0))
        self._refresh()

    def getStencilCells(self, cell):
        mn = (C.c_int32 * 3)()
        mx = (C.c_int32 * 3)()
        capi.check(capi.lib().cb_stencil_get_cells(C.byref(self._v.stencil_grid), self._v.cell_range,
                                                   int(cell), mn, mx))
        return tuple(mn), tuple(mx)

    def binningData(self) -> "BinningData":
        # :745-749
        return BinningData(self._v.begin, self._v.end, self.counts, self.offsets, self.permutes,
                           self.totalBins())

    def getParticle(self, offset):
        # :863-872
        return self.permutation(offset) if not self.sorted() else offset + self._v.begin


class BinningData:
    """Cabana::BinningData<MemorySpace> (core/src/Cabana_Sort.hpp:37-136): a value type over the
    binning arrays of a LinkedCellList (zero-copy device tensors)."""

    def __init__(self, begin, end, counts, offsets, permute_vector, nbin):
        self._begin, self._end, self._nbin = int(begin), int(end), int(nbin)
        self.counts, self.offsets, self.permute_vector = counts, offsets, permute_vector

    def numBin(self):
        return self._nbin

    def binSize(self, bin_id):
        return int(self.counts[bin_id])

    def binOffset(self, bin_id):
        return int(self.offsets[bin_id])

    def permutation(self, tuple_id):
        return int(self.permute_vector[tuple_id])

    def rangeBegin(self):
        return self._begin

    def rangeEnd(self):
        return self._end


def permute(binning, *fields: Slice):
    """Cabana::permute(LinkedCellList&, aosoa|slice) (Cabana_LinkedCellList.hpp:1130-1145) and
    Cabana::permute(BinningData, aosoa|slice) (Cabana_Sort.hpp:549-715).

    Pass every member slice of the AoSoA to permute the whole AoSoA.
    """
    arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
    if isinstance(binning, BinningData):
        capi.check(capi.lib().cb_binning_permute(
            C.c_int64(binning.rangeBegin()), C.c_int64(binning.rangeEnd()),
            C.c_void_p(binning.permute_vector.data_ptr()), arr, len(fields), _stream()))
        return
    capi.check(capi.lib().cb_lcl_permute(binning._h, arr, len(fields), _stream()))
    binning._refresh()


# --------------------------------------------------------------------------------- VerletList
class VerletListData:
    """Cabana::VerletListData<MemorySpace, LayoutTag> (core/src/Cabana_VerletList.hpp:50-113)."""

    counts: torch.Tensor
    offsets: torch.Tensor | None
    neighbors: torch.Tensor
    max_n: int


class VerletList:
    """Cabana::VerletList<MemorySpace, AlgorithmTag, LayoutTag, BuildTag, 3> (:824-1495).

    The template tags become keyword arguments.  `_data.counts/offsets/neighbors` are
    zero-copy views of the handle's device buffers, valid until the next build().
    """

    def __init__(self, x: Slice | None = None, begin=None, end=None, neighborhood_radius=None,
                 cell_size_ratio=None, grid_min=None, grid_max=None, max_neigh=0, *,
                 algorithm=FULL, layout=CSR, build_tag=OP_TEAM_VECTOR, row_placement=ROWS_REFERENCE):
        self.algorithm, self.layout, self.build_tag = algorithm, layout, build_tag
        self._h = C.c_void_p()
        capi.check(capi.lib().cb_verlet_create(C.byref(self._h)))
        # ROWS_BINNED: CSR rows stay in cell order (no offsets scan / reorder pass); see
        # cb_verlet_set_row_placement in include/cabana_b200.h
        self.row_placement = row_placement
        if row_placement != ROWS_REFERENCE:
            capi.check(capi.lib().cb_verlet_set_row_placement(self._h, C.c_int(row_placement)))
        self._data = VerletListData()
        if x is not None:
            self.build(x, begin, end, neighborhood_radius, cell_size_ratio, grid_min, grid_max, max_neigh)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                capi.lib().cb_verlet_destroy(h)
            except Exception:
                pass

    def build(self, x: Slice, begin, end, neighborhood_radius, cell_size_ratio, grid_min, grid_max,
              max_neigh=0):
        b = 0 if begin is None else int(begin)
        e = x.size() if end is None else int(end)
        d = x.positions_desc()
        capi.check(capi.lib().cb_verlet_build(
            self._h, C.byref(d), C.c_int64(b), C.c_int64(e), C.c_double(neighborhood_radius),
            C.c_double(cell_size_ratio), capi.d3(grid_min), capi.d3(grid_max), C.c_int64(max_neigh),
            C.c_int(self.algorithm), C.c_int(self.layout), C.c_int(self.build_tag), _stream()))
        self._refresh()

    def build_radii(self, x: Slice, begin, end, background_radius, neighborhood_radius: Slice,
                    cell_size_ratio, grid_min, grid_max, max_neigh=0):
        """VerletList( x, begin, end, background_radius, neighborhood_radius (slice/view, one double
        per particle), cell_size_ratio, grid_min, grid_max, max_neigh ) -- the per-particle cutoff
        constructor (core/src/Cabana_VerletList.hpp:989-1017)."""
        b = 0 if begin is None else int(begin)
        e = x.size() if end is None else int(end)
        d = x.positions_desc()
        rd = neighborhood_radius.field_desc()
        capi.check(capi.lib().cb_verlet_build_radii(
            self._h, C.byref(d), C.byref(rd), C.c_int64(b), C.c_int64(e), C.c_double(background_radius),
            C.c_double(cell_size_ratio), capi.d3(grid_min), capi.d3(grid_max), C.c_int64(max_neigh),
            C.c_int(self.algorithm), C.c_int(self.layout), C.c_int(self.build_tag), _stream()))
        self._refresh()

    def build_2d(self, x2: Slice, begin, end, neighborhood_radius, cell_size_ratio, grid_min2, grid_max2,
                 max_neigh=0):
        """VerletList<MemorySpace, AlgorithmTag, LayoutTag, BuildTag, 2>: positions with two
        components, two-component grid bounds (Cabana_VerletList.hpp:377-392, :626-639)."""
        b = 0 if begin is None else int(begin)
        e = x2.size() if end is None else int(end)
        d = x2.positions_desc()
        g2 = C.c_double * 2
        capi.check(capi.lib().cb_verlet_build_2d(
            self._h, C.byref(d), C.c_int64(b), C.c_int64(e), C.c_double(neighborhood_radius),
            C.c_double(cell_size_ratio), g2(*[float(v) for v in grid_min2]), g2(*[float(v) for v in grid_max2]),
            C.c_int64(max_neigh), C.c_int(self.algorithm), C.c_int(self.layout), C.c_int(self.build_tag),
            _stream()))
        self._refresh()

    def build_host(self, x_host, begin, end, neighborhood_radius, cell_size_ratio, grid_min, grid_max,
                   max_neigh=0):
        """End-to-end entry: positions in HOST memory (numpy / pinned torch CPU tensor, (n,3))."""
        t = x_host if isinstance(x_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_host))
        assert t.dtype == torch.float64 and not t.is_cuda and t.is_contiguous()
        n = t.shape[0]
        d = capi.Positions(t.data_ptr(), n, 3, 1, 1)
        b = 0 if begin is None else int(begin)
        e = n if end is None else int(end)
        capi.check(capi.lib().cb_verlet_build_host(
            self._h, C.byref(d), C.c_int64(b), C.c_int64(e), C.c_double(neighborhood_radius),
            C.c_double(cell_size_ratio), capi.d3(grid_min), capi.d3(grid_max), C.c_int64(max_neigh),
            C.c_int(self.algorithm), C.c_int(self.layout), C.c_int(self.build_tag), _stream()))
        self._refresh()

    def filter_selftest(self, x: Slice, neighborhood_radius, grid_min, grid_max):
        """cb_verlet_filter_selftest: (largest |tensor-core filter - exact FP64| seen over the tested
        pairs with s <= 4 r^2, the bound the build's decisions assume, pairs beyond the bound,
        values outside the exact tier's band with the wrong sign)."""
        d = x.positions_desc()
        out = (C.c_double * 4)()
        capi.check(capi.lib().cb_verlet_filter_selftest(
            self._h, C.byref(d), C.c_double(neighborhood_radius), capi.d3(grid_min), capi.d3(grid_max),
            C.c_int(self.algorithm), out, _stream()))
        return float(out[0]), float(out[1]), int(out[2]), int(out[3])

    def copy_to_host(self, counts_h: torch.Tensor, offsets_h: torch.Tensor | None, neighbors_h: torch.Tensor):
        capi.check(capi.lib().cb_verlet_copy_to_host(
            self._h, C.c_void_p(counts_h.data_ptr()),
            C.c_void_p(offsets_h.data_ptr() if offsets_h is not None else 0),
            C.c_void_p(neighbors_h.data_ptr()), C.c_int64(neighbors_h.numel()), _stream()))

    def _refresh(self):
        v = capi.VerletView()
        capi.check(capi.lib().cb_verlet_get(self._h, C.byref(v)))
        self._view = v
        d = self._data
        d.counts = _as_tensor(v.counts, (v.n,), "<i4", torch.int32, self)
        if v.layout == CSR:
            d.offsets = _as_tensor(v.offsets, (v.n,), "<i4", torch.int32, self)
            d.neighbors = _as_tensor(v.neighbors, (v.extent,), "<i4", torch.int32, self)
        else:
            d.offsets = None
            d.neighbors = _as_tensor(v.neighbors, (v.n, v.width), "<i4", torch.int32, self)
        d.max_n = int(v.max_n)
        self.total = int(v.total)
        self.extent = int(v.extent)
        self.width = int(v.width)
        self.refilled = bool(v.refilled)

    def setNeighbor(self, particle_index, neighbor_index, new_index):
        capi.check(capi.lib().cb_verlet_set_neighbor(
            self._h, C.c_int64(particle_index), C.c_int64(neighbor_index), C.c_int32(new_index), _stream()))


class NeighborList:
    """Cabana::NeighborList<VerletList<...>> traits (Cabana_VerletList.hpp:1603-1698)."""

    @staticmethod
    def totalNeighbor(lst: VerletList) -> int:
        return lst.total

    @staticmethod
    def maxNeighbor(lst: VerletList) -> int:
        return lst._data.max_n

    @staticmethod
    def numNeighbor(lst: VerletList, i: int) -> int:
        return int(lst._data.counts[i])

    @staticmethod
    def getNeighbor(lst: VerletList, i: int, n: int) -> int:
        if lst.layout == CSR:
            return int(lst._data.neighbors[int(lst._data.offsets[i]) + n])
        return int(lst._data.neighbors[i, n])


# --------------------------------------------------------------------------------- traversal
def neighbor_parallel_for_lj(begin, end, lst: VerletList, x: Slice, f: Slice, eps, sigma, rc,
                             op_tag=OP_SERIAL, newton=None):
    """neighbor_parallel_for(RangePolicy(begin,end), LJ functor, list, FirstNeighborsTag, op_tag)
    (core/src/Cabana_Parallel.hpp:251-293 / :386-435).  Forces are accumulated into f."""
    if newton is None:
        newton = lst.algorithm == HALF
    d = x.positions_desc()
    fd = f.field_desc()
    capi.check(capi.lib().cb_neighbor_for_lj(
        C.byref(lst._view), C.byref(d), C.byref(fd), C.c_double(eps), C.c_double(sigma), C.c_double(rc),
        C.c_int(1 if newton else 0), C.c_int(op_tag), C.c_int64(begin), C.c_int64(end), _stream()))


def neighbor_parallel_reduce_lj(begin, end, lst: VerletList, x: Slice, eps, sigma, rc,
                                op_tag=OP_SERIAL, scale=None) -> float:
    """neighbor_parallel_reduce with the LJ pair energy (Cabana_Parallel.hpp:638-685 / :787-844)."""
    if scale is None:
        scale = 1.0 if lst.algorithm == HALF else 0.5
    d = x.positions_desc()
    out = C.c_double(0.0)
    capi.check(capi.lib().cb_neighbor_reduce_lj(
        C.byref(lst._view), C.byref(d), C.c_double(eps), C.c_double(sigma), C.c_double(rc),
        C.c_double(scale), C.c_int(op_tag), C.c_int64(begin), C.c_int64(end), C.byref(out), _stream()))
    return out.value


def lcl_neighbor_parallel_for_lj(begin, end, lcl: LinkedCellList, x: Slice, f: Slice, eps, sigma, rc,
                                 op_tag=OP_SERIAL):
    """neighbor_parallel_for(policy, LJ functor, linked_cell_list, FirstNeighborsTag, op_tag):
    list-free traversal (core/src/Cabana_Parallel.hpp:1511-1595)."""
    d = x.positions_desc()
    fd = f.field_desc()
    capi.check(capi.lib().cb_lcl_neighbor_for_lj(
        lcl._h, C.byref(d), C.byref(fd), C.c_double(eps), C.c_double(sigma), C.c_double(rc),
        C.c_int(op_tag), C.c_int64(begin), C.c_int64(end), _stream()))


def lcl_neighbor_parallel_for_count(begin, end, lcl: LinkedCellList, x: Slice, cutoff,
                                    result: torch.Tensor, op_tag=OP_SERIAL):
    """The functor of checkLinkedCellNeighborInterface (tstLinkedCellList.hpp:704-780):
    result[i] += 1 for every stencil candidate j != i with r2 <= cutoff^2."""
    assert result.dtype == torch.int32 and result.is_cuda
    d = x.positions_desc()
    capi.check(capi.lib().cb_lcl_neighbor_for_count(
        lcl._h, C.byref(d), C.c_double(cutoff), C.c_void_p(result.data_ptr()), C.c_int(op_tag),
        C.c_int64(begin), C.c_int64(end), _stream()))


def neighbor_parallel_for_id_sum(begin, end, lst: VerletList, result: torch.Tensor, op_tag=OP_SERIAL):
    """The reference unit tests' functor: result[i] += j (neighbor_unit_test.hpp:291-348)."""
    assert result.dtype == torch.int64 and result.is_cuda
    capi.check(capi.lib().cb_neighbor_for_id_sum(
        C.byref(lst._view), C.c_void_p(result.data_ptr()), C.c_int(op_tag),
        C.c_int64(begin), C.c_int64(end), _stream()))
