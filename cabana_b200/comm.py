"""Halo / Distributor for the slab decomposition, over torch.distributed (NCCL on NVLink).

Mirrors the reference's communication-plan surface for the hot path:
  CommunicationPlan   core/src/Cabana_CommunicationPlanBase.hpp (counts, steering,
                      neighbour order: self first, then ascending ranks, :374-394)
  Halo / gather / scatter      core/src/Cabana_Halo.hpp:59-268, :392-682;
                               core/src/impl/Cabana_Halo_Mpi.hpp:41-125, :269-350
  Distributor / migrate        core/src/Cabana_Distributor.hpp:62-146, :275-337;
                               core/src/impl/Cabana_Migrate_Mpi.hpp:41-177
with CommSpaceType = Nccl instead of Mpi: export counts are
exchanged with ONE all_gather of a world-size vector (replacing a point-to-point per
neighbour), payloads with grouped send/recv on the current CUDA stream between the pack
and unpack kernels (no host fence, no barrier).

The device work (count/steer, pack, unpack, scatter-add, slab selection) is the C ABI's
(include/cabana_b200.h, csrc/cb_comm.cu).  `kernels` is injectable so the host-side plan
logic can be exercised with world_size-2 gloo tests on a CPU-only box; the package itself
only ships the CUDA implementation (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import capi
from .core import Slice, _stream


class CudaCommKernels:
    """The product implementation: csrc/cb_comm.cu through the C ABI."""

    device = "cuda"

    def count_and_steer(self, export_ranks: torch.Tensor, export_ids: torch.Tensor | None, num_ranks: int):
        n = export_ranks.numel()
        counts = (C.c_int64 * num_ranks)()
        offsets = (C.c_int64 * (num_ranks + 1))()
        steering = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
        assert export_ranks.dtype == torch.int32 and export_ranks.is_cuda
        ids_ptr = C.c_void_p(export_ids.data_ptr()) if export_ids is not None else None
        capi.check(capi.lib().cb_comm_count_and_steer(
            C.c_void_p(export_ranks.data_ptr()), C.c_int64(n), C.c_int(num_ranks), counts, offsets,
            C.c_void_p(steering.data_ptr()), ids_ptr, _stream()))
        return list(counts), list(offsets), steering

    def tuple_bytes(self, fields) -> int:
        arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
        return int(capi.lib().cb_comm_tuple_bytes(arr, len(fields)))

    def pack(self, fields, steering: torch.Tensor, count: int, out: torch.Tensor):
        arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
        capi.check(capi.lib().cb_comm_pack(arr, len(fields), C.c_void_p(steering.data_ptr()),
                                           C.c_int64(count), C.c_void_p(out.data_ptr()), _stream()))

    def unpack(self, fields, dst_begin: int, count: int, buf: torch.Tensor):
        arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
        capi.check(capi.lib().cb_comm_unpack(arr, len(fields), C.c_int64(dst_begin), C.c_int64(count),
                                             C.c_void_p(buf.data_ptr()), _stream()))

    _DTYPES = {torch.float64: 0, torch.float32: 1, torch.int32: 2, torch.int64: 3}

    def scatter_dtype(self, field) -> int:
        """Value types Cabana::scatter can sum (checked BEFORE any communication is posted)."""
        code = self._DTYPES.get(field.data.dtype)
        if code is None:
            raise TypeError(f"scatter: unsupported slice value type {field.data.dtype}")
        return code

    def scatter_add(self, field, steering: torch.Tensor, count: int, buf: torch.Tensor):
        fd = field.field_desc()
        capi.check(capi.lib().cb_comm_scatter_add_typed(
            C.byref(fd), C.c_void_p(steering.data_ptr()), C.c_int64(count), C.c_void_p(buf.data_ptr()),
            C.c_int(self.scatter_dtype(field)), _stream()))

    def pack_range(self, fields, src_begin: int, count: int, out: torch.Tensor):
        arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
        capi.check(capi.lib().cb_comm_pack_range(arr, len(fields), C.c_int64(src_begin), C.c_int64(count),
                                                 C.c_void_p(out.data_ptr()), _stream()))

    def slab_halo_select(self, x: Slice, num_local, lo_thresh, hi_thresh, lo_rank, hi_rank):
        ranks = torch.empty(2 * max(num_local, 1), dtype=torch.int32, device="cuda")
        ids = torch.empty(2 * max(num_local, 1), dtype=torch.int32, device="cuda")
        d = x.positions_desc()
        capi.check(capi.lib().cb_slab_halo_select(
            C.byref(d), C.c_int64(num_local), C.c_double(lo_thresh), C.c_double(hi_thresh),
            C.c_int(lo_rank), C.c_int(hi_rank), C.c_void_p(ranks.data_ptr()),
            C.c_void_p(ids.data_ptr()), _stream()))
        return ids[: 2 * num_local], ranks[: 2 * num_local]

    def slab_halo_plan(self, x: Slice, num_local, lo_thresh, hi_thresh, lo_rank, hi_rank):
        """Fused selection + stable compaction (one kernel, one read-back)."""
        steer = torch.empty(2 * max(num_local, 1), dtype=torch.int32, device="cuda")
        counts = (C.c_int64 * 2)()
        d = x.positions_desc()
        capi.check(capi.lib().cb_slab_halo_plan(
            C.byref(d), C.c_int64(num_local), C.c_double(lo_thresh), C.c_double(hi_thresh),
            C.c_int(1 if lo_rank >= 0 else 0), C.c_int(1 if hi_rank >= 0 else 0),
            C.c_void_p(steer.data_ptr()), C.c_void_p(steer.data_ptr() + 4 * max(num_local, 1)),
            counts, _stream()))
        return steer, int(counts[0]), int(counts[1])

    def slab_destinations(self, x: Slice, num_local, bounds):
        out = torch.empty(max(num_local, 1), dtype=torch.int32, device="cuda")
        d = x.positions_desc()
        b = (C.c_double * len(bounds))(*[float(v) for v in bounds])
        capi.check(capi.lib().cb_slab_migrate_destinations(
            C.byref(d), C.c_int64(num_local), b, C.c_int(len(bounds) - 1),
            C.c_void_p(out.data_ptr()), _stream()))
        return out[:num_local]


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class CommunicationPlan:
    """Export-built plan with the reference's neighbour ordering.

    Without a topology (impl/Cabana_CommunicationPlan_Mpi.hpp:270-410): the ranks we export to in
    ascending order, this rank swapped to the front when it sends to itself (:318-328), then the
    ranks we only import from (the reference appends them in message-arrival order; here
    ascending).  With a topology (`neighbor_ranks`, :105-192): getUniqueTopology
    (Cabana_CommunicationPlanBase.hpp:374-394) -- sorted, unique, this rank swapped with the
    first entry -- whether or not there is traffic.  Export blocks and import blocks are laid
    out in that neighbour order (:626-637; impl/Cabana_Halo_Mpi.hpp:73-101).
    """

    def __init__(self, export_ranks: torch.Tensor | None, export_ids: torch.Tensor | None = None,
                 group=None, kernels=None, plan=None, neighbor_ranks=None):
        self.kernels = kernels if kernels is not None else CudaCommKernels()
        self.group = group
        self.rank, self.world = _world(group)
        if plan is not None:
            # precomputed (counts per rank, block offsets into steering, steering)
            counts, offsets, steering = plan
        else:
            counts, offsets, steering = self.kernels.count_and_steer(export_ranks, export_ids, self.world)
        # counts exchange: one all_gather of the export-count vector (replaces the
        # per-neighbour MPI_Send of one unsigned long, impl/Cabana_CommunicationPlan_Mpi.hpp:152-178)
        mine = torch.tensor(counts, dtype=torch.int64, device=self.kernels.device)
        if self.world > 1:
            allc = torch.empty(self.world * self.world, dtype=torch.int64, device=self.kernels.device)
            dist.all_gather_into_tensor(allc, mine, group=group)
            allc = allc.view(self.world, self.world).cpu()
        else:
            allc = mine.view(1, 1).cpu()
        self.export_matrix = allc  # [src, dst]
        imports = [int(allc[s, self.rank]) for s in range(self.world)]
        if neighbor_ranks is not None:
            topo = sorted(set(int(r) for r in neighbor_ranks if int(r) >= 0))
            if self.rank in topo:
                i = topo.index(self.rank)
                topo[0], topo[i] = topo[i], topo[0]
            extra = [r for r in range(self.world) if r not in topo and (counts[r] > 0 or imports[r] > 0)]
            if extra:
                raise ValueError(f"CommunicationPlan: traffic with ranks {extra} outside the given topology")
            self.neighbors = topo
        else:
            nb = [r for r in range(self.world) if counts[r] > 0]
            if self.rank in nb:
                i = nb.index(self.rank)
                nb[0], nb[i] = nb[i], nb[0]
            nb += [r for r in range(self.world) if imports[r] > 0 and r not in nb]
            self.neighbors = nb
        self.num_export = [counts[r] for r in self.neighbors]
        self.num_import = [imports[r] for r in self.neighbors]
        # steering comes back grouped by ascending rank; re-express as neighbour-ordered blocks
        self._steer_rank_offsets = offsets
        self._steering = steering
        self.export_block = {r: (offsets[r], counts[r]) for r in self.neighbors}
        self.total_export = sum(self.num_export)
        self.total_import = sum(self.num_import)
        off = 0
        self.import_offset = {}
        for r, c in zip(self.neighbors, self.num_import):
            self.import_offset[r] = off
            off += c

    def numNeighbor(self):
        return len(self.neighbors)

    def neighborRank(self, n):
        return self.neighbors[n]

    def numExport(self, n):
        return self.num_export[n]

    def numImport(self, n):
        return self.num_import[n]

    def totalNumExport(self):
        return self.total_export

    def totalNumImport(self):
        return self.total_import

    def steering_for(self, r) -> torch.Tensor:
        o, c = self.export_block[r]
        return self._steering[o : o + c]

    # -- payload exchange: sends[r] / recvs[r] are byte tensors per neighbour rank
    def exchange(self, sends: dict, recvs: dict):
        ops = []
        for r in self.neighbors:
            if r == self.rank:
                continue
            if r in recvs and recvs[r].numel() > 0:
                ops.append(dist.P2POp(dist.irecv, recvs[r], r, group=self.group))
        for r in self.neighbors:
            if r == self.rank:
                continue
            if r in sends and sends[r].numel() > 0:
                ops.append(dist.P2POp(dist.isend, sends[r], r, group=self.group))
        if self.rank in sends and self.rank in recvs and sends[self.rank].numel() > 0:
            recvs[self.rank].copy_(sends[self.rank])  # self-send short-circuits to a copy
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


def exports_from_imports(import_ids: torch.Tensor, import_ranks: torch.Tensor, group=None,
                         device=None):
    """The Import build of a plan (CommunicationPlan::createWithTopology( Import, ... ) /
    createWithoutTopology( Import, ... ), impl/Cabana_CommunicationPlan_Mpi.hpp:480-960).

    Every rank lists what it wants: import_ids[i] is a LOCAL id on rank import_ranks[i].  The
    requests are delivered to their owners (counts with one all_gather, ids with grouped
    point-to-point -- the reference sends one MPI message per id) and come back as the
    (export_ids, export_ranks) pair an Export-built plan is constructed from.  Requests keep
    their order per (requester, owner) pair and are grouped by requester in ascending rank
    order, so after `gather` the ghosts of this rank sit in neighbour order and, inside one
    neighbour's block, in the order this rank asked for them.
    """
    rank, world = _world(group)
    dev = device if device is not None else import_ids.device
    ids = import_ids.to(torch.int32)
    ranks = import_ranks.to(torch.int64)
    if ranks.numel() and (int(ranks.min()) < 0 or int(ranks.max()) >= world):
        raise ValueError("an import rank of -1 (or out of range) is not supported")
    want = torch.bincount(ranks.cpu(), minlength=world).to(torch.int64)
    if world > 1:
        allw = torch.empty(world * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allw, want.to(dev), group=group)
        allw = allw.view(world, world).cpu()     # [requester, owner]
    else:
        allw = want.view(1, 1)
    order = torch.argsort(ranks.cpu(), stable=True).to(ids.device)
    sorted_ids = ids[order]
    starts = torch.cumsum(want, 0) - want
    sends = {r: sorted_ids[int(starts[r]): int(starts[r]) + int(want[r])].contiguous()
             for r in range(world) if int(want[r]) > 0}
    recvs = {r: torch.empty(int(allw[r, rank]), dtype=torch.int32, device=ids.device)
             for r in range(world) if int(allw[r, rank]) > 0}
    ops = [dist.P2POp(dist.irecv, recvs[r], r, group=group) for r in recvs if r != rank]
    ops += [dist.P2POp(dist.isend, sends[r], r, group=group) for r in sends if r != rank]
    if rank in sends:
        recvs[rank].copy_(sends[rank])
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    requesters = sorted(recvs)
    if requesters:
        export_ids = torch.cat([recvs[r] for r in requesters])
        export_ranks = torch.cat([torch.full((recvs[r].numel(),), r, dtype=torch.int32,
                                             device=ids.device) for r in requesters])
    else:
        export_ids = torch.empty(0, dtype=torch.int32, device=ids.device)
        export_ranks = torch.empty(0, dtype=torch.int32, device=ids.device)
    return export_ids, export_ranks


class Halo(CommunicationPlan):
    """Cabana::Halo<MemorySpace, Export|Import, Nccl> (core/src/Cabana_Halo.hpp:59-268)."""

    def __init__(self, num_local: int, export_ids: torch.Tensor | None,
                 export_ranks: torch.Tensor | None, group=None, kernels=None, plan=None,
                 neighbor_ranks=None):
        super().__init__(export_ranks, export_ids, group, kernels, plan, neighbor_ranks)
        self._num_local = int(num_local)

    @classmethod
    def from_imports(cls, num_local: int, import_ids: torch.Tensor, import_ranks: torch.Tensor,
                     group=None, kernels=None):
        """Halo<MemorySpace, Import> (Cabana_Halo.hpp:174-330): built from the ghosts this rank
        wants (local ids on their owners) instead of from what it sends."""
        k = kernels if kernels is not None else CudaCommKernels()
        export_ids, export_ranks = exports_from_imports(import_ids, import_ranks, group, k.device)
        return cls(num_local, export_ids, export_ranks, group, k)

    def numLocal(self):
        return self._num_local

    def numGhost(self):
        return self.total_import


def gather(halo: Halo, *fields: Slice):
    """Cabana::gather(halo, aosoa|slice) (Cabana_Halo.hpp:677-682; impl/Cabana_Halo_Mpi.hpp:41-125).

    Every field must hold numLocal() + numGhost() elements; ghosts land at
    [numLocal(), numLocal()+numGhost()) grouped by source rank in neighbour order.
    """
    k = halo.kernels
    tb = k.tuple_bytes(fields)
    sends, recvs = {}, {}
    for r, ne, ni in zip(halo.neighbors, halo.num_export, halo.num_import):
        if ne > 0:
            buf = torch.empty(ne * tb, dtype=torch.uint8, device=k.device)
            k.pack(fields, halo.steering_for(r), ne, buf)
            sends[r] = buf
        if ni > 0:
            recvs[r] = torch.empty(ni * tb, dtype=torch.uint8, device=k.device)
    halo.exchange(sends, recvs)
    for r, ni in zip(halo.neighbors, halo.num_import):
        if ni > 0:
            k.unpack(fields, halo.numLocal() + halo.import_offset[r], ni, recvs[r])


def scatter(halo: Halo, field: Slice):
    """Cabana::scatter(halo, slice) (impl/Cabana_Halo_Mpi.hpp:236-350): ghost values are sent
    back to their owners and atomically summed into them."""
    k = halo.kernels
    k.scatter_dtype(field)             # reject unsupported value types before posting anything
    tb = k.tuple_bytes([field])        # packed single-field tuples (4-byte types pad to 8 bytes)
    sends, recvs = {}, {}
    for r, ne, ni in zip(halo.neighbors, halo.num_export, halo.num_import):
        if ni > 0:
            buf = torch.empty(ni * tb, dtype=torch.uint8, device=k.device)
            k.pack_range([field], halo.numLocal() + halo.import_offset[r], ni, buf)
            sends[r] = buf
        if ne > 0:
            recvs[r] = torch.empty(ne * tb, dtype=torch.uint8, device=k.device)
    halo.exchange(sends, recvs)
    for r, ne in zip(halo.neighbors, halo.num_export):
        if ne > 0:
            k.scatter_add(field, halo.steering_for(r), ne, recvs[r])


class _CommunicationData:
    """Persistent send/receive buffers of a Gather / Scatter (CommunicationData,
    Cabana_CommunicationPlanBase.hpp:700-960): sizes in TUPLES, grow-only on reserve(),
    reallocated by shrinkToFit()."""

    def __init__(self, halo: Halo, fields, overallocation: float, total_send, total_recv):
        if overallocation < 1.0:
            raise RuntimeError("Cabana::CommunicationPlan: Cannot allocate buffers with less space "
                               "than data to communicate!")
        self._overallocation = float(overallocation)
        self._send = torch.empty(0, dtype=torch.uint8, device=halo.kernels.device)
        self._recv = torch.empty(0, dtype=torch.uint8, device=halo.kernels.device)
        self._reserve(halo, fields, total_send, total_recv)

    def _reserve(self, halo, fields, total_send, total_recv):
        self.halo, self.fields = halo, list(fields)
        self._tb = halo.kernels.tuple_bytes(self.fields)
        new_send = int(total_send * self._overallocation)
        if new_send > self.sendCapacity():
            self._send = torch.empty(new_send * self._tb, dtype=torch.uint8, device=halo.kernels.device)
        new_recv = int(total_recv * self._overallocation)
        if new_recv > self.receiveCapacity():
            self._recv = torch.empty(new_recv * self._tb, dtype=torch.uint8, device=halo.kernels.device)
        self._send_size, self._recv_size = int(total_send), int(total_recv)

    def sendSize(self):
        return self._send_size

    def receiveSize(self):
        return self._recv_size

    def sendCapacity(self):
        return self._send.numel() // max(self._tb, 1)

    def receiveCapacity(self):
        return self._recv.numel() // max(self._tb, 1)

    def shrinkToFit(self, use_overallocation: bool = False):
        f = self._overallocation if use_overallocation else 1.0
        dev = self.halo.kernels.device
        self._send = torch.empty(int(self._send_size * f) * self._tb, dtype=torch.uint8, device=dev)
        self._recv = torch.empty(int(self._recv_size * f) * self._tb, dtype=torch.uint8, device=dev)

    def _blocks(self, buf, counts):
        out, at = {}, 0
        for r, c in zip(self.halo.neighbors, counts):
            out[r] = buf[at * self._tb:(at + c) * self._tb]
            at += c
        return out


class Gather(_CommunicationData):
    """Cabana::Gather<HaloType, AoSoA|Slice> (Cabana_Halo.hpp:392-640): a gather with persistent
    buffers.  `apply()` = impl/Cabana_Halo_Mpi.hpp:41-125; reserve() re-targets the object at a
    new halo / data and only grows the buffers."""

    def __init__(self, halo: Halo, *fields: Slice, overallocation: float = 1.0):
        super().__init__(halo, fields, overallocation, halo.totalNumExport(), halo.totalNumImport())

    def reserve(self, halo: Halo, *fields: Slice, overallocation: float | None = None):
        if overallocation is not None:
            if overallocation < 1.0:
                raise RuntimeError("Cabana::CommunicationPlan: Cannot allocate buffers with less "
                                   "space than data to communicate!")
            self._overallocation = float(overallocation)
        self._reserve(halo, fields, halo.totalNumExport(), halo.totalNumImport())

    def apply(self):
        h, k = self.halo, self.halo.kernels
        sends = self._blocks(self._send, h.num_export)
        recvs = self._blocks(self._recv, h.num_import)
        for r, ne in zip(h.neighbors, h.num_export):
            if ne > 0:
                k.pack(self.fields, h.steering_for(r), ne, sends[r])
        h.exchange(sends, recvs)
        for r, ni in zip(h.neighbors, h.num_import):
            if ni > 0:
                k.unpack(self.fields, h.numLocal() + h.import_offset[r], ni, recvs[r])


class Scatter(_CommunicationData):
    """Cabana::Scatter<HaloType, Slice> (Cabana_Halo.hpp:700-870): ghost values go back to their
    owners and are summed into them; persistent buffers (send = ghosts, receive = exports)."""

    def __init__(self, halo: Halo, field: Slice, overallocation: float = 1.0):
        super().__init__(halo, [field], overallocation, halo.totalNumImport(), halo.totalNumExport())

    def reserve(self, halo: Halo, field: Slice, overallocation: float | None = None):
        if overallocation is not None:
            if overallocation < 1.0:
                raise RuntimeError("Cabana::CommunicationPlan: Cannot allocate buffers with less "
                                   "space than data to communicate!")
            self._overallocation = float(overallocation)
        self._reserve(halo, [field], halo.totalNumImport(), halo.totalNumExport())

    def apply(self):
        h, k = self.halo, self.halo.kernels
        field = self.fields[0]
        k.scatter_dtype(field)
        sends = self._blocks(self._send, h.num_import)
        recvs = self._blocks(self._recv, h.num_export)
        for r, ni in zip(h.neighbors, h.num_import):
            if ni > 0:
                k.pack_range([field], h.numLocal() + h.import_offset[r], ni, sends[r])
        h.exchange(sends, recvs)
        for r, ne in zip(h.neighbors, h.num_export):
            if ne > 0:
                k.scatter_add(field, h.steering_for(r), ne, recvs[r])


def createGather(halo: Halo, *fields: Slice, overallocation: float = 1.0) -> Gather:
    """Cabana::createGather (Cabana_Halo.hpp:653-664)."""
    return Gather(halo, *fields, overallocation=overallocation)


def createScatter(halo: Halo, field: Slice, overallocation: float = 1.0) -> Scatter:
    """Cabana::createScatter (Cabana_Halo.hpp:833-843)."""
    return Scatter(halo, field, overallocation)


class Distributor(CommunicationPlan):
    """Cabana::Distributor<MemorySpace, Nccl> (core/src/Cabana_Distributor.hpp:62-146).

    export_ranks[i] = destination rank of element i, -1 to drop it.
    """

    def __init__(self, export_ranks: torch.Tensor, group=None, kernels=None, neighbor_ranks=None):
        # neighbor_ranks: the topology constructor (Cabana_Distributor.hpp:103-122)
        super().__init__(export_ranks, None, group, kernels, neighbor_ranks=neighbor_ranks)


def migrate(distributor: Distributor, src_fields, dst_fields):
    """Cabana::migrate(distributor, src, dst) (Cabana_Distributor.hpp:330-337;
    impl/Cabana_Migrate_Mpi.hpp:41-177): dst holds totalNumImport() elements, staying
    elements first (self is neighbour 0), then one block per source rank."""
    k = distributor.kernels
    tb = k.tuple_bytes(src_fields)
    sends, recvs = {}, {}
    for r, ne, ni in zip(distributor.neighbors, distributor.num_export, distributor.num_import):
        if ne > 0:
            buf = torch.empty(ne * tb, dtype=torch.uint8, device=k.device)
            k.pack(src_fields, distributor.steering_for(r), ne, buf)
            sends[r] = buf
        if ni > 0:
            recvs[r] = torch.empty(ni * tb, dtype=torch.uint8, device=k.device)
    distributor.exchange(sends, recvs)
    for r, ni in zip(distributor.neighbors, distributor.num_import):
        if ni > 0:
            k.unpack(dst_fields, distributor.import_offset[r], ni, recvs[r])


# --------------------------------------------------------------------------------- slab domain
class SlabDecomposition:
    """1-D slab partition along x over the ranks of `group`.

    Rank g owns x in [bounds[g], bounds[g+1]) (the last rank owns its upper face).
    """

    def __init__(self, bounds, halo_width: float, group=None, kernels=None):
        self.kernels = kernels if kernels is not None else CudaCommKernels()
        self.group = group
        self.rank, self.world = _world(group)
        assert len(bounds) == self.world + 1
        self.bounds = [float(b) for b in bounds]
        # a hair wider than r so that rounding of lo + r can never lose a neighbour
        self.halo_width = float(halo_width) * (1.0 + 2.0**-40)
        self.lo = self.bounds[self.rank]
        self.hi = self.bounds[self.rank + 1]
        # a slab thinner than the halo would need ghosts from ranks two or more away, which the
        # nearest-neighbour topology below would silently lose
        for g in range(self.world):
            if self.world > 1 and self.bounds[g + 1] - self.bounds[g] < self.halo_width:
                raise ValueError(f"SlabDecomposition: slab {g} is thinner ({self.bounds[g + 1] - self.bounds[g]:g}) "
                                 f"than the halo width ({self.halo_width:g})")
        self.lo_rank = self.rank - 1 if self.rank > 0 else -1
        self.hi_rank = self.rank + 1 if self.rank < self.world - 1 else -1

    def local_grid_x(self):
        """x-extent of the local Verlet grid: the slab widened by the halo on interior faces."""
        gmin = self.lo - self.halo_width if self.lo_rank >= 0 else self.lo
        gmax = self.hi + self.halo_width if self.hi_rank >= 0 else self.hi
        return gmin, gmax

    def create_halo(self, x: Slice, num_local: int) -> Halo:
        lo_t, hi_t = self.lo + self.halo_width, self.hi - self.halo_width
        if hasattr(self.kernels, "slab_halo_plan"):
            # fused plan: one kernel + one read-back; same plan as the general path below
            steer, n_lo, n_hi = self.kernels.slab_halo_plan(x, num_local, lo_t, hi_t,
                                                            self.lo_rank, self.hi_rank)
            counts = [0] * self.world
            offsets = [0] * (self.world + 1)
            if self.lo_rank >= 0:
                counts[self.lo_rank] = n_lo
                offsets[self.lo_rank] = 0
            if self.hi_rank >= 0:
                counts[self.hi_rank] = n_hi
                offsets[self.hi_rank] = max(num_local, 1)
            return Halo(num_local, None, None, self.group, self.kernels, plan=(counts, offsets, steer),
                        neighbor_ranks=self.topology())
        ids, ranks = self.kernels.slab_halo_select(x, num_local, lo_t, hi_t, self.lo_rank, self.hi_rank)
        return Halo(num_local, ids, ranks, self.group, self.kernels, neighbor_ranks=self.topology())

    def topology(self):
        """The known point-to-point topology of the slab halo (neighbour ranks incl. this one), as
        Cabana::Grid passes it to Halo's topology constructor."""
        return [r for r in (self.lo_rank, self.rank, self.hi_rank) if r >= 0]

    def create_distributor(self, x: Slice, num_local: int) -> Distributor:
        dest = self.kernels.slab_destinations(x, num_local, self.bounds)
        return Distributor(dest, self.group, self.kernels)

    def create_peer_halo(self, fields, capacity: int) -> "PeerHalo":
        """Peer-memory halo (CUDA IPC windows over NVLink) for `fields`-shaped tuples."""
        return PeerHalo(self, fields, capacity)


class PeerHalo:
    """Ghost gather for the slab decomposition through peer memory instead of send/recv.

    Each rank owns two receive windows (from its lower and upper neighbour) that the
    neighbours map with CUDA IPC once, at construction.  `gather` then runs, with no NCCL
    call and ONE host synchronisation:
        fused plan + pack straight into the neighbours' HBM  (cb_slab_halo_push)
        bounded device-side wait for the neighbours' pushes    (cb_slab_halo_wait)
        unpack from the local windows                          (cb_comm_unpack)
    and produces exactly the ghosts of `gather(slab.create_halo(...))`, in the same order
    (lower neighbour's first).  Replaces impl/Cabana_Halo_Mpi.hpp:41-125 for this topology.
    All ranks of the group must be processes on one NVLink/NVSwitch box.
    """

    def __init__(self, slab: SlabDecomposition, fields, capacity: int):
        self.slab = slab
        L = capi.lib()
        self.capacity = int(capacity)
        if slab.world > 1:
            # a pusher addresses the neighbour's window with ITS OWN idea of the layout, so every
            # rank must use the same capacity: the largest one asked for
            t = torch.tensor([self.capacity], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=slab.group)
            self.capacity = int(t.item())
        arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
        self.tuple_bytes = int(L.cb_comm_tuple_bytes(arr, len(fields)))
        self._win = [C.c_void_p(), C.c_void_p()]      # from_lo, from_hi
        self._peer = [C.c_void_p(), C.c_void_p()]     # where I push my lo / hi face
        handles, err = [], None
        try:
            for w in self._win:
                capi.check(L.cb_p2p_window_create(C.byref(w), C.c_int64(self.capacity),
                                                  C.c_int64(self.tuple_bytes)))
                h = (C.c_ubyte * 64)()
                capi.check(L.cb_p2p_window_get_handle(w, h))
                handles.append(bytes(h))
        except Exception as e:
            handles, err = None, str(e)
        everyone = [None] * slab.world
        dist.all_gather_object(everyone, handles, group=slab.group)
        try:
            if err is None and any(h is None for h in everyone):
                err = "a neighbour could not export its window"
            if err is None and slab.lo_rank >= 0:   # my low face lands in the lower neighbour's "from_hi" window
                h = (C.c_ubyte * 64).from_buffer_copy(everyone[slab.lo_rank][1])
                capi.check(L.cb_p2p_window_open(h, C.byref(self._peer[0])))
            if err is None and slab.hi_rank >= 0:
                h = (C.c_ubyte * 64).from_buffer_copy(everyone[slab.hi_rank][0])
                capi.check(L.cb_p2p_window_open(h, C.byref(self._peer[1])))
        except Exception as e:   # e.g. CUDA IPC not permitted between these processes
            err = str(e)
        # all ranks agree on the outcome before anyone pushes (and nobody is left in a barrier)
        errs = [None] * slab.world
        dist.all_gather_object(errs, err, group=slab.group)
        self._seq = 0
        self._steer = None
        if any(e is not None for e in errs):
            self._release()
            raise RuntimeError("PeerHalo: peer-memory windows unavailable: " +
                               "; ".join(f"rank {r}: {e}" for r, e in enumerate(errs) if e))
        dist.barrier(group=slab.group)   # every window is mapped before anyone pushes

    def gather(self, x: Slice, fields, num_local: int):
        """Push my face layers, receive the neighbours'; ghosts land at [num_local, ...).

        `x` is the position slice the selection reads (its first num_local tuples), `fields`
        the slices to ship (normally including x) sized for the ghosts.  Returns
        (num_ghost_from_lo, num_ghost_from_hi)."""
        L = capi.lib()
        s = self.slab
        self._seq += 1
        if self._steer is None or self._steer.numel() < 2 * max(num_local, 1):
            self._steer = torch.empty(2 * max(num_local, 1), dtype=torch.int32, device="cuda")
        arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
        d = x.positions_desc()
        st = _stream()
        capi.check(L.cb_slab_halo_push(
            C.byref(d), arr, len(fields), C.c_int64(num_local),
            C.c_double(s.lo + s.halo_width), C.c_double(s.hi - s.halo_width),
            self._peer[0], self._peer[1], C.c_int64(self.capacity), C.c_uint64(self._seq),
            C.c_void_p(self._steer.data_ptr()), st))
        counts = (C.c_int64 * 2)()
        data = [C.c_void_p(), C.c_void_p()]
        capi.check(L.cb_slab_halo_wait(
            self._win[0] if s.lo_rank >= 0 else None, self._win[1] if s.hi_rank >= 0 else None,
            C.c_uint64(self._seq), counts, C.byref(data[0]), C.byref(data[1]), st))
        n_lo, n_hi = int(counts[0]), int(counts[1])
        need = num_local + n_lo + n_hi
        for f in fields:
            if f.n < need:
                raise ValueError(f"PeerHalo.gather: slice holds {f.n} tuples, needs {need}")
        if n_lo:
            capi.check(L.cb_comm_unpack(arr, len(fields), C.c_int64(num_local), C.c_int64(n_lo),
                                        data[0], st))
        if n_hi:
            capi.check(L.cb_comm_unpack(arr, len(fields), C.c_int64(num_local + n_lo),
                                        C.c_int64(n_hi), data[1], st))
        return n_lo, n_hi

    def step(self, lst, x_all: Slice, fields, num_local: int, neighborhood_radius, cell_size_ratio,
             grid_min, grid_max, max_neigh=0):
        """gather() followed by lst.build( x[0:num_local+ghosts], 0, num_local, ... ) in ONE host
        call (cb_slab_step): push, wait, unpack and the owner-local build are issued from C.
        `x_all` / `fields` are sized for owned + ghost tuples.  Returns (n_lo, n_hi)."""
        L = capi.lib()
        s = self.slab
        self._seq += 1
        if self._steer is None or self._steer.numel() < 2 * max(num_local, 1):
            self._steer = torch.empty(2 * max(num_local, 1), dtype=torch.int32, device="cuda")
        arr = (capi.Field * len(fields))(*[f.field_desc() for f in fields])
        d = x_all.positions_desc()
        counts = (C.c_int64 * 2)()
        capi.check(L.cb_slab_step(
            lst._h, C.byref(d), arr, len(fields), C.c_int64(num_local),
            C.c_double(s.lo + s.halo_width), C.c_double(s.hi - s.halo_width),
            self._peer[0], self._peer[1],
            self._win[0] if s.lo_rank >= 0 else None, self._win[1] if s.hi_rank >= 0 else None,
            C.c_int64(self.capacity), C.c_uint64(self._seq), C.c_void_p(self._steer.data_ptr()),
            C.c_double(neighborhood_radius), C.c_double(cell_size_ratio), capi.d3(grid_min),
            capi.d3(grid_max), C.c_int64(max_neigh), C.c_int(lst.algorithm), C.c_int(lst.layout),
            C.c_int(lst.build_tag), counts, _stream()))
        lst._refresh()
        return int(counts[0]), int(counts[1])

    def close(self):
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.slab.group)
        self._release()

    def _release(self):
        L = capi.lib()
        for p in self._peer:
            if p:
                L.cb_p2p_window_close(p)
        for w in self._win:
            if w:
                L.cb_p2p_window_destroy(w)
        self._peer = [C.c_void_p(), C.c_void_p()]
        self._win = [C.c_void_p(), C.c_void_p()]
