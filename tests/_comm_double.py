"""TEST DOUBLE for cabana_b200.comm.CudaCommKernels, implemented with torch CPU ops.

It exists only so the host-side plan logic (count exchange, neighbour ordering, buffer
layout, ghost placement) can be exercised with world_size-2 gloo process groups on a
CPU-only box.  It lives in tests/ on purpose: the package ships no CPU implementation.
The semantics restate csrc/cb_comm.cu / the reference:
  count_and_steer   Cabana_CommunicationPlanBase.hpp:96-224, :596-657 (deterministic order)
  pack / unpack     impl/Cabana_Halo_Mpi.hpp:58-65, :113-121
  scatter_add       impl/Cabana_Halo_Mpi.hpp:334-347
"""
import torch


class CpuSlice:
    """Dense (n,k) CPU field standing in for cabana_b200.core.Slice."""

    def __init__(self, t: torch.Tensor):
        self.t = t
        self.n = t.shape[0]
        self.num_comp = t.shape[1]

    @property
    def data(self):
        return self.t

    def size(self):
        return self.n

    def to_array(self):
        return self.t


class CpuCommKernels:
    device = "cpu"

    def count_and_steer(self, export_ranks, export_ids, num_ranks):
        counts = [int((export_ranks == r).sum()) for r in range(num_ranks)]
        offsets = [0]
        for c in counts:
            offsets.append(offsets[-1] + c)
        idx = torch.arange(export_ranks.numel())
        steer = []
        for r in range(num_ranks):
            sel = idx[export_ranks == r]
            steer.append(export_ids[sel].to(torch.int64) if export_ids is not None else sel)
        steering = torch.cat(steer) if steer else torch.zeros(0, dtype=torch.int64)
        return counts, offsets, steering

    def tuple_bytes(self, fields):
        return sum(f.num_comp * f.t.element_size() for f in fields)

    def pack(self, fields, steering, count, out):
        cols = [f.t[steering.long()].contiguous().view(torch.uint8).reshape(count, -1) for f in fields]
        out.copy_(torch.cat(cols, dim=1).reshape(-1))

    def unpack(self, fields, dst_begin, count, buf):
        rows = buf.reshape(count, -1)
        at = 0
        for f in fields:
            w = f.num_comp * f.t.element_size()
            f.t[dst_begin : dst_begin + count] = rows[:, at : at + w].contiguous().view(f.t.dtype).reshape(count, f.num_comp)
            at += w

    def scatter_dtype(self, field):
        if field.t.dtype not in (torch.float64, torch.float32, torch.int32, torch.int64):
            raise TypeError(f"scatter: unsupported slice value type {field.t.dtype}")
        return 0

    def scatter_add(self, field, steering, count, buf):
        vals = buf.view(field.t.dtype).reshape(count, field.num_comp)
        field.t.index_add_(0, steering.long(), vals)

    def pack_range(self, fields, src_begin, count, out):
        cols = [f.t[src_begin:src_begin + count].contiguous().view(torch.uint8).reshape(count, -1) for f in fields]
        out.copy_(torch.cat(cols, dim=1).reshape(-1))

    def slab_halo_select(self, x, num_local, lo_thresh, hi_thresh, lo_rank, hi_rank):
        px = x.t[:num_local, 0]
        ranks = torch.full((2 * num_local,), -1, dtype=torch.int32)
        if lo_rank >= 0:
            ranks[0::2] = torch.where(px < lo_thresh, lo_rank, -1).to(torch.int32)
        if hi_rank >= 0:
            ranks[1::2] = torch.where(px >= hi_thresh, hi_rank, -1).to(torch.int32)
        ids = torch.arange(num_local, dtype=torch.int32).repeat_interleave(2)
        return ids, ranks

    def slab_destinations(self, x, num_local, bounds):
        px = x.t[:num_local, 0]
        b = torch.tensor(bounds, dtype=torch.float64)
        nr = len(bounds) - 1
        dest = torch.full((num_local,), nr - 1, dtype=torch.int32)
        for g in range(nr - 2, -1, -1):
            dest = torch.where(px < b[g + 1], torch.tensor(g, dtype=torch.int32), dest)
        dest = torch.where((px >= b[0]) & (px <= b[nr]), dest, torch.tensor(-1, dtype=torch.int32))
        return dest
