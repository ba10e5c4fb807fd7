#!/usr/bin/env python
"""Halo / Distributor timings shaped like the reference's benchmark/core/Cabana_CommPerformance.cpp
(:31-415): particles with members (double[3], double[3], double, int); a fraction of them is sent
to the neighbour ranks; timed over 10 runs each: distributor create (with topology = "fast",
without = "general"), AoSoA migrate (all members in one tuple), slice migrate (one member), halo
create, AoSoA gather, slice gather, slice scatter.  The rank topology is the one the hot path
uses: a 1-D ring of slabs (self + lower + upper neighbour), periodic like the reference's.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/comm_bench.py [--particles 1000000] [--runs 10]

Prints one JSON object (rank 0): times in milliseconds (device time, CUDA events, max over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    # libraries (NCCL banner) must not pollute stdout: only the JSON line goes there
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--runs", type=int, default=10)
    ap.add_argument("--fractions", type=float, nargs="*", default=[0.0001, 0.001, 0.01, 0.1, 0.5])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    from cabana_b200 import comm
    from cabana_b200 import core as cb

    n = args.particles
    lo, hi = (rank - 1) % world, (rank + 1) % world
    topo = sorted({rank, lo, hi})
    n_other = len(topo) - 1
    rng = np.random.Generator(np.random.Philox(key=77 + rank))

    def members(count):
        return [cb.slice_from_array(rng.random((count, 3)), vlen=32),
                cb.slice_from_array(rng.random((count, 3)), vlen=32),
                cb.slice_from_array(rng.random((count, 1)), vlen=32),
                cb.slice_from_array(np.arange(count, dtype=np.int32).reshape(-1, 1), vlen=32)]

    def timed(fn):
        ts = []
        for _ in range(args.runs + 2):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor(ts[2:], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.cpu().numpy()
        return {"min": float(t.min()), "max": float(t.max()), "avg": float(t.mean())}, out

    results = {"world": world, "particles_per_rank": n, "runs": args.runs,
               "bytes_per_particle": 64, "fractions": {}}
    src = members(n)
    for frac in args.fractions:
        per_nb = int(n * frac) // max(n_other, 1) if n_other else 0
        num_send = per_nb * n_other
        ranks = np.full(n, rank, dtype=np.int32)
        for k, r in enumerate([t for t in topo if t != rank]):
            ranks[k * per_nb:(k + 1) * per_nb] = r
        d_ranks = torch.from_numpy(ranks).cuda()
        out = {"comm_bytes_per_neighbor": per_nb * 64}
        out["distributor_fast_create"], dfast = timed(
            lambda: comm.Distributor(d_ranks, neighbor_ranks=topo))
        out["distributor_general_create"], dgen = timed(lambda: comm.Distributor(d_ranks))
        n_in = dgen.totalNumImport()
        dst = members(n_in)
        out["distributor_aosoa_migrate"], _ = timed(lambda: comm.migrate(dgen, src, dst))
        out["distributor_slice_migrate"], _ = timed(lambda: comm.migrate(dgen, src[:1], dst[:1]))
        # halo: the same elements are ghosted on the neighbours
        ids = torch.arange(num_send, dtype=torch.int32, device="cuda")
        hr = d_ranks[:num_send].contiguous()
        out["halo_fast_create"], halo = timed(lambda: comm.Halo(n, ids, hr, neighbor_ranks=topo))
        out["halo_general_create"], _ = timed(lambda: comm.Halo(n, ids, hr))
        n_tot = halo.numLocal() + halo.numGhost()
        big = members(n_tot)
        out["halo_aosoa_gather"], _ = timed(lambda: comm.gather(halo, *big))
        out["halo_slice_gather"], _ = timed(lambda: comm.gather(halo, big[0]))
        out["halo_slice_scatter"], _ = timed(lambda: comm.scatter(halo, big[0]))
        results["fractions"][str(frac)] = out
        del dst, big
    if rank == 0:
        os.write(real_stdout, (json.dumps(results) + "\n").encode())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
