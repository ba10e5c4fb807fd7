#!/usr/bin/env python
"""Headline benchmark: Verlet-build neighbours/second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one complete VerletList build -- LinkedCellList binning, gather-permute,
count pass, offsets scan, fill pass, including the allocation-size read-back -- of the
16 078 716-atom FCC Lennard-Jones configuration (BASELINE.json configs[2]:
FullNeighborTag, VerletLayoutCSR, cutoff 2.5 sigma + 0.3 sigma skin), which fits one B200.
At N > 1 the SAME configuration is slab-decomposed along x over N GPUs (strong scaling):
every step then also selects the ghost layer, builds the Halo plan and gathers ghost
positions over NCCL before the local build.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is defined.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "verlet_build_neighbors_per_sec"
UNIT = "neighbors/s"
FCC_CELLS = 159          # 4 * 159^3 = 16 078 716 atoms
RADIUS = 2.8             # 2.5 sigma cutoff + 0.3 sigma skin
CELL_RATIO = 1.0
CPU_SAMPLE_CELLS = int(os.environ.get("CB_BENCH_CPU_CELLS", "100"))  # cpu_baseline sample: 4 * 100^3 atoms


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def _workload_config(n_gpus):
    return {
        "workload": "cfg3: 16 078 716-atom FCC LJ (rho*=0.8442), r = 2.8 sigma, cell_size_ratio 1.0, "
                    "FullNeighborTag, VerletLayoutCSR" + ("" if n_gpus == 1 else
                    f", x-slab decomposition over {n_gpus} GPUs with Halo ghost gather each step "
                    f"({os.environ.get('CB_BENCH_HALO', 'peer')} halo)"),
        "particles": 4 * FCC_CELLS**3,
        "radius": RADIUS,
        "cell_size_ratio": CELL_RATIO,
        "algorithm": "FullNeighborTag",
        "layout": "VerletLayoutCSR",
        "positions": "AoSoA<double[3]> slice, VectorLength 32",
        "l2_policy": "inputs_exceed_l2 (386 MB of positions + 5 GB of output per step vs 126 MB L2)",
    }


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []          # (arrival time, csv row)
        self.proc = None
        self.index = index
        self.t_begin = None     # timed region, perf_counter clock
        self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # the sampler runs from before the warm-up; keep the samples that fell inside the timed
        # region, or -- if the region was shorter than the sampling period -- the ones taken
        # under the identical warm-up load just before it
        window = "timed region"
        rows = [r for t, r in self.rows
                if self.t_begin is None or (self.t_begin <= t <= (self.t_end or t) + 0.03)]
        if not rows:
            window = "warm-up + timed region (timed region shorter than the sampling period)"
            rows = [r for _, r in self.rows[-8:]]
        for row in rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "samples": len(sm),
            "window": window,
            "reasons": sorted(reasons),
        }


# ----------------------------------------------------------------------------------- CPU arm
def _cpu_setup():
    """The timed CPU legs use the -march=native copy of the oracle (built on the host that runs
    it) and EVERY host core: torch.distributed.run exports OMP_NUM_THREADS=1, which the oracle
    must not inherit."""
    import oracle

    native = oracle.use_native_build()
    threads = oracle.use_all_cores()
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert threads > 1 or avail == 1, f"CPU arm would run on {threads} of {avail} cores"
    return oracle, threads, native


def cpu_reference_run(steps: int, warmup: int, cells: int = CPU_SAMPLE_CELLS):
    """Time the CPU port of the reference path (oracle/) on an FCC lattice of `cells`^3 unit
    cells (cells = 159 is the full cfg3 workload).  Returns the last build for parity checks."""
    oracle, threads, native = _cpu_setup()
    from cabana_b200 import datasets

    ps = datasets.fcc_lattice(cells)
    x = oracle.slice_from_xyz(ps.xyz, vlen=16)  # host AoSoA vector length (PerformanceTraits)
    times, res = [], None
    for it in range(warmup + steps):
        res = None   # (one 5 GB list at a time)
        t0 = time.perf_counter()
        res = oracle.verlet_build(x, 0, ps.n, ps.radius, CELL_RATIO, ps.grid_min, ps.grid_max,
                                  algo=oracle.FULL, layout=oracle.CSR)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    avg = float(np.mean(times))
    what = ("the full workload" if cells == FCC_CELLS else
            f"{ps.n}-atom FCC sub-lattice ({cells}^3 cells) of the same workload")
    return {
        "value": res.total / avg,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"{what}, {ps.n} atoms, {len(times)} timed builds, avg {avg*1e3:.1f} ms "
                  f"(min {min(times)*1e3:.1f}, max {max(times)*1e3:.1f}); C++/OpenMP restatement of the "
                  "reference (Kokkos is not installable here), -O3 "
                  + ("-march=native " if native else "") + "-ffp-contract=off, dynamic schedule over cells",
    }, avg, res, ps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # the configuration the GPU arm names, in full (about 5 s per build on 16 cores);
    # CB_BENCH_REF_CELLS=<cells per side> shrinks it and the line then says so
    cells = int(os.environ.get("CB_BENCH_REF_CELLS", str(FCC_CELLS)))
    steps = max(1, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 1))
    cpu, avg, res, ps = cpu_reference_run(steps, warmup, cells)
    cfg = _workload_config(args.gpus)
    cfg["particles"] = int(ps.n)
    if cells != FCC_CELLS:
        cfg["workload"] += f" -- CPU arm ran a {ps.n}-atom sub-lattice ({cells}^3 cells) of it"
        cfg["same_config"] = False
    if args.gpus > 1:
        cfg["workload"] += " (the CPU arm builds the undecomposed box on rank 0's host cores)"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": avg * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "neighbors_per_step": float(res.total),
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


# ----------------------------------------------------------------------------------- GPU arm
def _fcc_slab(rank, world):
    """Owned atoms of this rank: lattice cells [c0,c1) along x; bounds at multiples of a."""
    from cabana_b200 import datasets

    a = (4.0 / datasets.FCC_DENSITY) ** (1.0 / 3.0)
    cuts = [round(FCC_CELLS * g / world) for g in range(world + 1)]
    bounds = [c * a for c in cuts]
    bounds[-1] = FCC_CELLS * a
    c0, c1 = cuts[rank], cuts[rank + 1]
    ps = datasets.fcc_lattice(c1 - c0, radius=RADIUS, cells_yz=FCC_CELLS)
    xyz = ps.xyz
    xyz[:, 0] += c0 * a
    return xyz, bounds, (FCC_CELLS * a,) * 3


def _fcc_first_id(rank, world):
    """Global index of a slab's first atom (lattice order: x cell slowest)."""
    cuts = [round(FCC_CELLS * g / world) for g in range(world + 1)]
    return cuts[rank] * 4 * FCC_CELLS * FCC_CELLS


# ----------------------------------------------------------------------------------- parity
def _lsr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def _mix64(z):
    """splitmix64 finaliser on int64 tensors (two's-complement wrap-around == uint64 arithmetic)."""
    z = z + (-7046029254386353131)            # 0x9e3779b97f4a7c15
    z = (z ^ _lsr(z, 30)) * (-4658895280553007687)   # 0xbf58476d1ce4e5b9
    z = (z ^ _lsr(z, 27)) * (-7723592293110705685)   # 0x94d049bb133111eb
    return z ^ _lsr(z, 31)


def _row_hashes_gpu(torch, counts, offsets, neighbors, row0, row1, idmap=None, chunk=1 << 21):
    """Order-independent 64-bit hash of rows [row0,row1) of a CSR list on the device: sum over the
    row of mix64(global id).  Chunked so the temporaries stay small."""
    out = torch.empty(row1 - row0, dtype=torch.int64, device=counts.device)
    for r0 in range(row0, row1, chunk):
        r1 = min(r0 + chunk, row1)
        cnt = counts[r0:r1].to(torch.int64)
        off = offsets[r0:r1].to(torch.int64)
        lo = int(off[0])
        hi = int(off[-1] + cnt[-1])
        ids = neighbors[lo:hi].to(torch.int64)
        if idmap is not None:
            ids = idmap[ids]
        cs = torch.cumsum(_mix64(ids), 0)
        cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=cs.device), cs])
        out[r0 - row0:r1 - row0] = cs[off - lo + cnt] - cs[off - lo]
    return out




def run_cfg5(args):
    """cfg5 (BASELINE.json configs[4], informational): WEAK scaling.  8 M uniform particles per
    GPU (rho = 0.455, r = 3, Half CSR); every step = ballistic drift (synthetic stand-in for the
    integrator: ~0.4 % of the particles change slab), Distributor/migrate over NCCL to the new
    owners, ghost gather through the peer-memory halo, owner-local Half CSR build."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    from cabana_b200 import build as cb_build
    if rank == 0:
        cb_build.build()
    dist.barrier()
    from cabana_b200 import capi, comm
    from cabana_b200 import core as cb

    Lc = capi.lib()
    n_per = int(os.environ.get("CB_CFG5_N", "8000000"))
    radius = 3.0
    Lbox = 1.3 * float(n_per) ** (1.0 / 3.0)
    bounds = [Lbox * g for g in range(world + 1)]
    rng = np.random.Generator(np.random.Philox(key=20240105 + rank))
    xyz = rng.random((n_per, 3)) * Lbox
    xyz[:, 0] += bounds[rank]
    vel = rng.normal(0.0, 1.0, (n_per, 3))
    cap = int(n_per * 1.25) + 4096
    slab = comm.SlabDecomposition(bounds, radius)
    lgx = slab.local_grid_x()
    lmin, lmax = (lgx[0], 0.0, 0.0), (lgx[1], Lbox, Lbox)

    def alloc(a):
        buf = np.zeros((cap, 3))
        buf[: a.shape[0]] = a
        return cb.slice_from_array(buf, vlen=32)

    xs = [alloc(xyz), alloc(xyz[:1])]      # double-buffered: migrate is out of place
    vs = [alloc(vel), alloc(vel[:1])]
    state = {"cur": 0, "n": n_per}
    peer = slab.create_peer_halo([xs[0]], cap // 8)
    lst = cb.VerletList(algorithm=cb.HALF, layout=cb.CSR)
    hi_global = bounds[-1]
    DT = 1.0    # displacement = DT * v, v ~ N(0,1)

    def sub(sl, n):
        return cb.Slice(sl.data, n, sl.outer_stride, sl.vlen, sl.comp_stride, 3)

    def drift(x, v, n):
        # synthetic integrator (plumbing, torch ops): x += DT v, reflected at the global box
        nso = (n + 31) // 32
        X = x.data[: nso * 96].view(nso, 3, 32)
        V = v.data[: nso * 96].view(nso, 3, 32)
        X.add_(V, alpha=DT)
        for d, hi in ((0, hi_global), (1, Lbox), (2, Lbox)):
            xd, vd = X[:, d, :], V[:, d, :]
            lo_m = xd < 0.0
            hi_m = xd >= hi
            xd.copy_(torch.where(lo_m, -xd, torch.where(hi_m, 2.0 * hi - xd, xd)))
            xd.clamp_(0.0, float(np.nextafter(hi, 0.0)))
            vd.copy_(torch.where(lo_m | hi_m, -vd, vd))

    def step():
        c, n = state["cur"], state["n"]
        x, v = xs[c], vs[c]
        drift(x, v, n)
        distributor = slab.create_distributor(sub(x, n), n)
        n_new = distributor.totalNumImport()
        assert n_new <= int(n_per * 1.1), "slab over capacity"
        comm.migrate(distributor, [sub(x, n), sub(v, n)], [sub(xs[1 - c], n_new), sub(vs[1 - c], n_new)])
        x2 = xs[1 - c]
        n_lo, n_hi = peer.gather(sub(x2, n_new), [x2], n_new)
        n_tot = n_new + n_lo + n_hi
        lst.build(sub(x2, n_tot), 0, n_new, radius, 1.0, lmin, lmax)
        state["cur"], state["n"] = 1 - c, n_new
        return lst.total, n - distributor.numExport(0)

    def sync_all():
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    launches0 = Lc.cb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    ev0.record()
    tot_sum, moved_sum = 0.0, 0.0
    for _ in range(args.steps):
        t_, m_ = step()
        tot_sum += t_
        moved_sum += m_
    ev1.record()
    sync_all()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = Lc.cb_kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, tot_sum / args.steps, moved_sum / args.steps, float(launches)],
                     dtype=torch.float64, device="cuda")
    tmax, tsum = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    peer.close()
    if rank == 0:
        ms_g = float(tmax[0])
        _emit({
            "metric": METRIC, "value": float(tsum[1]) / (ms_g * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_g,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "cfg5 (informational): %d uniform particles per GPU, rho 0.455, r = 3, "
                                   "HalfNeighborTag CSR; step = drift + Distributor migrate (NCCL) + "
                                   "peer-memory halo + build" % n_per,
                       "particles": n_per * world, "radius": radius,
                       "migrated_per_step": float(tsum[2])},
            "neighbors_per_step": float(tsum[1]), "gpu_launches": int(tsum[3]), "clocks": clocks,
        })
    dist.destroy_process_group()
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from cabana_b200 import build as cb_build
    if rank == 0:
        cb_build.build()
    if world > 1:
        dist.barrier()
    from cabana_b200 import capi, comm
    from cabana_b200 import core as cb
    import ctypes as C

    L = capi.lib()
    algo, layout, radius = cb.FULL, cb.CSR, RADIUS
    if args.workload == "cfg3":
        xyz, bounds, gmax = _fcc_slab(rank, world)
    else:
        assert world == 1, "only cfg3 is sharded"
        from cabana_b200 import datasets
        if args.workload == "cfg1":
            ps = datasets.uniform_box(100_000, 20240101)
        elif args.workload == "cfg2":
            ps = datasets.uniform_box(1_000_000, 20240102)
            algo, layout = cb.HALF, cb.LAYOUT_2D
        else:
            ps = datasets.clustered(8_000_000)
        xyz, gmax, radius, bounds = ps.xyz, ps.grid_max, ps.radius, None
    num_local = xyz.shape[0]
    gmin = (0.0, 0.0, 0.0)

    lst = cb.VerletList(algorithm=algo, layout=layout)
    capi.check(L.cb_verlet_set_profiling(lst._h, 1))

    if world == 1:
        x = cb.slice_from_array(xyz, vlen=32)

        def step():
            lst.build(x, 0, num_local, radius, CELL_RATIO, gmin, gmax)
            return lst.total
    else:
        slab = comm.SlabDecomposition(bounds, RADIUS)
        lgx = slab.local_grid_x()
        lmin = (lgx[0], 0.0, 0.0)
        lmax = (lgx[1], gmax[1], gmax[2])
        # capacity for owned + ghosts: two faces of thickness ~r
        cap = num_local + int(2.2 * RADIUS / (bounds[rank + 1] - bounds[rank]) * num_local) + 1024
        store = np.zeros((cap, 3))
        store[:num_local] = xyz
        x_all = cb.slice_from_array(store, vlen=32)

        trace = os.environ.get("CB_BENCH_TRACE") == "1"
        tr = {"plan": 0.0, "gather": 0.0, "build": 0.0, "n": 0}
        # ghost exchange: peer-memory windows over NVLink (default) or NCCL send/recv
        halo_impl = os.environ.get("CB_BENCH_HALO", "peer")
        peer = None
        if halo_impl == "peer":
            try:
                peer = slab.create_peer_halo([x_all], cap - num_local)
            except RuntimeError as e:   # raised on EVERY rank (the outcome is agreed collectively)
                if rank == 0:
                    sys.stderr.write(f"bench.py: {e}; falling back to the NCCL send/recv halo\n")
                halo_impl = "nccl"
                os.environ["CB_BENCH_HALO"] = "nccl"   # (named in config.workload)

        def step():
            t0 = time.perf_counter()
            x_own = cb.Slice(x_all.data, num_local, x_all.outer_stride, x_all.vlen, x_all.comp_stride, 3)
            if peer is not None and not trace:
                # the whole step from one C entry: push -> wait -> unpack -> owner-local build
                peer.step(lst, x_all, [x_all], num_local, RADIUS, CELL_RATIO, lmin, lmax)
                return lst.total
            if peer is not None:
                n_lo, n_hi = peer.gather(x_own, [x_all], num_local)
                n_tot = num_local + n_lo + n_hi
                x_tot = cb.Slice(x_all.data, n_tot, x_all.outer_stride, x_all.vlen, x_all.comp_stride, 3)
                if trace:
                    torch.cuda.synchronize()
                    t1 = time.perf_counter()
            else:
                halo = slab.create_halo(x_own, num_local)
                n_tot = halo.numLocal() + halo.numGhost()
                assert n_tot <= cap, "ghost capacity exceeded"
                x_tot = cb.Slice(x_all.data, n_tot, x_all.outer_stride, x_all.vlen, x_all.comp_stride, 3)
                if trace:
                    torch.cuda.synchronize()
                    t1 = time.perf_counter()
                comm.gather(halo, x_tot)
            if trace:
                torch.cuda.synchronize()
                t2 = time.perf_counter()
            lst.build(x_tot, 0, num_local, RADIUS, CELL_RATIO, lmin, lmax)
            if trace:
                torch.cuda.synchronize()
                t3 = time.perf_counter()
                tr["plan"] += t1 - t0
                tr["gather"] += t2 - t1
                tr["build"] += t3 - t2
                tr["n"] += 1
            return lst.total

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler is already running so that it is warm too)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        local_total = step()
    sync_all()

    # ---- timed region: device events on the launching stream, barrier+sync both sides
    # (per-phase times: the library records CUDA events around every phase of every timed build
    # and averages them; they are read ONCE after the region -- a read per step would make the host
    # wait for each build to finish before it may launch the next one)
    capi.check(L.cb_verlet_set_profiling(lst._h, 1))
    launches0 = L.cb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        local_total = step()
    ev1.record()
    sync_all()
    sampler.mark_end()
    ph = (C.c_double * 6)()
    capi.check(L.cb_verlet_get_phase_times(lst._h, ph))
    phase_sum = np.array(list(ph)) * args.steps
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = L.cb_kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([elapsed_ms, float(local_total), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms = float(tmax[0])
        global_total = float(tsum[1])
        launches = int(tsum[2])
    else:
        global_total = float(local_total)
    ms_per_step = elapsed_ms / args.steps
    value = global_total / (ms_per_step * 1e-3)

    # ---- per-kernel roofline, rank 0's numbers (CUDA events recorded inside the library on the
    # launching stream around each phase of every timed build).  v2 kernels: fused binning on the
    # internal pencil grid -> tile plan -> count pass (k_tile_count: tensor-core distance tiles,
    # hit masks, counts) -> offsets scan -> fill pass (k_tile_fill: rows written once at
    # offsets[i]).  Algorithmic bytes per particle follow SURVEY.md 8d: bin 32, build 36 + 4 K_s
    # (count pass: positions 24 + permutation 4 + counts 4; fill pass: offsets 4 + ids 4 K_s).
    peak, peak_kind = _peaks()
    phases = phase_sum / args.steps  # ms: bin, plan, count pass, scan, fill pass, total
    n_rows = num_local
    k_s = local_total / max(n_rows, 1)
    alg = {"binning": 32.0 * n_rows, "tile_plan": 0.0, "count_pass": 32.0 * n_rows,
           "offset_scan": 0.0, "fill_pass": n_rows * (4.0 + 4.0 * k_s)}
    names = ["binning", "tile_plan", "count_pass", "offset_scan", "fill_pass"]
    kernel_of = {"binning": "k_tbin_count + k_tbin_scatter", "tile_plan": "k_plan_blocks + k_plan_tiles + scans",
                 "count_pass": "k_tile_count", "offset_scan": "k_exclusive_scan + k_max_and_sum + k_sorted_dst",
                 "fill_pass": "k_tile_fill"}
    tj = {}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
        except Exception:
            tj = {}
    if world != 1 or args.workload != "cfg3":
        # the capture is of the single-GPU cfg3 build: it says nothing per launch about a slab or
        # another workload -- report null rather than a number that does not belong to this run
        tj = {"source": "none for this configuration (the committed ncu capture is the 1-GPU cfg3 build: "
                        "profiles/r02_traffic.json)"}
    per_kernel = {}
    for i, nm in enumerate(names):
        ms_k = float(phases[i])
        gbs = alg[nm] / (ms_k * 1e-3) / 1e9 if ms_k > 0 else 0.0
        per_kernel[nm] = {"kernels": kernel_of[nm], "ms": ms_k, "algorithmic_bytes": alg[nm],
                          "achieved_gbs": gbs, "frac": gbs / peak,
                          "dram_traffic_bytes": (tj.get("dram_bytes_per_launch") or {}).get(nm)}
    dom = "count_pass" if phases[2] >= phases[4] else "fill_pass"
    step_bytes = n_rows * (68.0 + 4.0 * k_s)
    step_gbs = step_bytes / (phases[5] * 1e-3) / 1e9 if phases[5] > 0 else 0.0
    roofline = {
        "bound": "hbm",
        "kernel": f"{kernel_of[dom]} ({dom}: the longest kernel of the step)",
        "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": per_kernel[dom]["frac"],
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
        "traffic": per_kernel[dom]["dram_traffic_bytes"],
        "traffic_source": tj.get("source"),
        "algorithmic_bytes_per_launch": alg[dom],
        "kernel_ms": per_kernel[dom]["ms"],
        "other_units": tj.get("other_units"),
        "step": {"achieved": step_gbs, "frac": step_gbs / peak, "algorithmic_bytes": step_bytes,
                 "ms": float(phases[5]),
                 "dram_traffic_bytes": (tj.get("dram_bytes_per_launch") or {}).get("step"),
                 "note": "whole build (bin + plan + count + scan + fill), N(68 + 4 K_s) bytes, against the HBM roof"},
        "phases": per_kernel,
    }

    # ---- end-to-end: host buffers in, list back to host, copies inside the timed region
    e2e = None
    if world == 1:
        host_x = torch.from_numpy(np.ascontiguousarray(xyz)).pin_memory()
        counts_h = torch.empty(num_local, dtype=torch.int32).pin_memory()
        offsets_h = torch.empty(num_local, dtype=torch.int32).pin_memory()
        nb_h = torch.empty(int(lst.total) if layout == cb.CSR else num_local * int(lst.width),
                           dtype=torch.int32).pin_memory()
        e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            lst.build_host(host_x, 0, num_local, radius, CELL_RATIO, gmin, gmax)
            lst.copy_to_host(counts_h, offsets_h if layout == cb.CSR else None, nb_h)

        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e_steps
        assert int(counts_h.to(torch.int64).sum()) == lst.total
        e2e = {"value": lst.total / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": e_steps,
               "h2d_bytes_per_step": int(host_x.numel() * 8),
               "d2h_bytes_per_step": int(4 * (2 * num_local + lst.total)),
               "api": "cb_verlet_build_host + cb_verlet_copy_to_host (pinned host buffers)"}

    if world > 1:
        # end to end at N GPUs: every rank uploads its owned positions from pinned host memory,
        # exchanges the halo, builds, and copies its part of the list back to pinned host memory
        own_elems = ((num_local + 31) // 32) * x_all.outer_stride      # whole SoAs of the owned part
        host_own = x_all.data[:own_elems].cpu().pin_memory()
        local_total = step()
        counts_h = torch.empty(cap, dtype=torch.int32).pin_memory()
        offsets_h = torch.empty(cap, dtype=torch.int32).pin_memory()
        nb_h = torch.empty(int(lst.total * 1.02) + 1024, dtype=torch.int32).pin_memory()
        e_steps = max(1, min(args.steps, 5))

        def e2e_step_n():
            x_all.data[:own_elems].copy_(host_own, non_blocking=True)
            step()
            lst.copy_to_host(counts_h, offsets_h, nb_h)

        e2e_step_n()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step_n()
        sync_all()
        dt = (time.perf_counter() - t0) / e_steps
        tt = torch.tensor([dt, float(lst.total), float(host_own.numel() * 8),
                           float(4 * (2 * lst._view.n + lst.total))], dtype=torch.float64, device="cuda")
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e = {"value": float(tsum[1]) / float(tmax[0]), "unit": UNIT, "ms_per_step": float(tmax[0]) * 1e3,
               "steps": e_steps, "h2d_bytes_per_step": int(tsum[2]), "d2h_bytes_per_step": int(tsum[3]),
               "api": "per rank: pinned host positions -> device, halo exchange, cb_verlet_build, "
                      "cb_verlet_copy_to_host (pinned); wall clock, max over ranks"}

    # ---- informational: LJ neighbor_parallel_for over the list just built (SURVEY.md 8d
    # "t_traverse"), thread-per-particle and warp-per-particle
    traverse = None
    if world == 1:
        traverse = {}
        f = cb.view_from_array(np.zeros((num_local, 3)))
        for name, op in (("serial", cb.OP_SERIAL), ("team", cb.OP_TEAM)):
            for _ in range(2):
                cb.neighbor_parallel_for_lj(0, num_local, lst, x, f, 1.0, 1.0, 2.5, op)
            torch.cuda.synchronize()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_steps = 5
            t0e.record()
            for _ in range(t_steps):
                cb.neighbor_parallel_for_lj(0, num_local, lst, x, f, 1.0, 1.0, 2.5, op)
            t1e.record()
            torch.cuda.synchronize()
            t_ms = t0e.elapsed_time(t1e) / t_steps
            # compulsory bytes: ids 4 K_s + every position once as x_i (24) and once as somebody's
            # x_j (24) + f_i 24 = N (72 + 4 K_s); worst case (no cache hit on x_j): N (48 + 28 K_s).
            # The roofline fraction is quoted on the COMPULSORY figure (the traversal is bound by the
            # x_j gathers in L1/L2, not by HBM: profiles/r02_final_lj_ncu_summary.txt).
            k_s_t = lst.total / max(num_local, 1)
            c_bytes = num_local * (72.0 + 4.0 * k_s_t)
            w_bytes = num_local * (48.0 + 28.0 * k_s_t)
            traverse[name] = {"ms": t_ms, "neighbors_per_s": lst.total / (t_ms * 1e-3),
                              "compulsory_gbs": c_bytes / (t_ms * 1e-3) / 1e9,
                              "frac_of_hbm_peak": c_bytes / (t_ms * 1e-3) / 1e9 / peak,
                              "worst_case_gbs": w_bytes / (t_ms * 1e-3) / 1e9}
        del f

    # ---- parity, outside the timed region.  N > 1: every rank rebuilds the UNDECOMPOSED box on
    # its own GPU and compares counts and a hash of every owned row (local ids mapped to global
    # ids, ghosts through a gid field shipped by the NCCL halo) with its sharded list.
    parity = None
    if world > 1 and args.workload == "cfg3" and not args.no_parity_check:
        local_total = step()            # the list under test: the timed path, once more
        x_own = cb.Slice(x_all.data, num_local, x_all.outer_stride, x_all.vlen, x_all.comp_stride, 3)
        halo = slab.create_halo(x_own, num_local)
        n_tot = halo.numLocal() + halo.numGhost()
        g0 = _fcc_first_id(rank, world)
        gid_np = np.full((cap, 1), -1, dtype=np.int32)
        gid_np[:num_local, 0] = g0 + np.arange(num_local)
        gid = cb.view_from_array(gid_np)
        x_chk = cb.slice_from_array(store, vlen=32)
        comm.gather(halo, cb.Slice(x_chk.data, n_tot, x_chk.outer_stride, x_chk.vlen, x_chk.comp_stride, 3),
                    cb.Slice(gid.data, n_tot, 1, 1, 1, 1))
        x_cmp = cb.Slice(x_all.data, n_tot, x_all.outer_stride, x_all.vlen, x_all.comp_stride, 3)
        x_ref = cb.Slice(x_chk.data, n_tot, x_chk.outer_stride, x_chk.vlen, x_chk.comp_stride, 3)
        same_ghosts = bool(torch.equal(x_cmp.to_array(), x_ref.to_array()))
        idmap = gid.data[:n_tot].to(torch.int64)
        h_loc = _row_hashes_gpu(torch, lst._data.counts, lst._data.offsets, lst._data.neighbors,
                                0, num_local, idmap)
        c_loc = lst._data.counts[:num_local].clone()
        full_xyz = np.concatenate([_fcc_slab(r, world)[0] for r in range(world)])
        x_full = cb.slice_from_array(full_xyz, vlen=32)
        lst_full = cb.VerletList(algorithm=algo, layout=layout)
        lst_full.build(x_full, 0, full_xyz.shape[0], RADIUS, CELL_RATIO, gmin, gmax)
        h_ref = _row_hashes_gpu(torch, lst_full._data.counts, lst_full._data.offsets,
                                lst_full._data.neighbors, g0, g0 + num_local)
        ok = (same_ghosts and bool(torch.equal(c_loc, lst_full._data.counts[g0:g0 + num_local]))
              and bool(torch.equal(h_loc, h_ref)))
        flag = torch.tensor([1.0 if ok else 0.0, float(local_total), 0.0], dtype=torch.float64, device="cuda")
        flag[2] = float(lst_full.total) if rank == 0 else 0.0
        fmin = flag.clone()
        dist.all_reduce(fmin, op=dist.ReduceOp.MIN)
        fsum = flag.clone()
        dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
        ok_all = bool(fmin[0] == 1.0) and float(fsum[1]) == float(fsum[2])
        parity = {"parity_check": "ok" if ok_all else "FAILED",
                  "against": "the undecomposed 16 078 716-atom build on one GPU (every rank): counts and a "
                             "64-bit hash of every owned row in global ids; ghost positions of the peer halo "
                             "bit-equal to the NCCL halo's",
                  "global_total": float(fsum[1]), "single_gpu_total": float(fsum[2])}
        del lst_full, x_full, h_ref, h_loc

    if world > 1 and os.environ.get("CB_BENCH_TRACE") == "1":
        sys.stderr.write("rank %d trace (ms/step): plan %.3f gather %.3f build %.3f\n" % (
            rank, 1e3 * tr["plan"] / max(tr["n"], 1), 1e3 * tr["gather"] / max(tr["n"], 1),
            1e3 * tr["build"] / max(tr["n"], 1)))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _, res, ps_s = cpu_reference_run(steps=3, warmup=1)
        if args.workload == "cfg3" and not args.no_parity_check:
            # N = 1: the oracle's list of the cpu_baseline sample against the CUDA path's list of the
            # same sample (counts, offsets, a hash of every row); the full-size comparison is
            # tests/test_gpu_parity.py::test_fcc_16m_matches_oracle
            import oracle
            xs_ = cb.slice_from_array(ps_s.xyz, vlen=32)
            ls_ = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
            ls_.build(xs_, 0, ps_s.n, ps_s.radius, CELL_RATIO, ps_s.grid_min, ps_s.grid_max)
            c_ = ls_._data.counts.cpu().numpy()
            o_ = ls_._data.offsets.cpu().numpy()
            n_ = ls_._data.neighbors[: ls_.total].cpu().numpy()
            ok = (ls_.total == res.total and np.array_equal(c_, res.counts) and np.array_equal(o_, res.offsets)
                  and np.array_equal(oracle.row_hashes(oracle.CSR, c_, o_, n_),
                                     oracle.row_hashes(oracle.CSR, res.counts, res.offsets, res.neighbors)))
            parity = {"parity_check": "ok" if ok else "FAILED",
                      "against": f"the CPU oracle on the cpu_baseline sample ({ps_s.n} atoms): counts, offsets "
                                 "and a 64-bit hash of every row"}
            del ls_, xs_, n_
        del res

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _workload_config(world) if args.workload == "cfg3" else
        {"workload": args.workload + " (BASELINE.json parity configuration, informational)",
         "particles": int(num_local),
         # per-row work spread the scheduler has to absorb (BASELINE.md section 4 asks for it on the
         # clustered set): rows are handed out as tiles of 16 by a global ticket, so the spread
         # shows up as tail time, not as idle SMs
         "load_imbalance": {"max_neighbors_per_row": int(lst._data.max_n),
                            "mean_neighbors_per_row": float(global_total) / max(int(num_local), 1),
                            "factor_max_over_mean": float(lst._data.max_n) * int(num_local) / max(float(global_total), 1.0)}},
        "neighbors_per_step": global_total,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "lj_traverse": traverse,
        "parity_check": parity["parity_check"] if parity else None,
        "parity": parity,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _emit(line: dict) -> None:
    """Print THE one JSON line on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    # Libraries (NCCL prints its version banner) must not pollute stdout: everything but the
    # final JSON line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="cfg3 (default) is the headline line; the others are the BASELINE.json "
                         "parity configurations, timed for DESIGN.md only (1 GPU)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "cfg5":
        return run_cfg5(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
