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
CPU_SAMPLE_CELLS = int(os.environ.get("CB_BENCH_CPU_CELLS", "64"))  # FCC cells per side of the CPU sample


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
def cpu_reference_run(steps: int, warmup: int, cells: int = CPU_SAMPLE_CELLS):
    """Time the CPU port of the reference path (oracle/) on a bounded FCC sample."""
    import oracle
    from cabana_b200 import datasets

    ps = datasets.fcc_lattice(cells)
    x = oracle.slice_from_xyz(ps.xyz, vlen=16)  # host AoSoA vector length (PerformanceTraits)
    threads = oracle.num_threads()
    times, total = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        res = oracle.verlet_build(x, 0, ps.n, ps.radius, CELL_RATIO, ps.grid_min, ps.grid_max,
                                  algo=oracle.FULL, layout=oracle.CSR)
        dt = time.perf_counter() - t0
        total = res.total
        if it >= warmup:
            times.append(dt)
    avg = float(np.mean(times))
    return {
        "value": total / avg,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"{ps.n}-atom FCC sub-lattice ({cells}^3 cells) of the same workload, "
                  f"{len(times)} timed builds, avg {avg*1e3:.1f} ms (min {min(times)*1e3:.1f}, "
                  f"max {max(times)*1e3:.1f}); C++/OpenMP restatement of the reference "
                  "(Kokkos is not installable here), -O3 -ffp-contract=off",
    }, avg, total, ps.n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 2))
    cpu, avg, total, n = cpu_reference_run(steps, warmup)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": avg * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _workload_config(args.gpus),
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
    phase_sum = np.zeros(6)
    launches0 = L.cb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        local_total = step()
        ph = (C.c_double * 6)()
        capi.check(L.cb_verlet_get_phase_times(lst._h, ph))
        phase_sum += np.array(list(ph))
    ev1.record()
    sync_all()
    sampler.mark_end()
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

    # ---- per-kernel roofline for the dominant kernel, rank 0's numbers.  The dominant
    # kernel is the single test pass (k_verlet_column: every pair tested once, rows written
    # to the binned temporary, counts produced); the reorder kernel is timed separately.
    peak, peak_kind = _peaks()
    phases = phase_sum / args.steps  # ms: bin, gather, test pass, scan, reorder, total
    n_rows = num_local
    k_s = local_total / max(n_rows, 1)
    # test-pass algorithmic bytes per launch: read positions 24 + ids 4 + cell offsets ~4,
    # write counts 4 + 4*K_s neighbour ids   (36 + 4 K_s per particle)
    fill_bytes = n_rows * (36.0 + 4.0 * k_s)
    fill_gbs = fill_bytes / (phases[2] * 1e-3) / 1e9 if phases[2] > 0 else 0.0
    # whole step: bin 32 + build 36 + 4 K_s
    step_bytes = n_rows * (68.0 + 4.0 * k_s)
    step_gbs = step_bytes / (phases[5] * 1e-3) / 1e9 if phases[5] > 0 else 0.0
    traffic, other_roofs = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get("k_verlet_column_dram_bytes_per_launch")
            # SURVEY.md 8d asks for the FP64 roof beside the HBM one: the pair tests run in FP32
            # (tier 1), only the ambiguous band reaches the exact FP64 tier, so the FP64 pipe is
            # idle; the unit that actually limits the kernel is L1TEX (ncu, profiles/)
            other_roofs = {"fp64_pipe_active_pct": tj.get("k_verlet_column_fp64_pipe_active_pct"),
                           "l1tex_throughput_pct": tj.get("k_verlet_column_l1tex_throughput_pct"),
                           "issue_active_pct": tj.get("k_verlet_column_issue_active_pct"),
                           "source": "ncu --set full, profiles/r01_final_ncu_summary.txt"}
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": "k_verlet_column (single test pass; dominant kernel of the step)",
        "achieved": fill_gbs, "peak": peak, "unit": "GB/s", "frac": fill_gbs / peak,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "traffic": traffic,
        "algorithmic_bytes_per_launch": fill_bytes,
        "kernel_ms": float(phases[2]),
        "other_units": other_roofs,
        "step": {"achieved": step_gbs, "frac": step_gbs / peak, "algorithmic_bytes": step_bytes,
                 "note": "whole build (bin+gather+count+scan+fill) against the HBM roof"},
        "phase_ms": {"binning": float(phases[0]), "gather_permute": float(phases[1]),
                     "test_pass": float(phases[2]), "offset_scan": float(phases[3]),
                     "row_reorder": float(phases[4]), "build_total": float(phases[5])},
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

    # ---- informational: the same build with CB_ROWS_BINNED (rows stay in cell order; no
    # offsets scan / reorder pass).  NOT the headline: `value` above is the reference layout.
    binned = None
    if world == 1 and layout == cb.CSR:
        lst_b = cb.VerletList(algorithm=algo, layout=layout, row_placement=cb.ROWS_BINNED)
        for _ in range(max(args.warmup, 2)):
            lst_b.build(x, 0, num_local, radius, CELL_RATIO, gmin, gmax)
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b_steps = max(1, min(args.steps, 10))
        b0.record()
        for _ in range(b_steps):
            lst_b.build(x, 0, num_local, radius, CELL_RATIO, gmin, gmax)
        b1.record()
        torch.cuda.synchronize()
        b_ms = b0.elapsed_time(b1) / b_steps
        assert lst_b.total == lst.total
        binned = {"ms_per_step": b_ms, "value": lst_b.total / (b_ms * 1e-3), "unit": UNIT,
                  "note": "opt-in cb_verlet_set_row_placement(CB_ROWS_BINNED): same neighbour sets, "
                          "rows left in cell order inside `neighbors` (offsets not monotone)"}
        del lst_b

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
            # worst-case algorithmic bytes: ids 4 K_s + x_i 24 + gathered x_j 24 K_s + f_i 24
            t_bytes = num_local * (48.0 + 28.0 * (lst.total / max(num_local, 1)))
            traverse[name] = {"ms": t_ms, "neighbors_per_s": lst.total / (t_ms * 1e-3),
                              "worst_case_gbs": t_bytes / (t_ms * 1e-3) / 1e9,
                              "frac_of_hbm_peak": t_bytes / (t_ms * 1e-3) / 1e9 / peak}
        del f

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
        cpu, _, _, _ = cpu_reference_run(steps=3, warmup=1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _workload_config(world) if args.workload == "cfg3" else
        {"workload": args.workload + " (BASELINE.json parity configuration, informational)",
         "particles": int(num_local)},
        "neighbors_per_step": global_total,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "binned_rows": binned,
        "lj_traverse": traverse,
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
