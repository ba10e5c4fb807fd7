"""Sharded parity on ONE GPU: P "virtual" slabs (P = 2, 4, 8) run the same C entries the
multi-rank step runs -- fused ghost selection/plan (cb_slab_halo_plan), pack (cb_comm_pack),
unpack (cb_comm_unpack), owner-local build (cb_verlet_build with begin=0, end=num_local) --
with a device-to-device copy standing in for the NCCL send/recv or the peer-memory push.
The union of the owner-local lists, mapped to global ids, must equal the oracle's list of
the whole box exactly (SURVEY.md section 8e "parity under sharding"); forces from the sharded
half list + ghost scatter must match the oracle within 1e-12.

The real transports are covered by tests/test_gpu_comm.py on >= 2 GPUs and by
tests/test_comm_gloo.py (gloo, world 2 and 3) on CPU.
"""
import numpy as np
import pytest
import torch

from cabana_b200 import datasets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    assert torch.cuda.is_available()
    from cabana_b200 import core

    return core


def _virtual_slab_lists(cb, comm, ps, P, algo, with_forces=False):
    """Returns {global id: sorted global neighbour ids} over all owners (and LJ forces)."""
    k = comm.CudaCommKernels()
    L = ps.grid_max[0]
    bounds = [L * g / P for g in range(P + 1)]
    hw = ps.radius * (1.0 + 2.0**-40)
    owner = np.minimum(np.searchsorted(np.asarray(bounds[1:-1]), ps.xyz[:, 0], side="right"), P - 1)
    mine = [np.nonzero(owner == g)[0] for g in range(P)]
    cap = [len(m) + 2 * int(len(m) * 1.2 * hw / (bounds[1] - bounds[0])) + 4096 for m in mine]
    xs, gids, nloc = [], [], []
    for g in range(P):
        buf = np.zeros((cap[g], 3))
        buf[: len(mine[g])] = ps.xyz[mine[g]]
        xs.append(cb.slice_from_array(buf, vlen=32))
        gi = np.full((cap[g], 1), -1, dtype=np.int32)
        gi[: len(mine[g]), 0] = mine[g]
        gids.append(cb.view_from_array(gi))
        nloc.append(len(mine[g]))

    def sub(sl, n, nc):
        return cb.Slice(sl.data, n, sl.outer_stride, sl.vlen, sl.comp_stride, nc)

    # ---- plan + pack on every "rank", then deliver: lower neighbour's ghosts first
    packed = {}
    steers = {}
    for g in range(P):
        lo_rank = g - 1 if g > 0 else -1
        hi_rank = g + 1 if g < P - 1 else -1
        x_own = sub(xs[g], nloc[g], 3)
        steer, n_lo, n_hi = k.slab_halo_plan(x_own, nloc[g], bounds[g] + hw, bounds[g + 1] - hw,
                                             lo_rank, hi_rank)
        steers[g] = (steer, n_lo, n_hi)
        fields = [sub(xs[g], nloc[g], 3), sub(gids[g], nloc[g], 1)]
        tb = k.tuple_bytes(fields)
        for dst, cnt, st in ((lo_rank, n_lo, steer[:max(n_lo, 1)]),
                             (hi_rank, n_hi, steer[max(nloc[g], 1):])):
            if dst >= 0 and cnt > 0:
                buf = torch.empty(cnt * tb, dtype=torch.uint8, device="cuda")
                k.pack(fields, st.contiguous(), cnt, buf)
                packed[(g, dst)] = (buf.clone(), cnt)    # the "wire": a device-to-device copy
    nghost = []
    for g in range(P):
        at = nloc[g]
        for src in (g - 1, g + 1):
            if (src, g) in packed:
                buf, cnt = packed[(src, g)]
                assert at + cnt <= cap[g]
                k.unpack([sub(xs[g], at + cnt, 3), sub(gids[g], at + cnt, 1)], at, cnt, buf)
                at += cnt
        nghost.append(at - nloc[g])

    rows = {}
    forces = np.zeros((ps.n, 3)) if with_forces else None
    ghost_forces = {}
    for g in range(P):
        ntot = nloc[g] + nghost[g]
        gmin_x = bounds[g] - hw if g > 0 else bounds[g]
        gmax_x = bounds[g + 1] + hw if g < P - 1 else bounds[g + 1]
        lmin = (gmin_x, ps.grid_min[1], ps.grid_min[2])
        lmax = (gmax_x, ps.grid_max[1], ps.grid_max[2])
        x_tot = sub(xs[g], ntot, 3)
        lst = cb.VerletList(x_tot, 0, nloc[g], ps.radius, 1.0, lmin, lmax, algorithm=algo,
                            layout=cb.CSR)
        counts = lst._data.counts.cpu().numpy()
        offsets = lst._data.offsets.cpu().numpy()
        nb = lst._data.neighbors.cpu().numpy()
        gid = gids[g].to_array().cpu().numpy()[:ntot, 0]
        assert counts[nloc[g]:].sum() == 0, "ghost rows must stay empty (begin=0,end=num_local)"
        for i in range(nloc[g]):
            rows[int(gid[i])] = sorted(int(gid[j]) for j in nb[offsets[i]: offsets[i] + counts[i]])
        if with_forces:
            f = cb.view_from_array(np.zeros((ntot, 3)))
            cb.neighbor_parallel_for_lj(0, nloc[g], lst, x_tot, f, 1.0, 1.0, 2.5, cb.OP_TEAM)
            fa = f.to_array().cpu().numpy()
            forces[gid[:nloc[g]]] += fa[:nloc[g]]
            ghost_forces[g] = (gid[nloc[g]:ntot], fa[nloc[g]:ntot])
    if with_forces:
        # half lists: forces accumulated on ghosts go home (Cabana::scatter semantics)
        for g, (ids, fa) in ghost_forces.items():
            np.add.at(forces, ids, fa)
    return rows, forces


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("algo_name", ["full", "half"])
def test_virtual_slabs_union_equals_oracle(orc, cb, P, algo_name):
    from cabana_b200 import comm

    ps = datasets.fcc_lattice(28, jitter=0.03)   # 87 808 atoms, box 47 wide: 8 slabs of 2.1 r
    algo = cb.FULL if algo_name == "full" else cb.HALF
    rows, _ = _virtual_slab_lists(cb, comm, ps, P, algo)
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, 1.0, ps.grid_min,
                           ps.grid_max, algo=orc.FULL if algo_name == "full" else orc.HALF)
    assert len(rows) == ps.n
    for i in range(ps.n):
        assert rows[i] == sorted(int(v) for v in ref.row(i)), (P, algo_name, i)


@pytest.mark.parametrize("P", [2, 8])
def test_virtual_slabs_uniform_and_forces(orc, cb, P):
    from cabana_b200 import comm

    ps = datasets.uniform_box(60_000, 20240105, radius=3.0)
    ox = orc.view_from_xyz(ps.xyz)
    full = orc.verlet_build(ox, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algo=orc.FULL)
    f_ref, fabs = orc.lj_forces(ox, orc.CSR, full.counts, full.offsets, full.neighbors, 0, 0, ps.n,
                                1.0, 1.0, 2.5)
    for algo, oalgo in ((cb.FULL, orc.FULL), (cb.HALF, orc.HALF)):
        rows, forces = _virtual_slab_lists(cb, comm, ps, P, algo, with_forces=True)
        ref = full if oalgo == orc.FULL else orc.verlet_build(
            ox, 0, ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max, algo=oalgo)
        for i in range(ps.n):
            assert rows[i] == sorted(int(v) for v in ref.row(i)), (P, algo, i)
        assert np.all(np.abs(forces - f_ref) <= 1e-12 * np.maximum(fabs, 1e-300))


def _single_rank_step(rank, port, ret):
    import os

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        import oracle
        from cabana_b200 import comm
        from cabana_b200 import core as cb

        ps = datasets.fcc_lattice(12, jitter=0.03)
        slab = comm.SlabDecomposition([0.0, ps.grid_max[0]], ps.radius)
        cap = ps.n + 1000
        store = np.zeros((cap, 3))
        store[: ps.n] = ps.xyz
        x_all = cb.slice_from_array(store, vlen=32)
        peer = slab.create_peer_halo([x_all], 1000)
        lst = cb.VerletList(algorithm=cb.HALF, layout=cb.CSR)
        ok = True
        for _ in range(3):       # the entry is re-entrant: sequence numbers, buffers re-used
            n_lo, n_hi = peer.step(lst, x_all, [x_all], ps.n, ps.radius, 1.0, ps.grid_min, ps.grid_max)
            ok &= (n_lo, n_hi) == (0, 0)
        ref = oracle.verlet_build(oracle.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, 1.0, ps.grid_min,
                                  ps.grid_max, algo=oracle.HALF)
        counts = lst._data.counts.cpu().numpy()[: ps.n]
        got, _ = oracle.sorted_rows_flat(oracle.CSR, counts, lst._data.offsets.cpu().numpy()[: ps.n],
                                         lst._data.neighbors.cpu().numpy(), 0)
        ok &= bool(np.array_equal(counts, ref.counts)) and bool(np.array_equal(got, ref.sorted_rows_flat()[0]))
        peer.close()
        ret[0] = bool(ok)
    except Exception:
        import traceback

        ret[0] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_slab_step_entry_single_rank(orc):
    """cb_slab_step (push -> wait -> unpack -> build from one C entry) on a one-rank "slab
    decomposition": no neighbours, so the step must equal a plain build; the two-rank case is
    tests/test_gpu_comm.py::test_two_gpu_slab_build_equals_single_gpu."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_single_rank_step, args=(port, ret), nprocs=1, join=True)
    assert ret[0] is True, ret[0]


def _two_slab_step_lists(cb, ps, algo, reps):
    """Two virtual ranks in ONE process drive cb_slab_step with its device-side ghost counts: each
    rank's windows live on the same GPU and the neighbour pushes through their local base address.
    The pushes of both ranks are issued first (cb_slab_halo_push), so the step's own wait finds them
    -- its second push of the same sequence rewrites identical bytes."""
    import ctypes as C

    from cabana_b200 import capi

    L = capi.lib()
    P = 2
    Lx = ps.grid_max[0]
    bounds = [Lx * g / P for g in range(P + 1)]
    hw = ps.radius * (1.0 + 2.0**-40)
    owner = np.minimum(np.searchsorted(np.asarray(bounds[1:-1]), ps.xyz[:, 0], side="right"), P - 1)
    mine = [np.nonzero(owner == g)[0] for g in range(P)]
    capw = max(int(len(m) * 1.5 * hw / (bounds[1] - bounds[0])) + 1024 for m in mine)
    xs, gids, nloc = [], [], []
    for g in range(P):
        cap = len(mine[g]) + capw
        buf = np.zeros((cap, 3))
        buf[: len(mine[g])] = ps.xyz[mine[g]]
        xs.append(cb.slice_from_array(buf, vlen=32))
        gi = np.full((cap, 1), -1, dtype=np.int32)
        gi[: len(mine[g]), 0] = mine[g]
        gids.append(cb.view_from_array(gi))
        nloc.append(len(mine[g]))
    fields = [[xs[g], gids[g]] for g in range(P)]
    farr = [(capi.Field * 2)(*[f.field_desc() for f in fields[g]]) for g in range(P)]
    tb = int(L.cb_comm_tuple_bytes(farr[0], 2))
    # rank 0 receives from its upper neighbour, rank 1 from its lower one
    win = [C.c_void_p(), C.c_void_p()]
    base = [C.c_void_p(), C.c_void_p()]
    for g in range(P):
        capi.check(L.cb_p2p_window_create(C.byref(win[g]), C.c_int64(capw), C.c_int64(tb)))
        capi.check(L.cb_p2p_window_local_base(win[g], C.byref(base[g])))
    steer = [torch.empty(2 * nloc[g], dtype=torch.int32, device="cuda") for g in range(P)]
    lsts = [cb.VerletList(algorithm=algo, layout=cb.CSR) for _ in range(P)]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def args(g):
        peer_lo = base[0] if g == 1 else C.c_void_p()
        peer_hi = base[1] if g == 0 else C.c_void_p()
        from_lo = win[1] if g == 1 else None
        from_hi = win[0] if g == 0 else None
        return peer_lo, peer_hi, from_lo, from_hi

    out = []
    try:
        for seq in range(1, reps + 1):
            for g in range(P):
                peer_lo, peer_hi, _, _ = args(g)
                d = xs[g].positions_desc()
                d.n = nloc[g]
                capi.check(L.cb_slab_halo_push(
                    C.byref(d), farr[g], 2, C.c_int64(nloc[g]), C.c_double(bounds[g] + hw),
                    C.c_double(bounds[g + 1] - hw), peer_lo, peer_hi, C.c_int64(capw),
                    C.c_uint64(seq), C.c_void_p(steer[g].data_ptr()), st))
            rows = {}
            for g in range(P):
                peer_lo, peer_hi, from_lo, from_hi = args(g)
                lmin = (bounds[g] - hw if g > 0 else bounds[g], ps.grid_min[1], ps.grid_min[2])
                lmax = (bounds[g + 1] + hw if g < P - 1 else bounds[g + 1], ps.grid_max[1], ps.grid_max[2])
                d = xs[g].positions_desc()
                counts = (C.c_int64 * 2)()
                capi.check(L.cb_slab_step(
                    lsts[g]._h, C.byref(d), farr[g], 2, C.c_int64(nloc[g]),
                    C.c_double(bounds[g] + hw), C.c_double(bounds[g + 1] - hw), peer_lo, peer_hi,
                    from_lo, from_hi, C.c_int64(capw), C.c_uint64(seq),
                    C.c_void_p(steer[g].data_ptr()), C.c_double(ps.radius), C.c_double(1.0),
                    capi.d3(lmin), capi.d3(lmax), C.c_int64(0), C.c_int(algo), C.c_int(cb.CSR),
                    C.c_int(lsts[g].build_tag), counts, st))
                lsts[g]._refresh()
                nghost = int(counts[0]) + int(counts[1])
                assert nghost > 0
                ntot = nloc[g] + nghost
                assert lsts[g]._data.counts.numel() == ntot, "list size = owned + ghosts"
                cnt = lsts[g]._data.counts.cpu().numpy()
                off = lsts[g]._data.offsets.cpu().numpy()
                nb = lsts[g]._data.neighbors.cpu().numpy()
                gid = gids[g].to_array().cpu().numpy()[:ntot, 0]
                assert (gid >= 0).all()
                assert cnt[nloc[g]:].sum() == 0
                for i in range(nloc[g]):
                    rows[int(gid[i])] = sorted(int(gid[j]) for j in nb[off[i]: off[i] + cnt[i]])
            out.append(rows)
    finally:
        torch.cuda.synchronize()
        for g in range(P):
            L.cb_p2p_window_destroy(win[g])
    return out


@pytest.mark.parametrize("algo_name", ["full", "half"])
def test_slab_step_device_side_ghost_counts(orc, cb, algo_name, monkeypatch):
    """The sync-free step (k_halo_wait_unpack + a build that reads the particle count on the device
    + the speculative fill of the rebuild) on one GPU: union of the two owner-local lists equals the
    oracle's list of the whole box, on the first build and on two rebuilds; the older path
    (CB_SLAB_SYNC=1) gives the same lists."""
    ps = datasets.uniform_box(30_000, 20240177, radius=3.0)
    algo = cb.FULL if algo_name == "full" else cb.HALF
    ref = orc.verlet_build(orc.view_from_xyz(ps.xyz), 0, ps.n, ps.radius, 1.0, ps.grid_min,
                           ps.grid_max, algo=orc.FULL if algo_name == "full" else orc.HALF)
    want = [sorted(int(v) for v in ref.row(i)) for i in range(ps.n)]
    for rows in _two_slab_step_lists(cb, ps, algo, reps=3):
        assert len(rows) == ps.n
        for i in range(ps.n):
            assert rows[i] == want[i], f"particle {i}"
    monkeypatch.setenv("CB_SLAB_SYNC", "1")
    rows = _two_slab_step_lists(cb, ps, algo, reps=1)[0]
    for i in range(ps.n):
        assert rows[i] == want[i], f"particle {i} (sync path)"
