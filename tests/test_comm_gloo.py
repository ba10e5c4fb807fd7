"""world_size-2 (and 3) gloo tests of the Halo / Distributor host logic on CPU.

The device kernels are replaced by tests/_comm_double.py (a test double, not a product
fallback); what is under test is cabana_b200.comm: the count exchange, the neighbour
ordering (self first, ascending), buffer layout, ghost placement, scatter and migrate
semantics, following core/unit_test/tstHalo.hpp and tstDistributor.hpp.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover - surfaced through ret
        import traceback

        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def _spawn(fn, world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


# ------------------------------------------------------------------------------------ cases
def _global_particles(n=4000, seed=9, L=40.0):
    rng = np.random.Generator(np.random.Philox(key=seed))
    xyz = rng.random((n, 3)) * np.array([L, 10.0, 10.0])
    return xyz, L


def _case_slab_halo(rank, world):
    from _comm_double import CpuCommKernels, CpuSlice
    from cabana_b200 import comm

    xyz, L = _global_particles()
    r = 1.5
    bounds = [L * g / world for g in range(world + 1)]
    owner = np.minimum((xyz[:, 0] / (L / world)).astype(int), world - 1)
    mine = np.where(owner == rank)[0]
    num_local = len(mine)
    slab = comm.SlabDecomposition(bounds, r, kernels=CpuCommKernels())
    store = torch.zeros((num_local + 2000, 3), dtype=torch.float64)
    store[:num_local] = torch.from_numpy(xyz[mine])
    gid = torch.full((num_local + 2000, 1), -1, dtype=torch.int64)
    gid[:num_local, 0] = torch.from_numpy(mine)
    x = CpuSlice(store)
    halo = slab.create_halo(x, num_local)
    # neighbour order: self first, then ascending (Cabana_CommunicationPlanBase.hpp:374-394)
    assert halo.neighborRank(0) == rank
    assert halo.neighbors[1:] == sorted(halo.neighbors[1:])
    assert set(halo.neighbors[1:]) <= {rank - 1, rank + 1}
    assert halo.numLocal() == num_local
    comm.gather(halo, x, CpuSlice(gid))
    n_tot = halo.numLocal() + halo.numGhost()
    ghosts = gid[num_local:n_tot, 0].numpy()
    # expected ghosts: particles of the adjacent slabs within r of my faces
    lo, hi = bounds[rank], bounds[rank + 1]
    w = slab.halo_width
    exp = []
    if rank > 0:
        exp += list(np.where((owner == rank - 1) & (xyz[:, 0] >= lo - w))[0])
    if rank < world - 1:
        exp += list(np.where((owner == rank + 1) & (xyz[:, 0] < hi + w))[0])
    assert sorted(ghosts.tolist()) == sorted(exp)
    # ghosts carry bit-identical coordinates and are grouped by source rank, ascending
    assert np.array_equal(store[num_local:n_tot].numpy(), xyz[ghosts])
    src = owner[ghosts]
    assert np.all(np.diff(src) >= 0)
    # every cross-slab pair within r is now visible locally
    d2 = ((xyz[mine][:, None, :] - xyz[None, :, :]) ** 2).sum(-1)
    need = set(np.where((d2 <= r * r).any(axis=0) & (owner != rank))[0].tolist())
    assert need <= set(ghosts.tolist())

    # persistent Gather object (Cabana_Halo.hpp:392-640): same ghosts, buffers kept and re-used
    store2 = store.clone()
    store2[num_local:] = 0.0
    gid2 = gid.clone()
    gid2[num_local:] = -1
    g = comm.createGather(halo, CpuSlice(store2), CpuSlice(gid2), overallocation=1.5)
    assert g.sendSize() == halo.totalNumExport() and g.receiveSize() == halo.totalNumImport()
    assert g.sendCapacity() == int(halo.totalNumExport() * 1.5)
    g.apply()
    assert torch.equal(store2[:n_tot], store[:n_tot]) and torch.equal(gid2[:n_tot], gid[:n_tot])
    cap = g.sendCapacity()
    g.reserve(halo, CpuSlice(store2), CpuSlice(gid2))          # same sizes: no reallocation
    assert g.sendCapacity() == cap
    g.shrinkToFit()
    assert g.sendCapacity() == halo.totalNumExport() and g.receiveCapacity() == halo.totalNumImport()
    store2[num_local:] = 0.0
    g.apply()
    assert torch.equal(store2[:n_tot], store[:n_tot])
    with pytest.raises(RuntimeError):
        comm.createGather(halo, CpuSlice(store2), overallocation=0.5)

    # scatter: ghost contributions are summed into their owners (tstHalo.hpp:146-185)
    f = torch.zeros((num_local + 2000, 3), dtype=torch.float64)
    f[num_local:n_tot] = 1.0 + rank
    fs = CpuSlice(f)
    comm.scatter(halo, fs)
    sent_lo = (xyz[mine][:, 0] < lo + w) & (rank > 0)
    sent_hi = (xyz[mine][:, 0] >= hi - w) & (rank < world - 1)
    expect = np.zeros(num_local)
    expect += np.where(sent_lo, 1.0 + (rank - 1), 0.0)
    expect += np.where(sent_hi, 1.0 + (rank + 1), 0.0)
    assert np.array_equal(f[:num_local, 0].numpy(), expect)
    # persistent Scatter object: same sums
    f2 = torch.zeros((num_local + 2000, 3), dtype=torch.float64)
    f2[num_local:n_tot] = 1.0 + rank
    sc = comm.createScatter(halo, CpuSlice(f2))
    assert sc.sendSize() == halo.totalNumImport() and sc.receiveSize() == halo.totalNumExport()
    sc.apply()
    assert np.array_equal(f2[:num_local, 0].numpy(), expect)
    # every arithmetic value type scatters (buffers sized by the element size); an unsupported
    # one is rejected BEFORE any message is posted (mismatched send/recv sizes)
    for dt in (torch.float32, torch.int32, torch.int64):
        ft = torch.zeros((num_local + 2000, 2), dtype=dt)
        ft[num_local:n_tot] = 1 + rank
        comm.scatter(halo, CpuSlice(ft))
        assert np.array_equal(ft[:num_local, 1].numpy().astype(np.float64), expect)
        ft2 = torch.zeros((num_local + 2000, 2), dtype=dt)
        ft2[num_local:n_tot] = 1 + rank
        comm.createScatter(halo, CpuSlice(ft2)).apply()
        assert torch.equal(ft2, ft)
    with pytest.raises(TypeError):
        comm.scatter(halo, CpuSlice(torch.zeros((num_local + 2000, 1), dtype=torch.int16)))


def _case_migrate(rank, world):
    from _comm_double import CpuCommKernels, CpuSlice
    from cabana_b200 import comm

    xyz, L = _global_particles(seed=21)
    bounds = [L * g / world for g in range(world + 1)]
    owner = np.minimum((xyz[:, 0] / (L / world)).astype(int), world - 1)
    mine = np.where(owner == rank)[0]
    # move everyone by a drift so some cross faces and some leave the box (dropped, -1)
    moved = xyz.copy()
    moved[:, 0] += 3.0
    x = CpuSlice(torch.from_numpy(moved[mine].copy()))
    gid = CpuSlice(torch.from_numpy(mine.copy()).reshape(-1, 1))
    slab = comm.SlabDecomposition(bounds, 1.0, kernels=CpuCommKernels())
    distributor = slab.create_distributor(x, len(mine))
    new_owner = np.where(moved[:, 0] <= L, np.minimum((moved[:, 0] / (L / world)).astype(int), world - 1), -1)
    n_new = distributor.totalNumImport()
    assert n_new == int((new_owner == rank).sum())
    dst_x = CpuSlice(torch.zeros((n_new, 3), dtype=torch.float64))
    dst_id = CpuSlice(torch.zeros((n_new, 1), dtype=torch.int64))
    comm.migrate(distributor, [x, gid], [dst_x, dst_id])
    got = dst_id.t[:, 0].numpy()
    assert sorted(got.tolist()) == sorted(np.where(new_owner == rank)[0].tolist())
    assert np.array_equal(dst_x.t.numpy(), moved[got])
    # staying elements come first (self is neighbour 0; impl/Cabana_Migrate_Mpi.hpp:92-99)
    stay = distributor.numImport(0)
    assert np.all(owner[got[:stay]] == rank)
    assert np.all(owner[got[stay:]] != rank)


def _case_ring_distributor(rank, world):
    # tstDistributor.hpp test "ring": everything goes to the next rank
    from _comm_double import CpuCommKernels, CpuSlice
    from cabana_b200 import comm

    n = 100
    dest = torch.full((n,), (rank + 1) % world, dtype=torch.int32)
    d = comm.Distributor(dest, kernels=CpuCommKernels())
    assert d.totalNumImport() == n and d.totalNumExport() == n
    # no self-send: the neighbour list is the export rank, then import-only ranks
    # (impl/Cabana_CommunicationPlan_Mpi.hpp:305-328, :378-393)
    assert rank not in d.neighbors and d.neighbors[0] == (rank + 1) % world
    src = CpuSlice(torch.full((n, 1), float(rank)))
    dst = CpuSlice(torch.zeros((n, 1)))
    comm.migrate(d, [src], [dst])
    assert torch.all(dst.t == float((rank - 1) % world))


@pytest.mark.parametrize("world", [2, 3])
def test_slab_halo_gather_scatter_gloo(world):
    _spawn(_case_slab_halo, world)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_migrate_gloo(world):
    _spawn(_case_migrate, world)


def test_ring_distributor_gloo():
    _spawn(_case_ring_distributor, 2)


def _case_import_halo(rank, world):
    """Halo built from imports (tstHalo.hpp import-built variants; Cabana_Halo.hpp:174-330):
    every rank asks its neighbours for specific local ids; the ghosts must be exactly those
    elements, grouped by owner in neighbour order and, inside a block, in request order."""
    from _comm_double import CpuCommKernels, CpuSlice
    from cabana_b200 import comm

    num_local = 50 + 7 * rank
    nl = [50 + 7 * r for r in range(world)]
    # field value encodes (owner, local id)
    store = torch.zeros((num_local + 200, 2), dtype=torch.float64)
    store[:num_local, 0] = rank
    store[:num_local, 1] = torch.arange(num_local, dtype=torch.float64)
    rng = np.random.Generator(np.random.Philox(key=100 + rank))
    want_ranks, want_ids = [], []
    for r in range(world):
        if r == rank and world > 1:
            continue
        k = 5 + (rank + r) % 4
        want_ranks += [r] * k
        want_ids += rng.integers(0, nl[r], k).tolist()   # duplicates allowed, like a real halo
    perm = rng.permutation(len(want_ranks))               # interleave the owners in the request list
    want_ranks = [want_ranks[i] for i in perm]
    want_ids = [want_ids[i] for i in perm]
    halo = comm.Halo.from_imports(num_local, torch.tensor(want_ids, dtype=torch.int32),
                                  torch.tensor(want_ranks, dtype=torch.int32),
                                  kernels=CpuCommKernels())
    assert halo.numLocal() == num_local and halo.numGhost() == len(want_ids)
    # no topology given: export ranks ascending (self first only when it sends to itself), then
    # import-only ranks (impl/Cabana_CommunicationPlan_Mpi.hpp:305-328, :378-393)
    assert sorted(halo.neighbors) == sorted(set(halo.neighbors)) and (rank not in halo.neighbors or halo.neighborRank(0) == rank)
    x = CpuSlice(store)
    comm.gather(halo, x)
    got = store[num_local:num_local + len(want_ids)].numpy()
    # expected: neighbour order, request order inside a block
    exp = []
    for r in halo.neighbors:
        exp += [(r, i) for rr, i in zip(want_ranks, want_ids) if rr == r]
    assert got[:, 0].astype(int).tolist() == [e[0] for e in exp]
    assert got[:, 1].astype(int).tolist() == [e[1] for e in exp]
    # scatter sends ghost contributions back to the owners' elements (summed over duplicates)
    f = torch.zeros((num_local + 200, 1), dtype=torch.float64)
    f[num_local:num_local + len(want_ids)] = 1.0
    comm.scatter(halo, CpuSlice(f))
    total = torch.tensor([float(f[:num_local].sum())])
    dist.all_reduce(total)
    asked = torch.tensor([float(len(want_ids))])
    dist.all_reduce(asked)
    assert total.item() == asked.item()


@pytest.mark.parametrize("world", [2, 3])
def test_import_built_halo_gloo(world):
    _spawn(_case_import_halo, world)


def _case_import_edge(rank, world):
    """Import build corner cases: a rank that imports nothing, self-imports (allowed: the
    calling rank may be in the neighbour list), and a rejected rank of -1."""
    from _comm_double import CpuCommKernels, CpuSlice
    from cabana_b200 import comm

    num_local = 20
    store = torch.zeros((num_local + 40, 1), dtype=torch.float64)
    store[:num_local, 0] = 100.0 * rank + torch.arange(num_local, dtype=torch.float64)
    if rank == 0:
        ids, ranks = [3, 4, 19], [0, world - 1, 0]          # two self-imports and one remote
    else:
        ids, ranks = [], []                                  # imports nothing
    halo = comm.Halo.from_imports(num_local, torch.tensor(ids, dtype=torch.int32),
                                  torch.tensor(ranks, dtype=torch.int32), kernels=CpuCommKernels())
    assert halo.numGhost() == len(ids)
    comm.gather(halo, CpuSlice(store))
    if rank == 0:
        got = store[num_local:num_local + 3, 0].tolist()
        if world == 1:
            assert got == [3.0, 4.0, 19.0]
        else:   # self block first, then the remote owner
            assert got == [3.0, 19.0, 100.0 * (world - 1) + 4.0]
    with pytest.raises(ValueError):
        comm.exports_from_imports(torch.tensor([1], dtype=torch.int32), torch.tensor([-1], dtype=torch.int32))


@pytest.mark.parametrize("world", [1, 2])
def test_import_built_halo_edge_cases_gloo(world):
    _spawn(_case_import_edge, world)
