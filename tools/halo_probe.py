import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch, torch.distributed as dist
import bench
world=int(os.environ["WORLD_SIZE"]); rank=int(os.environ["RANK"]); lr=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from cabana_b200 import comm, core as cb
xyz, bounds, gmax = bench._fcc_slab(rank, world)
nl = xyz.shape[0]
slab = comm.SlabDecomposition(bounds, bench.RADIUS)
cap = nl + int(2.2*bench.RADIUS/(bounds[rank+1]-bounds[rank])*nl) + 1024
store = np.zeros((cap,3)); store[:nl]=xyz
x_all = cb.slice_from_array(store, vlen=32)
peer = slab.create_peer_halo([x_all], cap-nl)
x_own = cb.Slice(x_all.data, nl, x_all.outer_stride, x_all.vlen, x_all.comp_stride, 3)
lgx = slab.local_grid_x(); lmin=(lgx[0],0.0,0.0); lmax=(lgx[1],gmax[1],gmax[2])
lst = cb.VerletList(algorithm=cb.FULL, layout=cb.CSR)
def timeit(fn, n=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); t1=time.perf_counter()
    return e0.elapsed_time(e1)/n, (t1-t0)*1e3/n
g = timeit(lambda: peer.gather(x_own,[x_all],nl))
n_lo,n_hi = peer.gather(x_own,[x_all],nl)
x_tot = cb.Slice(x_all.data, nl+n_lo+n_hi, x_all.outer_stride, x_all.vlen, x_all.comp_stride, 3)
b = timeit(lambda: lst.build(x_tot,0,nl,bench.RADIUS,1.0,lmin,lmax))
s = timeit(lambda: peer.step(lst,x_all,[x_all],nl,bench.RADIUS,1.0,lmin,lmax))
t = torch.tensor([g[0],g[1],b[0],b[1],s[0],s[1]],device="cuda"); mx=t.clone(); dist.all_reduce(mx,op=dist.ReduceOp.MAX)
if rank==0: print("max over ranks (dev ms, wall ms): gather %.3f %.3f | build %.3f %.3f | step %.3f %.3f"%tuple(mx.tolist()))
peer.close(); dist.destroy_process_group()
