"""Probe: all-gather latency vs payload/dtype, and per-rank kernel time (diagnosing weak-scaling overheads)."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for nbytes, dt in ((6291456, torch.complex128), (6291456, torch.float64), (65536, torch.float64), (983040000 // 4, torch.float64)):
    n = nbytes // (16 if dt == torch.complex128 else 8)
    x = torch.ones(n, dtype=dt, device=dev); g = torch.empty(world * n, dtype=dt, device=dev)
    ms = timeit(lambda: dist.all_gather_into_tensor(g, x), n=30)
    if rank == 0: print("all_gather %9d B/rank %-12s %.3f ms  busbw %.1f GB/s" % (nbytes, str(dt), ms, nbytes * (world - 1) / ms / 1e6), flush=True)
import bench
class A: workload = "cfg2"; nw = 0; cases = 0; designs = 0
designs, cs, cfg = bench.build_workload(A, rank, world)
from raft_b200 import solver
sess = solver.DeviceSession(solver.DesignBatch(designs), solver.CaseTable(cs), device=dev)
ms = timeit(lambda: sess.solve(n_iter=10), n=30)
t = torch.tensor([ms], device=dev); allms = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(allms, t)
if rank == 0: print("kernel-only ms per rank:", ["%.3f" % float(v) for v in allms], flush=True)
Xi = sess.out["Xi"]; g = torch.empty((world,) + tuple(Xi.shape), dtype=Xi.dtype, device=dev)
def both():
    sess.solve(n_iter=10); dist.all_gather_into_tensor(g, Xi)
ms = timeit(both, n=30)
if rank == 0: print("solve + all_gather same stream: %.3f ms" % ms, flush=True)
dist.destroy_process_group()
