#!/usr/bin/env python
"""Where does a multi-GPU step spend its time?  torchrun --nproc-per-node N tools/peer_probe.py
Times, per rank, the solve kernel(s) and the arrival barrier separately (CUDA events), with and without the L2 flush
between steps, for the fused peer-store exchange."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raft_b200 import solver, sweep  # noqa: E402
from raft_b200._lib import RaftkSolveOpts, check, lib  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
a = argparse.Namespace(workload="cfg2", nw=0, cases=0, designs=0)
designs, cs, cfg = bench.build_workload(a, rank, world)
sh = sweep.ShardedSolve(designs, cs, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream


def run(n, do_flush, do_barrier=True, sleep_first=0.0):
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    for e in ev:
        if do_flush:
            flush.fill_(1)
        b, peers = sh.px.next()
        o_struct, _ = sh.o_structs[b]
        o = sh.sess._opts(10, 0.01, 0.0, 0)
        e[0].record()
        check(lib.raftk_solve_dynamics_gather_dev(C.byref(sh.sess.d_struct), C.byref(sh.sess.c_struct), C.byref(o), C.byref(o_struct), C.byref(peers),
                                                  sh.sess.workspace.data_ptr(), sh.sess.workspace_bytes, st))
        e[1].record()
        check(lib.raftk_peer_barrier_dev(C.byref(peers), sh.px.timeout.data_ptr(), st))
        e[2].record()
    torch.cuda.synchronize()
    s = np.array([[e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])] for e in ev])
    return s


for name, fl in (("warm", False), ("no flush", False), ("flush", True), ("no flush again", False), ("flush again", True)):
    dist.barrier()
    s = run(20, fl)
    t = torch.tensor(s[5:].mean(axis=0), device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        print(name, "per rank [solve ms, barrier ms]:", [[round(float(x), 4) for x in v.tolist()] for v in allt], flush=True)
assert not sh.timed_out()
sh.close()
dist.destroy_process_group()
