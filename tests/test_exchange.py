"""The multi-GPU exchange fused into the solve kernel (include/raftk.h raftk_peers), exercised on ONE GPU: two emulated
ranks in one process, each with its own gathered copy (plain raftk_peer_alloc memory, no IPC needed inside one
process) and its own stream.  Rank r's kernel must deliver its block into BOTH copies, the arrival barrier must
order the streams, and the result must equal the single-rank solve of the concatenated case table bit for bit.
(The cross-process version -- CUDA IPC handles over torch.distributed -- runs in bench.py at N > 1.)"""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def sea_states(seed, n):
    rng = np.random.default_rng(seed)
    return dict(Hs=rng.uniform(1, 10, n), Tp=rng.uniform(5, 18, n), gamma=np.zeros(n), beta_deg=rng.uniform(-180, 180, n),
                spec=np.zeros(n, dtype=np.int32))


def test_fused_exchange_two_emulated_ranks():
    import torch
    from raft_b200 import grid, solver, sweep
    from raft_b200._lib import RaftkPeers, check, lib
    _, P = load_golden("cfg2_VolturnUS-S_nw64")
    Q = grid.regrid(P, 256, 0.512)
    world, nC, nw = 2, 6, 256
    cs_all = sea_states(5, world * nC)
    ref = solver.solve_dynamics(solver.DesignBatch(Q), solver.CaseTable(cs_all), n_iter=10)

    dev = torch.device("cuda", 0)
    block = nC * 6 * nw
    xi_bytes = world * block * 16
    off_flags = (xi_bytes + 255) // 256 * 256
    off_status = off_flags + 256
    total = off_status + world * nC * 16
    ptrs = []
    for _ in range(world):
        p, h = C.c_void_p(), C.create_string_buffer(64)
        check(lib.raftk_peer_alloc(total, C.byref(p), h))
        ptrs.append(p.value)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    timeout = torch.zeros(1, dtype=torch.int32, device=dev)
    views, sessions = [], []
    for r in range(world):
        raw = torch.as_tensor(sweep._DevMem(ptrs[r], total), device=dev)
        g = torch.view_as_complex(raw[:xi_bytes].view(torch.float64).view(-1, 2)).view(world, 1, nC, 6, nw)
        s = raw[off_status:off_status + world * nC * 16].view(torch.int32).view(world, 1, nC, 4)
        views.append((g, s))
        cs = {k: v[r * nC:(r + 1) * nC] for k, v in cs_all.items()}
        sessions.append(solver.DeviceSession(solver.DesignBatch(Q), solver.CaseTable(cs), device=dev,
                                             out_tensors=dict(Xi=g[r], status=s[r])))
    for epoch in (1, 2):                                     # two steps: the flags count epochs
        for r in range(world):
            pr = RaftkPeers()
            pr.n_ranks, pr.rank, pr.epoch, pr.block_elems = world, r, epoch, block
            for q in range(world):
                pr.gathered[q], pr.flags[q], pr.status[q] = ptrs[q], ptrs[q] + off_flags, ptrs[q] + off_status
            with torch.cuda.stream(streams[r]):
                sessions[r].solve_gather(pr, n_iter=10, timeout_flag=timeout.data_ptr())
        torch.cuda.synchronize()
        assert timeout.item() == 0
        for r in range(world):
            g, s = views[r]
            got = g.cpu().numpy().reshape(world * nC, 6, nw)
            assert np.array_equal(got, ref["Xi"][0]), "copy of rank %d differs from the single-rank solve" % r
            assert np.array_equal(s.cpu().numpy().reshape(world * nC, 4), ref["status"][0])
        for g, s in views:
            g.zero_(); s.zero_()
        torch.cuda.synchronize()
    del views, sessions
    for p in ptrs:
        check(lib.raftk_peer_free(p))


def test_sharded_solve_single_rank_matches_plain():
    """ShardedSolve without a process group degenerates to the plain solve (double-buffered outputs)."""
    import torch
    from raft_b200 import solver, sweep
    _, P = load_golden("cfg2_VolturnUS-S_nw64")
    cs = sea_states(9, 4)
    ref = solver.solve_dynamics(solver.DesignBatch(P), solver.CaseTable(cs), n_iter=10)
    sh = sweep.ShardedSolve([P], cs)
    for _ in range(3):
        g, s = sh.step(n_iter=10)
        torch.cuda.synchronize()
        assert np.array_equal(g[0].cpu().numpy(), ref["Xi"]) and np.array_equal(s[0].cpu().numpy(), ref["status"])
    xi_h, st_h, h2d, d2h = sh.step_host(n_iter=10)
    assert np.array_equal(xi_h.numpy(), ref["Xi"]) and np.array_equal(st_h.numpy(), ref["status"]) and h2d > 0 and d2h > 0
    sh.close()


def test_second_device_has_its_own_library_state():
    """A process that drives two GPUs (DeviceSession(device=...)): shared-memory opt-ins, arenas and scratch are per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    from raft_b200 import grid, solver
    _, P = load_golden("cfg2_VolturnUS-S_nw64")
    Q = grid.regrid(P, 256, 0.512)
    cs = sea_states(3, 3)
    batch, cases = solver.DesignBatch(Q), solver.CaseTable(cs)
    outs = []
    for d in (0, 1):
        sess = solver.DeviceSession(batch, cases, device=torch.device("cuda", d))
        o = sess.solve(n_iter=10)
        torch.cuda.synchronize(d)
        outs.append((o["Xi"].cpu().numpy(), o["status"].cpu().numpy()))
        with torch.cuda.device(d):
            host = solver.solve_dynamics(batch, cases, n_iter=10)              # arena of the current device
        assert np.array_equal(host["Xi"], outs[-1][0])
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
