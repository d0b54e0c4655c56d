"""GPU parity at the sizes BASELINE.json quotes: the shapes bench.py times, not scaled-down
stand-ins.  Workloads come from bench.build_workload, so the arrays compared here are the arrays benchmarked.

  cfg2  VolturnUS-S strip theory, 1024 bins x 64 sea states (seed 2)   -> 8-CTA clusters, two waves of CTAs
  cfg3  OC4semi + WAMIT tables,   2048 bins x 256 sea states (seed 3)
  sweep 64 geometry variants (seed 40) x 16 sea states (seed 4) x 512 bins (a sample of the 10 000-design shard)
  farm  VolturnUS-S_farm, N = 2, 1024 bins (tests/test_farm.py)

Checker: the pinned C oracle (all host threads).  Tolerance 1e-10 on conftest.response_err, identical pass counts."""
import argparse

import numpy as np
import pytest

from conftest import response_err

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _workload(name, **kw):
    import bench
    a = argparse.Namespace(workload=name, nw=0, cases=0, designs=0)
    for k, v in kw.items():
        setattr(a, k, v)
    return bench.build_workload(a, 0, 1)


def _check(designs, cs, out, oracle):
    import os
    worst, mism = 0.0, 0
    for d, P in enumerate(designs):
        Xi_o, st_o, _ = oracle.solve_cases(oracle.OracleDesign(P), cs, nIter=10, nthreads=os.cpu_count() or 1)
        mism += int(np.sum(out["status"][d, :, 0] != st_o[:, 0]) + np.sum(out["status"][d, :, 1] != st_o[:, 1]))
        worst = max(worst, response_err(out["Xi"][d], Xi_o))
    return worst, mism


@pytest.mark.parametrize("cluster", [0, 4])
def test_cfg2_full_size_vs_oracle(cluster, oracle):
    import torch
    from raft_b200 import solver
    designs, cs, cfg = _workload("cfg2")
    assert cfg["nw"] == 1024 and cfg["cases_per_gpu"] == 64
    batch, cases = solver.DesignBatch(designs), solver.CaseTable(cs)
    host = solver.solve_dynamics(batch, cases, n_iter=10, cluster_size=cluster)
    worst, mism = _check(designs, cs, host, oracle)
    assert mism == 0 and worst < RTOL, (worst, mism)
    assert np.all(host["status"][..., 2] == 0)
    # the HBM-resident route bench.py times as `value` gives the same bits
    sess = solver.DeviceSession(batch, cases)
    dev = sess.solve(n_iter=10, cluster_size=cluster)
    torch.cuda.synchronize()
    assert np.array_equal(dev["Xi"].cpu().numpy(), host["Xi"]) and np.array_equal(dev["status"].cpu().numpy(), host["status"])


def test_cfg3_full_size_vs_oracle(oracle):
    from raft_b200 import solver
    designs, cs, cfg = _workload("cfg3")
    assert cfg["nw"] == 2048 and cfg["cases_per_gpu"] == 256
    out = solver.solve_dynamics(solver.DesignBatch(designs), solver.CaseTable(cs), n_iter=10)
    worst, mism = _check(designs, cs, out, oracle)
    assert mism == 0 and worst < RTOL, (worst, mism)


def test_sweep_sample_full_grid_vs_oracle(oracle):
    from raft_b200 import solver
    designs, cs, cfg = _workload("sweep", designs=64)
    assert cfg["nw"] == 512 and cfg["cases_per_gpu"] == 16 and len(designs) == 64
    # solved: the tables of the BATCHED builder (what bench.py times); checked: the oracle on the per-design builder's tables
    out = solver.solve_dynamics(designs.batch, solver.CaseTable(cs), n_iter=10)
    worst, mism = _check(designs, cs, out, oracle)
    assert mism == 0 and worst < RTOL, (worst, mism)
