#!/usr/bin/env python
"""Quick parity probe of the two-bin fused solver (k_rao_fused2) against the oracle, small enough for compute-sanitizer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, response_err  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from raft_b200 import grid, solver  # noqa: E402

orc.build()
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nC = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
_, P = load_golden("cfg2_VolturnUS-S_nw64")
Q = grid.regrid(P, nw, 0.512)
rng = np.random.default_rng(7)
cs = dict(Hs=rng.uniform(1, 10, nC), Tp=rng.uniform(5, 18, nC), gamma=np.zeros(nC), beta_deg=rng.uniform(-180, 180, nC), spec=np.zeros(nC, dtype=np.int32))
out = solver.solve_dynamics(solver.DesignBatch(Q), solver.CaseTable(cs), n_iter=10, cluster_size=cl, want=("Xi", "status", "B_drag", "F_drag", "F_iner"))
Xi_o, st_o, _ = orc.solve_cases(orc.OracleDesign(Q), cs, nIter=10)
print("nw", nw, "cases", nC, "cluster", cl, "passes", out["status"][0, :, 0].tolist(), "oracle", st_o[:, 0].tolist(), "flags", out["status"][0, :, 2].tolist(),
      "err %.3e" % response_err(out["Xi"][0], Xi_o), "launches", solver.launch_count())
