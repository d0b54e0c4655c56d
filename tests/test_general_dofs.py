"""Generalised degrees of freedom (raft_b200/csrc/raftk_general.cuh): flexible members, nDOF = 150, through the C ABI against
the reference run of VolturnUS-S-flexible (fixture flex_VolturnUS-S-flexible) and the pinned checker.  Tolerance 1e-10
(north-star): two independent LUs of this impedance agree to ~2e-11 (cond ~1e6), pass counts must match exactly.
The checker side is pinned separately: tests/test_oracle_golden.py::test_generalised_*."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr

try:
    import torch
    HAVE_CUDA = torch.cuda.is_available()
except Exception:                                         # pragma: no cover
    HAVE_CUDA = False

pytestmark = [pytest.mark.gpu]


def test_general_solve_vs_reference_run(oracle):
    from raft_b200 import solver
    z = np.load(os.path.join(GOLDEN, "flex_VolturnUS-S-flexible.npz"))
    P = {k[2:]: z[k] for k in z.files if k.startswith("P_")}
    cs = z["ref_run_solve_cases"]
    n = len(cs)
    table = dict(Hs=cs[:, 0], Tp=cs[:, 1], gamma=np.zeros(n), beta_deg=cs[:, 2], spec=np.zeros(n, dtype=np.int32))
    Xi, st = solver.general_solve_dynamics(P, z["gen_M"], z["gen_B"], z["gen_C"], solver.CaseTable(table), n_iter=int(z["n_iter"]),
                                           xi_start=float(z["xi_start"]))
    assert np.array_equal(st[:, 0], z["ref_run_solve_passes"])
    for i in range(n):
        assert relerr(Xi[i], z["ref_run_solve_Xi"][i]) < 1e-10
    gd = oracle.GeneralDesign(P)
    Xo, so = oracle.general_solve_dynamics(gd, z["gen_M"], z["gen_B"], z["gen_C"], 0, cs[0, 0], cs[0, 1], 0.0, cs[0, 2], nIter=int(z["n_iter"]),
                                           XiStart=float(z["xi_start"]))
    assert st[0, 1] == so[1] and relerr(Xi[0], Xo) < 1e-10
