"""STAGED kernels (raft_b200/csrc/raftk_general.cuh): generalised degrees of freedom, flexible members, nDOF = 150.

The kernels were written after the round's GPU budget was spent and have NEVER RUN ON HARDWARE.  During development they were
executed as written under a host emulation (tools/host_emu: CUDA threads as std::threads, __syncthreads / shuffles as
barriers; log profiles/r01_host_emu_general.txt): on this
fixture that reproduced the reference run to 1.5e-11 with identical pass counts, and the checker's F_iner / B_drag / F_drag to
5e-16.  What has not been exercised is the GPU itself, so the test is marked xfail(strict=False): an XPASS at the round-end run
means the row is built and parity-green, an XFAIL means debugging starts here next round.  It is the last file of the suite on
purpose.  The checker side is pinned separately: tests/test_oracle_golden.py::test_generalised_*."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr

try:
    import torch
    HAVE_CUDA = torch.cuda.is_available()
except Exception:                                         # pragma: no cover
    HAVE_CUDA = False

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="staged: validated by host emulation only, first hardware run"),
              pytest.mark.skipif(not HAVE_CUDA, reason="needs a CUDA device")]


def test_staged_general_solve_vs_reference_run(oracle):
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
        assert relerr(Xi[i], z["ref_run_solve_Xi"][i]) < 1e-9
    gd = oracle.GeneralDesign(P)
    Xo, so = oracle.general_solve_dynamics(gd, z["gen_M"], z["gen_B"], z["gen_C"], 0, cs[0, 0], cs[0, 1], 0.0, cs[0, 2], nIter=int(z["n_iter"]),
                                           XiStart=float(z["xi_start"]))
    assert st[0, 1] == so[1] and relerr(Xi[0], Xo) < 1e-9
