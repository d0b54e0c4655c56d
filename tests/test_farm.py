"""Coupled 6N-DOF farm response (BASELINE.json configs[4]; raft_model.py:1164-1236) on the device, pinned to a run of the
UNMODIFIED reference: designs/VolturnUS-S_farm.yaml (two FOWTs 1600 m apart, the first turned by 180 deg) with the
SURVEY.md 8c recipe -- array-level mooring replaced by a seeded SPD 12 x 12 stiffness injected through
model.ms.getCoupledStiffnessA (fixture farm_VolturnUS-S_farm_nw48, tests/golden/make_golden.py:fixture_farm)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, response_err

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _farm_fixture():
    z = np.load(os.path.join(GOLDEN, "farm_VolturnUS-S_farm_nw48.npz"))
    N = int(z["n_fowt"])
    packs = [{k[len("P%d_" % i):]: z[k] for k in z.files if k.startswith("P%d_" % i)} for i in range(N)]
    return z, packs


def _cases(rows):
    n = len(rows)
    return dict(Hs=rows[:, 0], Tp=rows[:, 1], gamma=np.zeros(n), beta_deg=rows[:, 2], spec=np.zeros(n, dtype=np.int32))


def test_farm_response_vs_reference_run():
    from raft_b200 import solver
    z, packs = _farm_fixture()
    N, n = len(packs), 6 * len(packs)
    out = solver.solve_dynamics_farm(solver.DesignBatch(packs), solver.CaseTable(_cases(z["cases"])), C_arr=z["C_array"],
                                     n_iter=int(z["n_iter"]), xi_start=float(z["xi_start"]))
    assert not np.any(out["info"]) and np.all(out["status"][..., 2] == 0)
    assert np.array_equal(out["status"][:, :, 0].T, z["ref_run_passes"])                 # passes per (case, FOWT)
    ref = z["ref_run_Xi"][:, 0]                                                             # [nCases, 12, nw] (wave train 0)
    assert out["Xi_sys"].shape == ref.shape
    err = max(response_err(out["Xi_sys"][:, 6 * i:6 * i + 6], ref[:, 6 * i:6 * i + 6]) for i in range(N))
    assert err < RTOL, err
    # caller-owned page-locked result arrays (the e2e path of bench.py --workload farm): the same numbers in the same arrays
    nC, nw = len(z["cases"]), len(packs[0]["w"])
    pin = dict(Xi=solver.pinned_empty([N, nC, 6, nw], np.complex128), status=solver.pinned_empty([N, nC, 4], np.int32),
               Xi_sys=solver.pinned_empty([nC, n, nw], np.complex128), info=solver.pinned_empty([nC, nw], np.int32))
    o2 = solver.solve_dynamics_farm(solver.DesignBatch(packs), solver.CaseTable(_cases(z["cases"])), C_arr=z["C_array"],
                                    n_iter=int(z["n_iter"]), xi_start=float(z["xi_start"]), out=pin)
    assert o2["Xi_sys"] is pin["Xi_sys"] and np.array_equal(o2["Xi_sys"], out["Xi_sys"]) and np.array_equal(o2["Xi"], out["Xi"])
    assert np.array_equal(o2["status"], out["status"])
    # without the coupling stiffness the system response is the per-FOWT response
    unc = solver.solve_dynamics_farm(solver.DesignBatch(packs), solver.CaseTable(_cases(z["cases"])), n_iter=int(z["n_iter"]),
                                     xi_start=float(z["xi_start"]))
    per = np.concatenate([unc["Xi"][i] for i in range(N)], axis=1)                          # [nC, 12, nw]
    assert max(response_err(unc["Xi_sys"][:, 6 * i:6 * i + 6], per[:, 6 * i:6 * i + 6]) for i in range(N)) < RTOL


def test_farm_model_api_vs_reference_run():
    from raft_b200.model import Model
    z, packs = _farm_fixture()
    D = json.load(open(os.path.join(GOLDEN, "designs.json")))["farm_VolturnUS-S_farm_nw48"]
    design = dict(settings=D["settings"], site=D["site"], platform=D["platform"], array=D["array"])
    mats = [dict(M_struc=P["M0"] - z["A_hydro_morison%d" % i], C_struc=P["C0"] - z["C_moor%d" % i], C_moor=z["C_moor%d" % i])
            for i, P in enumerate(packs)]
    model = Model(design, matrices=mats, array_stiffness=z["C_array"])
    assert model.nFOWT == 2 and model.nDOF == 12 and model.nw == 48
    for ic, (Hs, Tp, beta) in enumerate(z["cases"]):
        Xi = model.solveDynamics(dict(wave_spectrum="JONSWAP", wave_height=Hs, wave_period=Tp, wave_heading=beta))
        ref = z["ref_run_Xi"][ic]
        assert Xi.shape == ref.shape and np.all(Xi[-1] == 0)
        assert max(response_err(Xi[0, 6 * i:6 * i + 6], ref[0, 6 * i:6 * i + 6]) for i in range(2)) < RTOL
        assert np.array_equal(model.fowtList[1].Xi[0], Xi[0, 6:12])


def test_farm_baseline_size_vs_oracle(oracle):
    """configs[4] at its stated size: N = 2, 1024 bins (max_freq 0.1024 Hz), against the oracle's per-FOWT solves + the
    explicit-inverse system response (raft_model.py:1189-1216)."""
    from raft_b200 import grid, solver
    z, packs = _farm_fixture()
    Q = [grid.regrid(P, 1024, 0.1024) for P in packs]
    cs = _cases(np.array([[6.0, 12.0, 0.0], [4.0, 9.0, 35.0]]))
    out = solver.solve_dynamics_farm(solver.DesignBatch(Q), solver.CaseTable(cs), C_arr=z["C_array"], n_iter=10)
    nw = 1024
    for c in range(2):
        Zs = np.zeros([nw, 12, 12], dtype=complex)
        F = np.zeros([nw, 12], dtype=complex)
        for i, P in enumerate(Q):
            Xi_i, st, Z_i, _ = oracle.solve_dynamics(oracle.OracleDesign(P), 0, cs["Hs"][c], cs["Tp"][c], 0.0, cs["beta_deg"][c], nIter=10, want_Z=True)
            assert st[0] == out["status"][i, c, 0]
            Zs[:, 6 * i:6 * i + 6, 6 * i:6 * i + 6] = Z_i
            F[:, 6 * i:6 * i + 6] = np.einsum("wab,bw->wa", Z_i, Xi_i)
        Xo = oracle.system_response(Zs + z["C_array"][None], F).T
        err = max(response_err(out["Xi_sys"][c, 6 * i:6 * i + 6], Xo[6 * i:6 * i + 6]) for i in range(2))
        assert err < 1e-9, err           # the checker goes through the explicit inverse like the reference (cond ~1e5)


@pytest.mark.parametrize("N", [4, 5, 9])
def test_farm_larger_arrays_vs_oracle(N, oracle):
    """6N = 24, 30 (warp-per-system kernel, 4 resp. 3 systems per CTA) and 6N = 54 (blocked LU, one CTA per system) against the oracle's
    per-FOWT solves + explicit-inverse system response."""
    import bench_extra
    from raft_b200 import solver
    packs, C_arr, _ = bench_extra.farm_designs(N, nw=96, max_freq=0.1024)
    cs = _cases(np.array([[6.0, 12.0, 0.0], [3.0, 8.0, -70.0]]))
    out = solver.solve_dynamics_farm(solver.DesignBatch(packs), solver.CaseTable(cs), C_arr=C_arr, n_iter=10)
    Xo, passes = bench_extra._oracle_farm(packs, C_arr, cs)
    assert np.array_equal(passes, out["status"][:, :, 0]) and not np.any(out["info"])
    err = max(response_err(out["Xi_sys"][:, 6 * i:6 * i + 6], Xo[:, 6 * i:6 * i + 6]) for i in range(N))
    assert err < 1e-9, err
