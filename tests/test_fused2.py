"""The two-bin fused solver (k_fused_plan + k_rao_fused2; runs when a CTA gets 193-256 bins) on the features the BASELINE-size
tests do not reach: ragged slices (a thread with one valid bin), MacCamy-Fuchs frequency tables, wave trains (primary /
secondary launches), continuing a loop from its own iterate (Xi_init / Xi_last), optional outputs -- all against the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, relerr, response_err

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def sea_states(seed, n):
    rng = np.random.default_rng(seed)
    return dict(Hs=rng.uniform(1, 10, n), Tp=rng.uniform(5, 18, n), gamma=np.zeros(n), beta_deg=rng.uniform(-180, 180, n),
                spec=np.zeros(n, dtype=np.int32))


@pytest.mark.parametrize("nw,cluster", [(500, 2), (1000, 4), (200, 1), (777, 0)])
def test_ragged_slices_vs_oracle(nw, cluster, oracle):
    from raft_b200 import grid, solver
    _, P = load_golden("cfg2_VolturnUS-S_nw64")
    Q = grid.regrid(P, nw, 0.512)
    cs = sea_states(21, 5)
    out = solver.solve_dynamics(solver.DesignBatch(Q), solver.CaseTable(cs), n_iter=10, cluster_size=cluster,
                                want=("Xi", "status", "B_drag", "F_drag", "F_iner", "zeta"))
    od = oracle.OracleDesign(Q)
    Xi_o, st_o, _ = oracle.solve_cases(od, cs, nIter=10)
    assert np.array_equal(out["status"][0, :, 0], st_o[:, 0]) and np.all(out["status"][0, :, 2] == 0)
    assert response_err(out["Xi"][0], Xi_o) < RTOL
    zeta, _, F_iner, u = oracle.calc_hydro_excitation(od, 0, cs["Hs"][2], cs["Tp"][2], 0.0, cs["beta_deg"][2])
    assert relerr(out["zeta"][2], zeta) < 1e-13 and relerr(out["F_iner"][0, 2], F_iner) < RTOL


def test_maccamy_fuchs_tables_on_the_two_bin_solver(oracle):
    """test_VolturnUS-S carries MacCamy-Fuchs columns: the builder recomputes the complex inertia tables on a 256-bin grid."""
    from raft_b200 import grid, solver
    from raft_b200.fowt import FOWT
    G, P = load_golden("test_VolturnUS-S")
    D = json.load(open(os.path.join(GOLDEN, "designs.json")))["test_VolturnUS-S"]
    w = grid.make_w(0.40 / 256, 0.40)
    mats = dict(M_struc=P["M0"] - G["A_hydro_morison"], C_struc=P["C0"] - G["C_moor"], C_moor=G["C_moor"])
    f = FOWT(D, w, depth=float(P["depth"]), matrices=mats)
    f.calcHydroConstants()
    Q = f.pack()
    assert Q.get("node_in_p1_w") is not None and len(Q["w"]) == 256
    cs = sea_states(22, 4)
    out = solver.solve_dynamics(solver.DesignBatch(Q), solver.CaseTable(cs), n_iter=10)
    Xi_o, st_o, _ = oracle.solve_cases(oracle.OracleDesign(Q), cs, nIter=10)
    assert np.array_equal(out["status"][0, :, 0], st_o[:, 0])
    assert response_err(out["Xi"][0], Xi_o) < RTOL


def test_wave_trains_on_the_two_bin_solver(oracle):
    from raft_b200 import grid, packer, solver
    _, P = load_golden("cfg2_VolturnUS-S_nw64")
    Q = grid.regrid(P, 256, 0.512)
    tr = np.array([[6.0, 12.0, 30.0], [2.5, 7.0, -100.0], [1.0, 16.0, 170.0]])
    case = dict(wave_spectrum=["JONSWAP"] * 3, wave_height=list(tr[:, 0]), wave_period=list(tr[:, 1]), wave_heading=list(tr[:, 2]), wave_gamma=[0.0] * 3)
    cases = [dict(wave_spectrum="JONSWAP", wave_height=2.0, wave_period=9.0, wave_heading=10.0), case]
    table, owner, first = packer.pack_case_trains(cases)
    out = solver.solve_dynamics(solver.DesignBatch(Q), solver.CaseTable(table), n_iter=10)
    od = oracle.OracleDesign(Q)
    Xo, st = oracle.solve_dynamics_trains(od, np.zeros(3, dtype=np.int32), tr[:, 0], tr[:, 1], np.zeros(3), tr[:, 2], nIter=10)
    assert response_err(out["Xi"][0, 1:4], Xo) < RTOL and out["status"][0, 1, 0] == st[0]
    assert np.array_equal(out["status"][0, 2:4, 3], [2, 2])
    solo, st1, _ = oracle.solve_cases(od, dict(Hs=[2.0], Tp=[9.0], gamma=[0.0], beta_deg=[10.0], spec=np.zeros(1, dtype=np.int32)), nIter=10)
    assert response_err(out["Xi"][0, 0], solo[0]) < RTOL


def test_continuing_a_loop_from_its_own_iterate():
    """Stop after two passes, restart from 0.2 XiLast + 0.8 Xi (raft_model.py:1133): the same iterates (up to the rounding of
    the relaxation done on the host here), the same pass counts, the same answer."""
    from raft_b200 import grid, solver
    _, P = load_golden("cfg2_VolturnUS-S_nw64")
    Q = grid.regrid(P, 256, 0.512)
    cs = sea_states(23, 3)
    b = solver.DesignBatch(Q)
    full = solver.solve_dynamics(b, solver.CaseTable(cs), n_iter=10)
    a = solver.solve_dynamics(b, solver.CaseTable(cs), n_iter=1, want=("Xi", "status", "Xi_last"))
    assert np.all(a["status"][0, :, 0] == 2) and np.all(a["status"][0, :, 1] == 0)
    nxt = 0.2 * a["Xi_last"] + 0.8 * a["Xi"]
    c = solver.solve_dynamics(b, solver.CaseTable(cs, Xi_init=nxt), n_iter=8)
    assert np.array_equal(c["status"][0, :, 0] + 2, full["status"][0, :, 0])
    assert response_err(c["Xi"][0], full["Xi"][0]) < 1e-12


@pytest.mark.parametrize("name", ["cfg2_VolturnUS-S_nw64", "cfg1_OC3spar"])
def test_drag_direction_masks_vs_oracle(name, oracle):
    """k_fused_plan's per-chunk direction masks (axial-only / transverse-only / both / none) and the accumulator slots they
    select: the drag columns are edited so that every combination occurs inside one chunk and across chunk boundaries."""
    from raft_b200 import grid, solver
    _, P = load_golden(name)
    Q = grid.regrid(P, 256, 0.40)
    Q = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in Q.items()}
    # the raw columns (what the checker reads) are edited and the derived ones (what the kernels read) recomputed exactly as
    # packer.py:110-112 does: cd_q = pref (a_q Cd_q + a_End Cd_End), cd_p = pref a_p Cd_p, pref = sqrt(8/pi) rho / 2
    pref = np.sqrt(8.0 / np.pi) * 0.5 * float(Q["rho"])
    n = len(Q["node_cd_q"])
    for key in ("node_a_q", "node_Cd_q", "node_a_End", "node_Cd_End", "node_a_p1", "node_Cd_p1", "node_a_p2", "node_Cd_p2"):
        assert len(Q[key]) == n
    a_ref = float(np.max(Q["node_a_p1"]))
    for j in range(n):
        kind = j % 5
        if kind == 0:                                   # both: axial and transverse drag on the same node (third slot)
            Q["node_a_End"][j], Q["node_Cd_End"][j] = 0.5 * a_ref, 0.6
            Q["node_a_p1"][j], Q["node_Cd_p1"][j], Q["node_a_p2"][j], Q["node_Cd_p2"][j] = a_ref, 0.8, a_ref, 0.8
        elif kind == 1:                                 # nothing active
            for key in ("node_Cd_q", "node_Cd_End", "node_Cd_p1", "node_Cd_p2"):
                Q[key][j] = 0.0
        elif kind == 2:                                 # one transverse coefficient only (the other is an exact zero)
            Q["node_Cd_q"][j] = Q["node_Cd_End"][j] = Q["node_Cd_p1"][j] = 0.0
            Q["node_a_p2"][j], Q["node_Cd_p2"][j] = a_ref, 0.8
        # kinds 3, 4 keep the design's own pattern (axial-only ends, transverse-only strips)
    Q["node_cd_q"] = pref * (Q["node_a_q"] * Q["node_Cd_q"] + Q["node_a_End"] * Q["node_Cd_End"])
    Q["node_cd_p1"] = pref * Q["node_a_p1"] * Q["node_Cd_p1"]
    Q["node_cd_p2"] = pref * Q["node_a_p2"] * Q["node_Cd_p2"]
    both = (Q["node_cd_q"] != 0) & ((Q["node_cd_p1"] != 0) | (Q["node_cd_p2"] != 0))
    none = (Q["node_cd_q"] == 0) & (Q["node_cd_p1"] == 0) & (Q["node_cd_p2"] == 0)
    assert both.any() and none.any()
    cs = sea_states(24, 4)
    out = solver.solve_dynamics(solver.DesignBatch(Q), solver.CaseTable(cs), n_iter=10, want=("Xi", "status", "B_drag", "F_drag"))
    od = oracle.OracleDesign(Q)
    Xi_o, st_o, _ = oracle.solve_cases(od, cs, nIter=10)
    assert np.array_equal(out["status"][0, :, 0], st_o[:, 0]) and np.all(out["status"][0, :, 2] == 0)
    assert response_err(out["Xi"][0], Xi_o) < RTOL
