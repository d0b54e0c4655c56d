"""WAMIT readers + readHydro (raft_b200/bem.py) against the reference's FOWT.readHydro outputs (CPU)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, relerr


def test_wamit_readers_roundtrip(tmp_path):
    from raft_b200 import bem
    rng = np.random.default_rng(0)
    T = np.array([-1.0, 0.0, 4.0, 8.0, 16.0])
    with open(tmp_path / "x.1", "w") as f:
        for t in T[::-1]:
            for i in range(1, 7):
                for j in range(1, 7):
                    f.write("%e %d %d %e %e\n" % (t, i, j, rng.normal(), rng.normal()))
    with open(tmp_path / "x.3", "w") as f:
        for t in T[2:]:
            for h in (0.0, 90.0, 180.0):
                for i in range(1, 7):
                    re, im = rng.normal(), rng.normal()
                    f.write("%e %e %d %e %e %e %e\n" % (t, h, i, np.hypot(re, im), np.degrees(np.arctan2(im, re)), re, im))
    A, B, w1 = bem.read_wamit1(str(tmp_path / "x.1"))
    assert A.shape == (6, 6, 5) and np.isinf(w1[1]) and w1[0] < 0 and np.allclose(w1[2:], 2 * np.pi / T[2:])
    mod, pha, re, im, w3, heads = bem.read_wamit3(str(tmp_path / "x.3"))
    assert re.shape == (3, 6, 3) and list(heads) == [0.0, 90.0, 180.0]
    assert np.allclose(mod, np.hypot(re, im), rtol=1e-6)
    H = bem.read_hydro_files(str(tmp_path / "x"), w=np.array([0.1, 0.5, 1.0, 1.5]))
    assert H["A_BEM"].shape == (6, 6, 4) and H["X_BEM"].shape == (3, 6, 4)
    assert np.allclose(H["A_BEM"][3:, :3], np.swapaxes(H["A_BEM"][:3, 3:], 0, 1))        # symmetrised like helpers.py:580
    with pytest.raises(ValueError):
        bem.read_hydro_files(str(tmp_path / "x"), w=np.array([0.1, 5.0]))                  # beyond the table (interp1d bounds error)


@pytest.mark.parametrize("name", ["cfg3_OC4semi-WAMIT_nw128", "test_OC4semi-WAMIT_Coefs"])
def test_read_hydro_vs_reference(name):
    """A_BEM, B_BEM, X_BEM of the reference's own readHydro (run under the stub harness; the test_ fixture's grid is the
    one of the reference's golden OC4semi-WAMIT_Coefs_true_BEM_forces.pkl) from the raw marin_semi tables."""
    from raft_b200 import bem
    z = np.load(os.path.join(GOLDEN, "wamit_marin_semi.npz"))
    G, P = load_golden(name)
    H = bem.read_hydro(z["A"], z["B"], z["w1"], z["Re"], z["Im"], z["w3"], z["heads"], P["w"], rho=float(P["rho"]), g=float(P["g"]))
    assert relerr(H["A_BEM"], P["A_w"].reshape(6, 6, -1)) < 1e-14
    assert relerr(H["B_BEM"], P["B_w"].reshape(6, 6, -1)) < 1e-14
    assert relerr(H["X_BEM"], P["X_BEM"]) < 1e-14
    assert np.array_equal(H["BEM_headings"], P["bem_headings"])


def test_qtf_reader_vs_reference_readQTF(tmp_path):
    """bem.read_qtf on the shipped marin_semi.12d rows == the state FOWT.readQTF left in the reference run."""
    from conftest import QTF_GOLDEN
    from raft_b200 import bem
    rows = np.load(os.path.join(GOLDEN, "wamit_marin_semi.npz"))["qtf_rows"]
    path = str(tmp_path / "semi.12d")
    np.savetxt(path, rows, fmt="%.5e")                     # the file carries 6 significant digits
    qtf, w, heads = bem.read_qtf(path, rho=1025.0, g=9.81)
    G, P = load_golden(QTF_GOLDEN)
    assert qtf.shape == P["qtf"].shape == (56, 56, 1, 6)
    assert np.array_equal(w, P["qtf_w"]) and np.array_equal(heads, P["qtf_heads"])
    assert np.array_equal(qtf, P["qtf"])
    off = ~np.eye(56, dtype=bool)                          # Hermitian fill of the other triangle (diagonal as read)
    assert np.array_equal(qtf[:, :, 0, :][off], np.conj(np.swapaxes(qtf[:, :, 0, :], 0, 1))[off])


def test_qtf_reader_errors(tmp_path):
    from raft_b200 import bem
    p = str(tmp_path / "bad.12d")
    with open(p, "w") as f:
        f.write("10 10 0 30 1 1 0 1 0\n")
    with pytest.raises(ValueError, match="unidirectional"):
        bem.read_qtf(p)
    with open(p, "w") as f:
        f.write("10 10 0 0 1 1 0 1 0\n10 8 0 0 1 1 0 1 0\n")
    with pytest.raises(ValueError, match="frequency columns"):
        bem.read_qtf(p)
