"""First-order potential-flow coefficient tables: WAMIT .1/.3 readers and the ``readHydro`` interpolation.

SURVEY.md section 8(f) row 3: a pyHAMS-free production reader so that designs with ``potFirstOrder: 1`` run from
their shipped data files.  Restates raft_fowt.py:1444-1510 (``FOWT.readHydro``) and the two pyHAMS readers it
calls (``pyhams.pyhams.read_wamit1/read_wamit3`` with ``TFlag=True``; pyHAMS is a third-party dependency that
is not vendored in the reference -- its readers' contract is pinned by the reference's golden
``OC4semi-WAMIT_Coefs_true_BEM_forces.pkl`` through ``FOWT.readHydro``; see tests/test_bem.py).

Output layouts are the reference's: ``A_BEM, B_BEM [6,6,nw]``, ``X_BEM [nhead,6,nw]`` (heading-relative),
``BEM_headings [nhead]`` in degrees, ascending in [0, 360).

Second order: ``read_qtf`` restates ``FOWT.readQTF`` (raft_fowt.py:2081-2128) for WAMIT ``.12d`` difference-frequency
QTF files (potSecOrder 2).
"""
import numpy as np


def read_wamit1(path, TFlag=True):
    """WAMIT .1 (added mass / damping): rows ``T i j A [B]``.  -> A[6,6,nT], B[6,6,nT], w[nT] ordered by
    ascending period (so descending frequency), ``w = 2 pi / T`` when ``TFlag``."""
    rows = np.loadtxt(path, ndmin=2)
    T = np.unique(rows[:, 0])
    A = np.zeros([6, 6, len(T)])
    B = np.zeros([6, 6, len(T)])
    it = np.searchsorted(T, rows[:, 0])
    i, j = rows[:, 1].astype(int) - 1, rows[:, 2].astype(int) - 1
    A[i, j, it] = rows[:, 3]
    if rows.shape[1] > 4:
        B[i, j, it] = rows[:, 4]
    with np.errstate(divide="ignore"):
        w = 2.0 * np.pi / T if TFlag else T
    return A, B, w


def read_wamit3(path, TFlag=True):
    """WAMIT .3 (excitation): rows ``T heading i mod phase re im``.
    -> mod, phase, re, im [nhead,6,nT], w[nT], headings[nhead] (ascending unique values)."""
    rows = np.loadtxt(path, ndmin=2)
    T, H = np.unique(rows[:, 0]), np.unique(rows[:, 1])
    out = [np.zeros([len(H), 6, len(T)]) for _ in range(4)]
    it, ih = np.searchsorted(T, rows[:, 0]), np.searchsorted(H, rows[:, 1])
    i = rows[:, 2].astype(int) - 1
    for c, arr in enumerate(out):
        arr[ih, i, it] = rows[:, 3 + c]
    with np.errstate(divide="ignore"):
        w = 2.0 * np.pi / T if TFlag else T
    return out[0], out[1], out[2], out[3], w, H


def _alternator(r):
    return np.array([[0, r[2], -r[1]], [-r[2], 0, r[0]], [r[1], -r[0], 0]])


def translate_matrix_6to6(M, r):
    """6x6 matrix about a reference point translated by r (helpers.py:563-585).  Like the reference, the
    lower-left block is the transpose of the (translated) upper-right block -- also for r = 0, which
    symmetrises slightly asymmetric WAMIT coupling terms."""
    H = _alternator(r)
    out = np.zeros([6, 6])
    out[:3, :3] = M[:3, :3]
    out[:3, 3:] = M[:3, :3] @ H + M[:3, 3:]
    out[3:, :3] = out[:3, 3:].T
    out[3:, 3:] = H @ M[:3, :3] @ H.T + M[3:, :3] @ H + H.T @ M[:3, 3:] + M[3:, 3:]
    return out


def _interp_last_axis(x_tab, y_tab, x):
    """Linear interpolation along the last axis, table abscissae in any order (scipy interp1d, assume_sorted=False)."""
    order = np.argsort(x_tab)
    xs, ys = np.asarray(x_tab)[order], np.asarray(y_tab)[..., order]
    if np.any(x < xs[0]) or np.any(x > xs[-1]):
        raise ValueError("A value in the model's frequency grid is outside the range of the hydrodynamic data.")
    hi = np.clip(np.searchsorted(xs, x, side="left"), 1, len(xs) - 1)
    lo = hi - 1
    slope = (ys[..., hi] - ys[..., lo]) / (xs[hi] - xs[lo])
    return slope * (x - xs[lo]) + ys[..., lo]


def read_hydro(A, B, w1, Re, Im, w3, heads, w, rho=1025.0, g=9.81, r_ref0=(0.0, 0.0, 0.0), r_ref=(0.0, 0.0, 0.0)):
    """``FOWT.readHydro`` on tables already read from the .1/.3 files (raft_fowt.py:1459-1501).

    The first two .1 entries are dropped and entry 0 is reused as the zero-frequency added mass (:1470-1471:
    they are expected to be the T = -1/0 rows); damping and excitation are padded with zero at w = 0.
    Returns dict(A_BEM, B_BEM, X_BEM, BEM_headings)."""
    w = np.asarray(w, dtype=float)
    heads = np.asarray(heads, dtype=float) % 360
    order = np.argsort(heads)
    heads, Re, Im = heads[order], np.asarray(Re)[order], np.asarray(Im)[order]
    nh = len(heads)
    Ai = _interp_last_axis(np.hstack([w1[2:], 0.0]), np.dstack([A[:, :, 2:], A[:, :, 0]]), w)
    Bi = _interp_last_axis(np.hstack([w1[2:], 0.0]), np.dstack([B[:, :, 2:], np.zeros([6, 6])]), w)
    Ri = _interp_last_axis(np.hstack([w3, 0.0]), np.dstack([Re, np.zeros([nh, 6])]), w)
    Ii = _interp_last_axis(np.hstack([w3, 0.0]), np.dstack([Im, np.zeros([nh, 6])]), w)
    r0 = -np.asarray(r_ref0, dtype=float)
    A_BEM = np.stack([translate_matrix_6to6(rho * Ai[:, :, i], r0) for i in range(len(w))], axis=2)
    B_BEM = np.stack([translate_matrix_6to6(w[i] * rho * Bi[:, :, i], r0) for i in range(len(w))], axis=2)
    Xt = rho * g * (Ri + 1j * Ii)
    s, c = np.sin(np.radians(heads))[:, None], np.cos(np.radians(heads))[:, None]
    X = np.zeros_like(Xt)
    X[:, 0] = c * Xt[:, 0] + s * Xt[:, 1]
    X[:, 1] = -s * Xt[:, 0] + c * Xt[:, 1]
    X[:, 2] = Xt[:, 2]
    X[:, 3] = c * Xt[:, 3] + s * Xt[:, 4]
    X[:, 4] = -s * Xt[:, 3] + c * Xt[:, 4]
    X[:, 5] = Xt[:, 5]
    off = -np.asarray(r_ref, dtype=float)
    if np.any(off):                                        # transformForce(offset=...): add offset x force to the moments
        X[:, 3:] += np.cross(off[None, None, :], np.moveaxis(X[:, :3], 1, 2)).transpose(0, 2, 1)
    for name, arr in (("added mass", A_BEM), ("damping", B_BEM), ("excitation", X)):
        if np.isnan(arr).any():
            raise Exception("NaN values detected in HAMS calculations for %s. Check the geometry." % name)
    return dict(A_BEM=A_BEM, B_BEM=B_BEM, X_BEM=X, BEM_headings=heads)


def read_hydro_files(hydro_path, w, **kw):
    """readHydro straight from ``<hydro_path>.1`` / ``.3``."""
    A, B, w1 = read_wamit1(hydro_path + ".1", TFlag=True)
    _, _, Re, Im, w3, heads = read_wamit3(hydro_path + ".3", TFlag=True)
    return read_hydro(A, B, w1, Re, Im, w3, heads, w, **kw)


def read_qtf(path, rho=1025.0, g=9.81, ULEN=1.0):
    """WAMIT ``.12d`` difference-frequency QTF (FOWT.readQTF, raft_fowt.py:2081-2128): rows
    ``T1 T2 head1 head2 dof |F| phase Re Im``, one triangle of the Hermitian matrix.
    -> qtf complex [nw1, nw2, nheads, 6] (dimensionalised by rho g ULEN, moments by a further ULEN, other triangle
    filled with the conjugate), w [nw1] rad/s ascending, heads [nheads] rad ascending.
    Only unidirectional tables (head1 == head2) with identical frequency columns, as the reference."""
    rows = np.loadtxt(path, ndmin=2)
    rows[:, 0:2] = 2.0 * np.pi / rows[:, 0:2]                       # periods -> rad/s (:2095)
    if not (rows[:, 2] == rows[:, 3]).all():
        raise ValueError("Only unidirectional QTFs are supported for now.")                      # :2099
    heads_deg = np.unique(rows[:, 2])
    w1, w2 = np.unique(rows[:, 0]), np.unique(rows[:, 1])
    if len(w1) != len(w2) or not (w1 == w2).all():
        raise ValueError("Both frequency columns in the input QTF must contain the same values.")  # :2110
    i1, i2 = np.searchsorted(w1, rows[:, 0]), np.searchsorted(w2, rows[:, 1])
    ih = np.searchsorted(heads_deg, rows[:, 2])
    idof = np.round(rows[:, 4] - 1).astype(int)
    factor = np.where(idof >= 3, rho * g * ULEN * ULEN, rho * g * ULEN)
    val = factor * (rows[:, 7] + 1j * rows[:, 8])
    qtf = np.zeros([len(w1), len(w2), len(heads_deg), 6], dtype=complex)
    for r in range(len(rows)):                                      # row order matters when a file repeats an entry
        qtf[i1[r], i2[r], ih[r], idof[r]] = val[r]
        if i1[r] != i2[r]:
            qtf[i2[r], i1[r], ih[r], idof[r]] = np.conj(val[r])      # :2127-2128
    return qtf, w1, heads_deg * 0.017453292519943295
