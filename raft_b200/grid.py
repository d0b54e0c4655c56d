"""Frequency grid and dispersion relation of the model (host pre-pass, once per Model).

Restates raft_model.py:57 (grid) and helpers.py:377-392 (waveNumber).  The wave number is NOT the
exact dispersion root: the reference stops its fixed-point iteration at a 1e-3 relative change, and
parity requires reproducing that, so the same iteration is run here (vectorised over frequency,
each bin keeps iterating only while its own criterion is unmet, exactly like the scalar loop).
"""
import numpy as np


def make_w(min_freq, max_freq):
    """w [rad/s]: np.arange(min_freq, max_freq + 0.5*min_freq, min_freq) * 2 pi  (raft_model.py:57)."""
    return np.arange(min_freq, max_freq + 0.5 * min_freq, min_freq) * 2 * np.pi


_K_CACHE = {}


def wave_number(w, depth, e=0.001, g=9.81):
    """k(w) by the reference's fixed-point iteration k <- w^2 / (g tanh(k h))  (helpers.py:377-392).

    Long waves (k h << 1) need tens of thousands of iterations.  All bins iterate together while many are active; the last
    few stragglers (the lowest frequencies) finish in the reference's own scalar loop, which costs ~1 us per iteration instead
    of the ~7 us of a NumPy round over a one- or two-element index set (same operations, same ``np.tanh``: identical bits)."""
    w = np.atleast_1d(np.asarray(w, dtype=float))
    key = (w.tobytes(), float(depth), float(e), float(g))
    if key in _K_CACHE:                                   # a sweep builds every shard on the same grid
        return _K_CACHE[key].copy()
    k1 = w * w / g
    k2 = w * w / (np.tanh(k1 * depth) * g)
    idx = np.nonzero(np.abs(k2 - k1) / k1 > e)[0]
    while idx.size > 4:
        k1[idx] = k2[idx]
        k2[idx] = w[idx] * w[idx] / (np.tanh(k1[idx] * depth) * g)
        idx = idx[np.abs(k2[idx] - k1[idx]) / k1[idx] > e]
    tanh = np.tanh
    for i in idx:
        ww, a1, a2 = w[i] * w[i], k1[i], k2[i]
        while abs(a2 - a1) / a1 > e:
            a1 = a2
            a2 = ww / (tanh(a1 * depth) * g)
        k2[i] = a2
    if len(_K_CACHE) >= 8:
        _K_CACHE.pop(next(iter(_K_CACHE)))
    _K_CACHE[key] = k2.copy()
    return k2


def regrid(P, nw, max_freq):
    """Copy of a packed design on a new grid of ``nw`` bins up to ``max_freq`` Hz (min_freq = max_freq/nw).

    Only for designs without frequency tables (A_w/B_w/X_BEM/MCF) -- those must be re-interpolated
    by their producer."""
    for key in ("A_w", "B_w", "X_BEM", "node_in_p1_w"):
        if key in P and P[key] is not None:
            raise ValueError("regrid: design carries the frequency table %r" % key)
    Q = dict(P)
    w = make_w(max_freq / nw, max_freq)
    if len(w) != nw:
        raise ValueError("grid recipe produced %d bins, wanted %d" % (len(w), nw))
    Q["w"] = w
    Q["k"] = wave_number(w, float(P["depth"]))
    Q["dw"] = np.float64(w[1] - w[0])
    return Q
