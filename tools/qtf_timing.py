#!/usr/bin/env python
"""Device timing of the second-order force kernel (k_qtf_force) at BASELINE config-3 size: nw = 2048, 256 sea states,
the shipped 56 x 56 x 1-heading QTF (and a synthetic 4-heading one).  CUDA events on torch's current stream."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raft_b200 import solver  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "cfg3q_OC4semi-QTF_nw96.npz"))
P = {k[2:]: z[k] for k in z.files if k.startswith("P_")}
nw, nC = 2048, 256
w = np.arange(1, nw + 1) * (2 * np.pi * 0.256 / nw)
Pb = dict(P, w=w, k=w ** 2 / 9.81, dw=w[1] - w[0])
for key in ("A_w", "B_w", "X_BEM", "bem_headings"):
    Pb.pop(key, None)
rng = np.random.default_rng(3)
cs = dict(Hs=rng.uniform(1, 10, nC), Tp=rng.uniform(5, 18, nC), gamma=np.zeros(nC), beta_deg=rng.uniform(-180, 180, nC),
          spec=np.zeros(nC, dtype=np.int32))
res = {}
for label, Pd in (("1head", Pb), ("4head", dict(Pb, qtf=np.stack([Pb["qtf"][:, :, 0, :] * s for s in (1.0, 0.7 + 0.2j, 1.3, -0.4 + 1j)], axis=2),
                                                qtf_heads=np.deg2rad([-90.0, 0.0, 45.0, 180.0])))):
    ses = solver.DeviceSession(solver.DesignBatch(Pd), solver.CaseTable(cs))
    for _ in range(3):
        ses.second_order_force()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    a.record()
    for _ in range(n):
        ses.second_order_force()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    pairs = nC * (nw * (nw + 1) // 2)
    res[label] = dict(ms=ms, pair_evals_per_s=pairs / (ms * 1e-3), cases_per_s=nC / (ms * 1e-3))
print(json.dumps(dict(workload="k_qtf_force nw=2048 nC=256 table 56x56x6", **res)))
