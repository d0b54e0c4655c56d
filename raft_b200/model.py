"""Model: host-side mirror of the reference's ``raft.Model`` for ``solveDynamics`` / ``analyzeCases``.

Reference: raft_model.py:30-176 (construction), :264-433 (analyzeCases), :966-1302 (solveDynamics).
The per-frequency work runs on the GPU through the C ABI; statics / mooring / aero are out of scope and are
injected per FOWT (see ``raft_b200.fowt``).  Farms: every FOWT is linearised independently on the GPU, then the
coupled 6N system (block-diagonal impedances + an injected array-mooring stiffness) is solved per frequency by
``raftk_system_solve`` (raft_model.py:1164-1216).
"""
import numpy as np

from . import grid, packer, solver
from .fowt import FOWT


class Model:
    def __init__(self, design, matrices=None, array_stiffness=None, channels=None):
        """``channels``: optional turbine output channels per FOWT (``packer.pack_turbine_channels`` dicts: nacelle
        accelerations, tower-base moment) -- the turbine itself is outside this path, its constants enter here."""
        s = design.setdefault("settings", {})
        min_freq, max_freq = float(s.get("min_freq", 0.01)), float(s.get("max_freq", 1.00))
        self.XiStart = float(s.get("XiStart", 0.1))
        self.nIter = int(s.get("nIter", 15))
        self.w = grid.make_w(min_freq, max_freq)
        self.nw = len(self.w)
        self.depth = float(design["site"]["water_depth"])
        self.k = grid.wave_number(self.w, self.depth)
        self.design = design
        self.fowtList, self.coords = [], []
        if "array" in design:
            keys, rows = design["array"]["keys"], design["array"]["data"]
            mats = matrices if isinstance(matrices, (list, tuple)) else [matrices] * len(rows)
            for i, row in enumerate(rows):
                info = dict(zip(keys, row))
                plats = design["platforms"] if "platforms" in design else [design["platform"]]     # raft_model.py:86-101
                d_i = dict(site=design["site"], platform=plats[int(info["platformID"]) - 1])
                self.fowtList.append(FOWT(d_i, self.w, depth=self.depth, x_ref=info["x_location"], y_ref=info["y_location"],
                                          heading_adjust=info.get("heading_adjust", 0), matrices=mats[i], k=self.k))
                self.coords.append([info["x_location"], info["y_location"]])
        else:
            self.fowtList.append(FOWT(design, self.w, depth=self.depth, matrices=matrices, k=self.k))
            self.coords.append([0.0, 0.0])
        self.nFOWT = len(self.fowtList)
        self.nDOF = 6 * self.nFOWT
        self.C_array = None if array_stiffness is None else np.array(array_stiffness, dtype=float)   # stands in for ms.getCoupledStiffnessA
        self.results = {}
        self.channels = list(channels) if isinstance(channels, (list, tuple)) else [channels] * self.nFOWT
        for f in self.fowtList:
            f.calcHydroConstants()

    # raft_model.py:966-1302 -------------------------------------------------------------------------------
    def solveDynamics(self, case, tol=0.01, conv_plot=0, RAO_plot=0, display=0):
        """Response amplitudes for one load case -> self.Xi [nWaves+1, nDOF, nw] (last row zero, as :1195)."""
        out = self._solve_batch([case], tol)
        trains = out["Xi_trains"][0]                                      # [nWaves, nDOF, nw]
        Xi = np.zeros([len(trains) + 1, self.nDOF, self.nw], dtype=complex)
        Xi[:-1] = trains
        self.Xi = Xi
        for i, f in enumerate(self.fowtList):
            f.Xi = Xi[:, 6 * i:6 * i + 6, :]
            f.Xi_fullDOF = f.Xi
        self.results["response"] = {}
        return self.Xi

    # raft_model.py:264-433 (dynamics part) ---------------------------------------------------------------------
    def analyzeCases(self, display=0, meshDir=None, RAO_plot=False, cases=None, tol=0.01):
        """All load cases in ONE batched GPU call.  Positional arguments as the reference's
        ``analyzeCases(display=0, meshDir=..., RAO_plot=False)`` (raft_model.py:264; meshDir / RAO_plot concern the BEM mesh
        and plotting, outside this path and ignored); ``cases=``: list of case dicts (default: the design's table).
        Fills results['freq_rad'], results['Xi'] [nCases, nDOF, nw], results['status'] [nCases, nFOWT, 4]."""
        if cases is None:
            keys = self.design["cases"]["keys"]
            cases = [dict(zip(keys, row)) for row in self.design["cases"]["data"]]
        out = self._solve_batch(cases, tol)
        self.results["freq_rad"] = self.w
        self.results["Xi"] = out["Xi"]
        self.results["Xi_trains"] = out["Xi_trains"]
        self.results["status"] = out["status"]
        # response statistics per case and FOWT (raft_fowt.py:2299-2353; zero mean offsets: statics are out of scope).
        # getRMS / getPSD sum the squares over a case's wave trains (helpers.py:678-700), so the per-train device
        # reductions are combined here: std = sqrt(sum std_t^2), PSD = sum PSD_t.
        nC = len(cases)
        owner = out["owner"]
        Xi_units = out["Xi_all"].reshape(len(owner), self.nFOWT, 6, self.nw)                  # [nTrains, nFOWT, 6, nw]
        sd_t, psd_t = solver.response_stats(Xi_units, self.w[1] - self.w[0])
        names = ("surge", "sway", "heave", "roll", "pitch", "yaw")
        ch_stats = [None if ch is None else solver.channel_stats(ch["coef"], Xi_units[:, i], self.w[1] - self.w[0])
                    for i, ch in enumerate(self.channels)]                                     # (std [nT,nch], PSD [nT,nch,nw], -)
        self.results["case_metrics"] = {}
        for ic in range(nC):
            idx = np.nonzero(owner == ic)[0]
            sd = np.sqrt((sd_t[idx] ** 2).sum(axis=0))
            psd = psd_t[idx].sum(axis=0)
            self.results["case_metrics"][ic] = {}
            for i in range(self.nFOWT):
                m = {}
                for k_, nm in enumerate(names):
                    m[nm + "_avg"], m[nm + "_std"] = 0.0, sd[i, k_]
                    m[nm + "_max"], m[nm + "_min"] = 3 * sd[i, k_], -3 * sd[i, k_]
                    m[nm + "_PSD"] = psd[i, k_]
                    ra = np.zeros([len(idx) + 1, self.nw], dtype=complex)                     # all trains + the zero row (:1195)
                    ra[:-1] = Xi_units[idx, i, k_] * (57.29577951308232 if k_ >= 3 else 1.0)
                    m[nm + "_RA"] = ra
                if ch_stats[i] is not None:                                                   # raft_fowt.py:2401-2444, 2504-2538
                    ch = self.channels[i]
                    nrot = 1 + max(ir for _, ir in ch["names"])
                    sd_c = np.sqrt((ch_stats[i][0][idx] ** 2).sum(axis=0))
                    psd_c = ch_stats[i][1][idx].sum(axis=0)
                    for k_, (nm, ir) in enumerate(ch["names"]):
                        for suffix in ("_avg", "_std", "_max", "_min"):
                            m.setdefault(nm + suffix, np.zeros(nrot))
                        m.setdefault(nm + "_PSD", np.zeros([self.nw, nrot]))
                        m[nm + "_avg"][ir], m[nm + "_std"][ir] = ch["avg"][k_], sd_c[k_]
                        m[nm + "_max"][ir], m[nm + "_min"][ir] = ch["avg"][k_] + 3 * sd_c[k_], ch["avg"][k_] - 3 * sd_c[k_]
                        m[nm + "_PSD"][:, ir] = psd_c[k_]
                self.results["case_metrics"][ic][i] = m
        return self.results

    def _solve_batch(self, cases, tol):
        table, owner, first = packer.pack_case_trains(cases)
        ct = solver.CaseTable(table)
        nC = len(cases)
        packs = [f.pack() for f in self.fowtList]
        batch = solver.DesignBatch([{k: v for k, v in P.items() if not k.startswith("qs_")} for P in packs])
        want = ("Xi", "status", "B_drag", "F_drag", "F_iner", "F_BEM", "zeta")
        sec = [int(getattr(f, "potSecOrder", 0)) for f in self.fowtList]
        if any(s_ == 1 for s_ in sec):                                      # slender-body QTF inside the loop (raft_model.py:1106-1131)
            if not all(s_ == 1 for s_ in sec):
                raise NotImplementedError("mixing potSecOrder 1 with other settings in one array is not supported")
            o = solver.solve_dynamics_slender(packs, ct, n_iter=self.nIter, tol=tol, xi_start=self.XiStart, want=want)
            for i, f in enumerate(self.fowtList):
                f.qtf = np.ascontiguousarray(o["qtf"][i, -1][:, :, None, :])
                f.heads_2nd = [float(ct.arrays["beta_deg"][-1]) * 0.017453292519943295]
        else:
            if batch.n_qtf_w:                                               # potSecOrder 2 (raft_model.py:1035-1038)
                want += ("F_2nd", "F_2nd_mean")
            if self.nFOWT > 1 and self.C_array is not None:
                # coupled array: per-FOWT linearisation + block assembly + 6N x 6N solve in one device call (raft_model.py:1164-1216)
                o = solver.solve_dynamics_farm(batch, ct, C_arr=self.C_array, n_iter=self.nIter, tol=tol, xi_start=self.XiStart, want=want)
                if np.any(o["info"]):
                    raise np.linalg.LinAlgError("Singular matrix")          # np.linalg.inv at raft_model.py:1191
            else:
                o = solver.solve_dynamics(batch, ct, n_iter=self.nIter, tol=tol, xi_start=self.XiStart, want=want)
        if "primary" in table:                                              # secondary trains share their primary's B_drag
            o["B_drag"] = o["B_drag"][:, table["primary"]]
        st = o["status"][:, first]                                          # [nFOWT, nC, 4] (train 0 of every case)
        solver.raise_on_flags(st)                                           # raft_model.py:1089 (LinAlgError), :1098-1099 (NaN)
        w = self.w
        for i, f in enumerate(self.fowtList):
            P = f.pack()
            f.B_hydro_drag, f.F_hydro_drag = o["B_drag"][i, -1], o["F_drag"][i, -1]
            f.zeta, f.F_BEM, f.F_hydro_iner = o["zeta"][-1:], o["F_BEM"][i, -1:], o["F_iner"][i, -1:]
            last = np.nonzero(owner == nC - 1)[0]                           # the trains of the last case
            f.Fhydro_2nd = np.zeros([len(last), 6, self.nw], dtype=complex)
            f.Fhydro_2nd_mean = np.zeros([len(last), 6])
            if "F_2nd" in o:
                f.Fhydro_2nd[:] = o["F_2nd"][i, last]
                f.Fhydro_2nd_mean[:] = o["F_2nd_mean"][i, last]
            M = P["M0"][:, :, None] + (P["A_w"] if "A_w" in P else 0.0)
            B = (P["B0"] + f.B_hydro_drag)[:, :, None] + (P["B_w"] if "B_w" in P else 0.0)
            f.Z = -w ** 2 * M + 1j * w * B + P["C0"][:, :, None]             # raft_model.py:1086, 1155 (last case)
        nT = ct.n_cases
        Xi_all = np.moveaxis(o["Xi"], 0, 1).reshape(nT, self.nDOF, self.nw)  # [nTrains, 6N, nw]
        if "Xi_sys" in o:
            Xi_all = o["Xi_sys"]                                            # coupled system response, computed on the device
        elif self.nFOWT > 1 and self.C_array is not None:
            # (slender-body QTF path) coupled system: Z_sys = blockdiag(Z_i) + C_array; F = Z_i Xi_i  (raft_model.py:1164-1216)
            Xi_all = self._couple(o, nT)
        Xi_trains = [Xi_all[owner == ic] for ic in range(nC)]
        return dict(Xi=Xi_all[first], Xi_trains=Xi_trains, status=np.moveaxis(st, 0, 1), Xi_all=Xi_all, owner=owner)

    def _couple(self, o, nC):
        n, nw, w = self.nDOF, self.nw, self.w
        Xi = np.zeros([nC, n, nw], dtype=complex)
        packs = [f.pack() for f in self.fowtList]
        for c in range(nC):
            Z = np.zeros([nw, n, n], dtype=complex)
            F = np.zeros([nw, n], dtype=complex)
            for i, P in enumerate(packs):
                M = P["M0"][:, :, None] + (P["A_w"] if "A_w" in P else 0.0)
                B = (P["B0"] + o["B_drag"][i, c])[:, :, None] + (P["B_w"] if "B_w" in P else 0.0)
                Zi = np.moveaxis(-w ** 2 * M + 1j * w * B + P["C0"][:, :, None], 2, 0)     # [nw,6,6]
                Z[:, 6 * i:6 * i + 6, 6 * i:6 * i + 6] = Zi
                Fi = o["F_BEM"][i, c] + o["F_iner"][i, c] + o["F_drag"][i, c]
                if "F_2nd" in o:
                    Fi = Fi + o["F_2nd"][i, c]                              # raft_model.py:1212
                F[:, 6 * i:6 * i + 6] = np.moveaxis(Fi, 0, 1)
            Z += self.C_array[None, :, :]
            X, info = solver.system_solve(Z, F)
            if np.any(info):
                raise np.linalg.LinAlgError("singular system impedance matrix")
            Xi[c] = X.T
        return Xi
