#!/usr/bin/env python
"""Development tool (see emu_general.cpp): run the staged generalised-DOF kernels under the host emulation on the
flex_VolturnUS-S-flexible fixture and compare with the CPU checker and the reference run.  ~30 minutes."""
import sys, os, subprocess, time, ctypes as C
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-o", os.path.join(HERE, "libgenemu.so"),
                       os.path.join(HERE, "emu_general.cpp")])
import numpy as np
from oracle import oracle as orc
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'flex_VolturnUS-S-flexible.npz'))
P = {k[2:]: z[k] for k in z.files if k.startswith("P_")}
n, nw, Ns = int(P["gen_nDOF"]), len(P["w"]), len(P["node_ls"])
mem = np.asarray(P["node_mem"], dtype=np.int64)
f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
frame = f8(np.concatenate([P["mem_q"][mem], P["mem_p1"][mem], P["mem_p2"][mem]], axis=1))
cd = f8(np.stack([P["node_a_q"]*P["node_Cd_q"], P["node_a_p1"]*P["node_Cd_p1"], P["node_a_p2"]*P["node_Cd_p2"], P["node_a_End"]*P["node_Cd_End"]], axis=1))
circ = np.ascontiguousarray(P["mem_circ"][mem], dtype=np.int32)
Iw = np.ascontiguousarray(P["node_Imat_w"], dtype=np.complex128) if "node_Imat_w" in P else None
cs = z["ref_run_solve_cases"]; nC = len(cs)
Hs, Tp, beta = f8(cs[:,0]), f8(cs[:,1]), f8(cs[:,2]); gam = np.zeros(nC); spec = np.zeros(nC, dtype=np.int32)
arrs = dict(w=f8(P["w"]), k=f8(P["k"]), node_r=f8(P["node_r"]), frame=frame, circ=circ, Imat=f8(P["node_Imat"]), a_i=f8(P["node_a_i"]), cd=cd,
            Tn=f8(P["gen_Tn"]), rr=f8(P["gen_rr"]), M=f8(z["gen_M"]), B=f8(z["gen_B"]), Cm=f8(z["gen_C"]))
Xi = np.zeros([nC, n, nw], dtype=np.complex128); st = np.zeros([nC,4], dtype=np.int32)
Fi = np.zeros([nC, n, nw], dtype=np.complex128); Bd = np.zeros([nC, n, n]); Fd = np.zeros([nC, n, nw], dtype=np.complex128)
lib = C.CDLL(os.path.join(HERE, 'libgenemu.so'))
p = lambda a: a.ctypes.data_as(C.c_void_p)
n_iter = int(z["n_iter"]); xi_start = float(z["xi_start"])
t = time.time()
lib.emu_general(C.c_int(n), C.c_int(nw), C.c_int(Ns), C.c_double(float(P["depth"])), C.c_double(float(P["rho"])), C.c_double(float(P["dw"])),
                p(arrs["w"]), p(arrs["k"]), p(arrs["node_r"]), p(frame), p(circ), p(arrs["Imat"]), p(Iw) if Iw is not None else None, p(arrs["a_i"]), p(cd),
                p(arrs["Tn"]), p(arrs["rr"]), p(arrs["M"]), p(arrs["B"]), p(arrs["Cm"]), C.c_int(nC), p(Hs), p(Tp), p(gam), p(beta), p(spec),
                C.c_int(n_iter), C.c_double(0.01), C.c_double(xi_start), p(Xi), p(st), p(Fi), p(Bd), p(Fd))
print('emu %.1f s' % (time.time()-t), 'status', st.tolist(), 'ref passes', z["ref_run_solve_passes"])
gd = orc.GeneralDesign(P)
for c in range(nC):
    _, Fo, u = orc.general_excitation(gd, 0, Hs[c], Tp[c], 0.0, beta[c])
    print(c, 'F_iner err', np.abs(Fi[c]-Fo).max()/np.abs(Fo).max())
    Bo, Fdo = orc.general_linearization(gd, u, np.full([n, nw], xi_start, dtype=complex))
    print(c, 'B_drag err', np.abs(Bd[c]-Bo).max()/np.abs(Bo).max(), 'F_drag err', np.abs(Fd[c]-Fdo).max()/np.abs(Fdo).max())
    print(c, 'Xi vs reference', np.abs(Xi[c]-z["ref_run_solve_Xi"][c]).max()/np.abs(z["ref_run_solve_Xi"][c]).max())
