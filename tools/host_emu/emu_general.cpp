// Development tool, NOT a product path and not used by tests, smoke() or bench.py: host emulation of the STAGED kernels in
// raft_b200/csrc/raftk_general.cuh, written when the round's GPU budget was spent.  The kernel source is compiled unchanged
// for the host: CUDA threads are std::threads, __syncthreads / warp shuffles are barriers, blocks run one after another.
// It answers one question -- is the kernel code as written logically right? -- and takes ~30 minutes for the fixture.
// Build + run: tools/host_emu/run_general.py (g++ -std=c++20 -pthread).
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>
#include <stdint.h>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define CUDART_PI 3.1415926535897931e+0
#define RAFTK_FLAG_NAN 1
#define RAFTK_SPEC_JONSWAP 0
#define RAFTK_SPEC_UNIT 1
#define RAFTK_SPEC_CONSTANT 2
static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim_;
static std::unique_ptr<std::barrier<>> g_block;
static std::vector<std::unique_ptr<std::barrier<>>> g_warp;
static double g_slots_d[1024];
static int g_slots_i[1024];
static inline void __syncthreads() { g_block->arrive_and_wait(); }
static inline double __shfl_xor_sync(unsigned, double v, int o)
{
    const int t = threadIdx.x, wp = t >> 5;
    g_slots_d[t] = v; g_warp[wp]->arrive_and_wait();
    const double r = g_slots_d[t ^ o]; g_warp[wp]->arrive_and_wait();
    return r;
}
static inline int __shfl_xor_sync(unsigned, int v, int o)
{
    const int t = threadIdx.x, wp = t >> 5;
    g_slots_i[t] = v; g_warp[wp]->arrive_and_wait();
    const int r = g_slots_i[t ^ o]; g_warp[wp]->arrive_and_wait();
    return r;
}
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
using std::isnan; using std::min; using std::max;

#include "../../raft_b200/csrc/raftk_common.cuh"
// copies of the (GPU-validated) sea-state helpers of raftk_tables.cuh
static inline double jonswap(double w, double Hs, double Tp, double Gamma)
{
    if (!(Gamma != 0.0)) { double t = Tp / sqrt(Hs); if (t <= 3.6) Gamma = 5.0; else if (t >= 5.0) Gamma = 1.0; else Gamma = exp(5.75 - 1.15 * t); }
    const double f = 0.5 / CUDART_PI * w;
    const double fpOvrf4 = pow(Tp * f, -4.0);
    const double C = 1.0 - (0.287 * log(Gamma));
    const double Sigma = (f <= 1.0 / Tp) ? 0.07 : 0.09;
    const double t = (f * Tp - 1.0) / Sigma;
    const double Alpha = exp(-0.5 * t * t);
    return 0.5 / CUDART_PI * C * 0.3125 * Hs * Hs * fpOvrf4 / f * exp(-1.25 * fpOvrf4) * pow(Gamma, Alpha);
}
static inline double sea_state_S(const CasesDev &Cs, int c, int i, int nw, double w, double dw)
{
    if (Cs.zeta_in) { const double z = Cs.zeta_in[(size_t)c * nw + i]; return z * z / (2.0 * dw); }
    const int spec = Cs.spec[c];
    if (spec == RAFTK_SPEC_JONSWAP) return jonswap(w, Cs.Hs[c], Cs.Tp[c], Cs.gamma[c]);
    if (spec == RAFTK_SPEC_UNIT) return 1.0;
    if (spec == RAFTK_SPEC_CONSTANT) return Cs.Hs[c];
    return 0.0;
}
static inline double sea_state_zeta(const CasesDev &Cs, int c, int i, int nw, double w, double dw)
{
    if (Cs.zeta_in) return Cs.zeta_in[(size_t)c * nw + i];
    return sqrt(2.0 * sea_state_S(Cs, c, i, nw, w, dw) * dw);
}
#include "../../raft_b200/csrc/raftk_general.cuh"

template <class F> static void launch(dim3 grid, unsigned threads, F &&body)
{
    blockDim_ = dim3(threads);
    g_block.reset(new std::barrier<>(threads));
    g_warp.clear();
    for (unsigned w = 0; w < (threads + 31) / 32; w++) g_warp.emplace_back(new std::barrier<>(std::min(32u, threads - 32 * w)));
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < threads; t++) ts.emplace_back([&, t]() { threadIdx = dim3(t); blockIdx = dim3(bx, by, bz); body(); });
        for (auto &th : ts) th.join();
    }
}

extern "C" int emu_general(int n, int nw, int Ns, double depth, double rho, double dw, const double *w, const double *k, const double *node_r,
                           const double *node_frame, const int *node_circ, const double *node_Imat, const double *node_Imat_w, const double *node_a_i,
                           const double *node_cd, const double *Tn, const double *rr, const double *M, const double *B, const double *C,
                           int nC, const double *Hs, const double *Tp, const double *gam, const double *beta, const int *spec,
                           int n_iter, double tol, double xi_start, double *Xi, int *status, double *F_iner_out, double *B_drag_out, double *F_drag_out)
{
    GenDev D; D.n = n; D.nw = nw; D.Ns = Ns; D.depth = depth; D.dw = dw; D.rho = rho; D.w = w; D.k = k; D.node_r = node_r; D.node_frame = node_frame;
    D.node_circ = node_circ; D.node_Imat = node_Imat; D.node_Imat_w = (const double2 *)node_Imat_w; D.node_a_i = node_a_i; D.node_cd = node_cd;
    D.Tn = Tn; D.rr = rr; D.M = M; D.B = B; D.C = C;
    CasesDev Cs; Cs.nC = nC; Cs.Hs = Hs; Cs.Tp = Tp; Cs.gamma = gam; Cs.beta_deg = beta; Cs.zeta_in = nullptr; Cs.spec = spec; Cs.primary = nullptr; Cs.F_2nd = nullptr;
    GenWork W;
    std::vector<double2> u((size_t)nC * Ns * 3 * nw), f6((size_t)nC * Ns * 6 * nw), Fi((size_t)nC * n * nw), Fd((size_t)nC * n * nw), XL((size_t)nC * n * nw);
    std::vector<double> Bm((size_t)nC * Ns * 9), Bd((size_t)nC * n * n);
    std::vector<double2> Z((size_t)nC * nw * n * (n + 1));
    std::vector<int> fl((size_t)nC * 4);
    W.u = u.data(); W.f6 = f6.data(); W.F_iner = Fi.data(); W.F_drag = Fd.data(); W.XiLast = XL.data(); W.Bmat = Bm.data(); W.B_drag = Bd.data(); W.Z = Z.data(); W.flags = fl.data();
    double2 *X = (double2 *)Xi;
    const unsigned fb = (nw + 127) / 128;
    launch(dim3(nC), 256, [&]() { k_gen_init(D, W, xi_start); });
    launch(dim3(fb, Ns, nC), 128, [&]() { k_gen_wave(D, Cs, W); });
    launch(dim3(fb, n, nC), 128, [&]() { k_gen_project(D, W, W.F_iner, 0); });
    memcpy(F_iner_out, Fi.data(), sizeof(double2) * Fi.size());
    for (int pass = 0; pass < n_iter + 1; pass++) {
        launch(dim3(Ns, nC), 128, [&]() { k_gen_node_pass(D, W); });
        launch(dim3(n, nC), 128, [&]() { k_gen_bdrag(D, W); });
        launch(dim3(fb, n, nC), 128, [&]() { k_gen_project(D, W, W.F_drag, 1); });
        if (pass == 0) { memcpy(B_drag_out, Bd.data(), sizeof(double) * Bd.size()); memcpy(F_drag_out, Fd.data(), sizeof(double2) * Fd.size()); }
        launch(dim3(nw, nC), 256, [&]() { k_gen_solve(D, W, X, tol); });
        launch(dim3(nC), 256, [&]() { k_gen_relax(D, W, X); });
        fprintf(stderr, "pass %d flags %d %d %d %d\n", pass, fl[0], fl[1], fl[2], fl[3]);
    }
    launch(dim3((nC + 127) / 128), 128, [&]() { k_gen_status(nC, W.flags, status); });
    return 0;
}
