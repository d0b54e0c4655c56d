// raftk.cu -- sm_100a kernels + C ABI for the RAO-solve hot path (see include/raftk.h).
//
// Kernels
//   k_depth_table   : depth-decay functions cosh/sinh ratios per (node, frequency)      helpers.py:207-222
//   k_excitation    : sea state -> zeta, node phase table, strip inertial + BEM excitation
//                     raft_fowt.py:1732-1888, raft_member.py:1899-1992, helpers.py:188-236,703-760
//   k_drag_solve    : per (design, case) CTA cluster: drag linearisation (cross-frequency RMS),
//                     B_drag, F_drag, impedance assembly, 6x6 complex LU per frequency, convergence,
//                     relaxation          raft_model.py:1052-1142, raft_fowt.py:1891-1957,
//                                         raft_member.py:1995-2152, helpers.py:149-184,678-684
//   k_system_solve  : dense n x n complex solve per frequency (farm)      raft_model.py:1164-1216
//   k_fp64_peak     : DFMA micro-benchmark for the FP64 roofline denominator
//
// Algebra used by the kernels: with member frame (q,p1,p2), node position
// r_j = rA + ls_j q and lever arm a = rA - r_ref, a 3-vector d in {q,p1,p2} at node j acts on the
// 6-DOF body through V_jd = [d ; r_j x d]:
//     V_jq = [q ; a x q],  V_jp1 = [p1 ; a x p1 + ls_j p2],  V_jp2 = [p2 ; a x p2 - ls_j p1]
// (q x p1 = p2, q x p2 = -p1).  Wave velocity projections are c_jd(w) = zeta w E_j (C_j h_d + i S_j d_z)
// with E_j = exp(-i k (x_j cos b + y_j sin b)), h_d = d_x cos b + d_y sin b, and (C_j,S_j) the depth
// functions.  Everything the reference does per (node, frequency) with 3x3 / 6x6 matrices reduces to
// complex scalars per node and a handful of sums per member.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <algorithm>
#include <type_traits>
#include <vector>

#include "../../include/raftk.h"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static long long g_launches = 0;

static int set_err(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return set_err(RAFTK_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// ---- optional per-kernel event timing (roofline report) ----------------------------------------------
struct ProfRec { cudaEvent_t a, b; int kind; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

static void prof_begin_call()
{
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
}
struct ProfScope {
    cudaStream_t st; int idx = -1;
    ProfScope(cudaStream_t s, int kind) : st(s)
    {
        if (!g_prof_on) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        ProfRec r; r.kind = kind;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, st);
        g_prof.push_back(r); idx = (int)g_prof.size() - 1;
    }
    ~ProfScope()
    {
        if (idx < 0) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        cudaEventRecord(g_prof[idx].b, st);
    }
};
extern "C" void raftk_profile_enable(int on) { g_prof_on = on != 0; }
extern "C" int raftk_profile_read(double ms[3], int launches[3])
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int t = 0; t < 3; t++) { ms[t] = 0.0; launches[t] = 0; }
    for (auto &r : g_prof) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) return RAFTK_ECUDA;
        float f = 0.f;
        if (cudaEventElapsedTime(&f, r.a, r.b) != cudaSuccess) return RAFTK_ECUDA;
        ms[r.kind] += f; launches[r.kind]++;
    }
    return RAFTK_OK;
}

extern "C" int raftk_version(void) { return RAFTK_VERSION; }
extern "C" const char *raftk_last_error(void) { return g_err; }
extern "C" long long raftk_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#define SOLVE_THREADS 128
#define CHUNK_NODES 10          // nodes per register-accumulator chunk in the RMS pass (3*10 <= 32)
#define MEM_STRIDE 24           // doubles per member in shared memory

struct DesignsDev {
    int nD, nw, max_nodes, max_members, n_bem_head;
    double depth, rho, g, dw;
    const double *w, *k;
    const int *member_offset;
    const double *mem_frame, *mem_rA, *mem_arm;
    const int *mem_node_start, *mem_circ;
    const double *node_ls, *node_cd_q, *node_cd_p1, *node_cd_p2, *node_in_q, *node_in_p1, *node_in_p2, *node_pa;
    const double2 *node_in_p1_w, *node_in_p2_w;
    const double *M0, *B0, *C0, *A_w, *B_w;
    const double *bem_headings, *X_BEM, *bem_xyh;
};

struct CasesDev {
    int nC;
    const double *Hs, *Tp, *gamma, *beta_deg, *zeta_in;
    const int *spec;
    const int *primary;     // [nC] or NULL: case whose drag linearisation this case reuses (secondary wave trains)
};

struct Work {          // workspace views for one chunk of designs [d0, d0+nDc)
    int d0, nDc;
    double2 *depth_tab;   // [nDc][max_nodes][nw]           (C, S)
    double2 *phase_tab;   // [nDc][nC][max_nodes][nw]       zeta*w*E
    double2 *F0;          // [nDc][nC][6][nw]               F_BEM + F_iner
    double *zeta;         // [nC][nw]
};

// depth functions of helpers.py:207-222 (k == 0 / k h > 89.4 / general)
__device__ __forceinline__ void depth_funcs(double k, double h, double z, double &S_, double &C_, double &P_)
{
    if (k == 0.0) { S_ = 1.0; C_ = 99999.0; P_ = 99999.0; }
    else if (k * h > 89.4) {
        double e = exp(k * z);
        S_ = e; C_ = e; P_ = e + exp(-k * (z + 2.0 * h));
    } else {
        double sh = sinh(k * h);
        S_ = sinh(k * (z + h)) / sh;
        C_ = cosh(k * (z + h)) / sh;
        P_ = cosh(k * (z + h)) / cosh(k * h);
    }
}

// ------------------------------------------------------------------------------------------------
// K0: depth table.  grid (ceil(nw/128), members of the chunk), block 128
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_depth_table(DesignsDev D, Work W)
{
    // blockIdx.y = local design, blockIdx.z unused; loop over the design's members and nodes
    const int dl = blockIdx.y, d = W.d0 + dl;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.nw) return;
    const int m0 = D.member_offset[d], m1 = D.member_offset[d + 1];
    const int nbase = D.mem_node_start[m0];
    const double k = D.k[i], h = D.depth;
    for (int m = m0; m < m1; m++) {
        const double qz = D.mem_frame[9 * m + 2], zA = D.mem_rA[3 * m + 2];
        const int j0 = D.mem_node_start[m], j1 = D.mem_node_start[m + 1];
        for (int j = j0; j < j1; j++) {
            double z = zA + D.node_ls[j] * qz;
            double S_, C_, P_;
            depth_funcs(k, h, z, S_, C_, P_);
            W.depth_tab[((size_t)dl * D.max_nodes + (j - nbase)) * D.nw + i] = make_double2(C_, S_);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1: excitation.  grid (ceil(nw/128), nC, nDc), block 128, thread = frequency
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double jonswap(double w, double Hs, double Tp, double Gamma)
{
    // helpers.py:733-760
    if (!(Gamma != 0.0)) {
        double t = Tp / sqrt(Hs);
        if (t <= 3.6) Gamma = 5.0;
        else if (t >= 5.0) Gamma = 1.0;
        else Gamma = exp(5.75 - 1.15 * t);
    }
    const double f = 0.5 / CUDART_PI * w;
    const double fpOvrf4 = pow(Tp * f, -4.0);
    const double C = 1.0 - (0.287 * log(Gamma));
    const double Sigma = (f <= 1.0 / Tp) ? 0.07 : 0.09;
    const double t = (f * Tp - 1.0) / Sigma;
    const double Alpha = exp(-0.5 * t * t);
    return 0.5 / CUDART_PI * C * 0.3125 * Hs * Hs * fpOvrf4 / f * exp(-1.25 * fpOvrf4) * pow(Gamma, Alpha);
}

// wave amplitude of one case at one frequency: explicit table or spectrum -> zeta = sqrt(2 S dw) (raft_fowt.py:1759-1774)
__device__ __forceinline__ double sea_state_zeta(const CasesDev &Cs, int c, int i, int nw, double w, double dw)
{
    if (Cs.zeta_in) return Cs.zeta_in[(size_t)c * nw + i];
    const int spec = Cs.spec[c];
    double S;
    if (spec == RAFTK_SPEC_JONSWAP) S = jonswap(w, Cs.Hs[c], Cs.Tp[c], Cs.gamma[c]);
    else if (spec == RAFTK_SPEC_UNIT) S = 1.0;
    else if (spec == RAFTK_SPEC_CONSTANT) S = Cs.Hs[c];
    else S = 0.0;
    return sqrt(2.0 * S * dw);
}

// BEM excitation of design d at frequency i for heading beta: bracket the heading in the (heading-relative)
// coefficient table with wrap-around, interpolate, rotate back to the global frame, scale by the wave amplitude
// and the array phase offset (raft_fowt.py:1796-1849).  Br/Bi receive the 6 complex force components.
__device__ __forceinline__ void bem_excitation(const DesignsDev &D, int d, int i, double k, double beta, double sb, double cb,
                                               double zeta, double (&Br)[6], double (&Bi)[6])
{
    const int nhs = D.n_bem_head, nw = D.nw;
    const double *hd = D.bem_headings;
    const double xr = D.bem_xyh[3 * d], yr = D.bem_xyh[3 * d + 1], hadj = D.bem_xyh[3 * d + 2];
    double bdeg = fmod(beta * (180.0 / CUDART_PI) - hadj, 360.0);
    if (bdeg < 0) bdeg += 360.0;                                   // python's % is non-negative
    int i1 = 0, i2 = 0; double f2 = 0;
    if (bdeg <= hd[0]) {
        const double hlast = hd[nhs - 1] - 360.0;
        i1 = nhs - 1; i2 = 0; f2 = (bdeg - hlast) / (hd[0] - hlast);
    } else if (bdeg >= hd[nhs - 1]) {
        const double hfirst = hd[0] + 360.0;
        i1 = nhs - 1; i2 = 0; f2 = (bdeg - hd[nhs - 1]) / (hfirst - hd[nhs - 1]);
    } else {
        for (int t = 0; t < nhs - 1; t++) if (hd[t + 1] > bdeg) { i1 = t; i2 = t + 1; f2 = (bdeg - hd[t]) / (hd[t + 1] - hd[t]); break; }
    }
    const double f1 = 1.0 - f2;
    const double2 *X = reinterpret_cast<const double2 *>(D.X_BEM) + (size_t)d * nhs * 6 * nw;
    double Xr[6], Xi_[6];
#pragma unroll
    for (int a = 0; a < 6; a++) {
        const double2 x1 = X[((size_t)i1 * 6 + a) * nw + i], x2 = X[((size_t)i2 * 6 + a) * nw + i];
        Xr[a] = x1.x * f1 + x2.x * f2; Xi_[a] = x1.y * f1 + x2.y * f2;
    }
    double Rr[6], Ri[6];
    Rr[0] = Xr[0] * cb - Xr[1] * sb; Ri[0] = Xi_[0] * cb - Xi_[1] * sb;
    Rr[1] = Xr[0] * sb + Xr[1] * cb; Ri[1] = Xi_[0] * sb + Xi_[1] * cb;
    Rr[2] = Xr[2];                   Ri[2] = Xi_[2];
    Rr[3] = Xr[3] * cb - Xr[4] * sb; Ri[3] = Xi_[3] * cb - Xi_[4] * sb;
    Rr[4] = Xr[3] * sb + Xr[4] * cb; Ri[4] = Xi_[3] * sb + Xi_[4] * cb;
    Rr[5] = Xr[5];                   Ri[5] = Xi_[5];
    double sp, cp;
    sincos(-(k * (xr * cb + yr * sb)), &sp, &cp);
    const double pr = zeta * cp, pi = zeta * sp;
#pragma unroll
    for (int a = 0; a < 6; a++) { Br[a] = Rr[a] * pr - Ri[a] * pi; Bi[a] = Rr[a] * pi + Ri[a] * pr; }
}

struct ExcOut { double2 *F_iner, *F_BEM; double *zeta; };

__global__ void __launch_bounds__(128) k_excitation(DesignsDev D, CasesDev Cs, Work W, ExcOut O)
{
    const int c = blockIdx.y, dl = blockIdx.z, d = W.d0 + dl;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.nw) return;
    const int nw = D.nw;
    const double w = D.w[i], k = D.k[i];

    const double zeta = sea_state_zeta(Cs, c, i, nw, w, D.dw);
    if (dl == 0) {
        W.zeta[(size_t)c * nw + i] = zeta;
        if (O.zeta && W.d0 == 0) O.zeta[(size_t)c * nw + i] = zeta;
    }
    const double beta = Cs.beta_deg[c] * (CUDART_PI / 180.0);   // np.deg2rad
    double sb, cb;
    sincos(beta, &sb, &cb);
    const double zw = zeta * w;

    const int m0 = D.member_offset[d], m1 = D.member_offset[d + 1];
    const int nbase = D.mem_node_start[m0];
    const size_t unit = (size_t)dl * Cs.nC + c;
    double2 *ptab = W.phase_tab + unit * D.max_nodes * nw;
    const double2 *dtab = W.depth_tab + (size_t)dl * D.max_nodes * nw;

    double Fr[6] = {0, 0, 0, 0, 0, 0}, Fi[6] = {0, 0, 0, 0, 0, 0};
    for (int m = m0; m < m1; m++) {
        const double *fr = D.mem_frame + 9 * m;
        const double q0 = fr[0], q1 = fr[1], q2 = fr[2], p10 = fr[3], p11 = fr[4], p12 = fr[5], p20 = fr[6], p21 = fr[7], p22 = fr[8];
        const double xA = D.mem_rA[3 * m], yA = D.mem_rA[3 * m + 1], zA = D.mem_rA[3 * m + 2];
        const double hq = q0 * cb + q1 * sb, h1 = p10 * cb + p11 * sb, h2 = p20 * cb + p21 * sb;
        const int j0 = D.mem_node_start[m], j1 = D.mem_node_start[m + 1];
        double Aqr = 0, Aqi = 0, A1r = 0, A1i = 0, A2r = 0, A2i = 0, L1r = 0, L1i = 0, L2r = 0, L2i = 0;
        for (int j = j0; j < j1; j++) {
            const double ls = D.node_ls[j];
            const double x = xA + ls * q0, y = yA + ls * q1;
            double se, ce;
            sincos(-(k * (cb * x + sb * y)), &se, &ce);          // E = exp(-i k (x cos b + y sin b))
            const double er = zw * ce, ei = zw * se;             // zeta*w*E
            ptab[(size_t)(j - nbase) * nw + i] = make_double2(er, ei);
            const double2 cs = dtab[(size_t)(j - nbase) * nw + i];
            const double inq = D.node_in_q[j], pa = D.node_pa[j];
            double in1 = D.node_in_p1[j], in2 = D.node_in_p2[j], in1i = 0.0, in2i = 0.0;
            if (D.node_in_p1_w) {                                 // MacCamy-Fuchs: complex, per frequency
                const double2 v1 = D.node_in_p1_w[(size_t)j * nw + i], v2 = D.node_in_p2_w[(size_t)j * nw + i];
                in1 = v1.x; in1i = v1.y; in2 = v2.x; in2i = v2.y;
            }
            if (inq != 0.0 || in1 != 0.0 || in2 != 0.0 || in1i != 0.0 || in2i != 0.0 || pa != 0.0) {
                // c_d = zeta w E (C h_d + i S d_z); inertial force coefficient along d: i w in_d c_d
                double gr, gi, cr, ci;
                gr = cs.x * hq; gi = cs.y * q2; cr = er * gr - ei * gi; ci = er * gi + ei * gr;
                double fqr = -w * inq * ci, fqi = w * inq * cr;
                gr = cs.x * h1; gi = cs.y * p12; cr = er * gr - ei * gi; ci = er * gi + ei * gr;
                const double f1r = -w * (in1 * ci + in1i * cr), f1i = w * (in1 * cr - in1i * ci);
                gr = cs.x * h2; gi = cs.y * p22; cr = er * gr - ei * gi; ci = er * gi + ei * gr;
                const double f2r = -w * (in2 * ci + in2i * cr), f2i = w * (in2 * cr - in2i * ci);
                if (pa != 0.0) {                                  // dynamic pressure on end area (member:1988)
                    double S_, C_, P_;
                    depth_funcs(k, D.depth, zA + ls * q2, S_, C_, P_);
                    fqr += pa * P_ * zeta * ce; fqi += pa * P_ * zeta * se;
                }
                Aqr += fqr; Aqi += fqi; A1r += f1r; A1i += f1i; A2r += f2r; A2i += f2i;
                L1r += ls * f1r; L1i += ls * f1i; L2r += ls * f2r; L2i += ls * f2i;
            }
        }
        const double *arm = D.mem_arm + 3 * m;
        const double a0 = arm[0], a1 = arm[1], a2 = arm[2];
        // a x q, a x p1, a x p2
        const double aq0 = a1 * q2 - a2 * q1, aq1 = a2 * q0 - a0 * q2, aq2 = a0 * q1 - a1 * q0;
        const double b10 = a1 * p12 - a2 * p11, b11 = a2 * p10 - a0 * p12, b12 = a0 * p11 - a1 * p10;
        const double b20 = a1 * p22 - a2 * p21, b21 = a2 * p20 - a0 * p22, b22 = a0 * p21 - a1 * p20;
        Fr[0] += q0 * Aqr + p10 * A1r + p20 * A2r;  Fi[0] += q0 * Aqi + p10 * A1i + p20 * A2i;
        Fr[1] += q1 * Aqr + p11 * A1r + p21 * A2r;  Fi[1] += q1 * Aqi + p11 * A1i + p21 * A2i;
        Fr[2] += q2 * Aqr + p12 * A1r + p22 * A2r;  Fi[2] += q2 * Aqi + p12 * A1i + p22 * A2i;
        Fr[3] += aq0 * Aqr + b10 * A1r + b20 * A2r + p20 * L1r - p10 * L2r;
        Fi[3] += aq0 * Aqi + b10 * A1i + b20 * A2i + p20 * L1i - p10 * L2i;
        Fr[4] += aq1 * Aqr + b11 * A1r + b21 * A2r + p21 * L1r - p11 * L2r;
        Fi[4] += aq1 * Aqi + b11 * A1i + b21 * A2i + p21 * L1i - p11 * L2i;
        Fr[5] += aq2 * Aqr + b12 * A1r + b22 * A2r + p22 * L1r - p12 * L2r;
        Fi[5] += aq2 * Aqi + b12 * A1i + b22 * A2i + p22 * L1i - p12 * L2i;
    }
    const size_t ogl = ((size_t)d * Cs.nC + c) * 6 * nw;     // global output index base
    if (O.F_iner)
        for (int a = 0; a < 6; a++) O.F_iner[ogl + (size_t)a * nw + i] = make_double2(Fr[a], Fi[a]);

    double Br[6] = {0, 0, 0, 0, 0, 0}, Bi[6] = {0, 0, 0, 0, 0, 0};
    if (D.n_bem_head > 0) bem_excitation(D, d, i, k, beta, sb, cb, zeta, Br, Bi);
    if (O.F_BEM)
        for (int a = 0; a < 6; a++) O.F_BEM[ogl + (size_t)a * nw + i] = make_double2(Br[a], Bi[a]);
    double2 *F0 = W.F0 + unit * 6 * nw;
    for (int a = 0; a < 6; a++) F0[(size_t)a * nw + i] = make_double2(Br[a] + Fr[a], Bi[a] + Fi[a]);
}

// ------------------------------------------------------------------------------------------------
// K2: drag linearisation + impedance solve
// ------------------------------------------------------------------------------------------------
struct SolveParams {
    int n_iter, CS, nwl, mode;          // mode 0: solve loop; 1: single linearisation pass with Xi_in
    double tol, xi_start;
    const double2 *Xi_in;               // [nD][nC][6][nw] (mode 1)
    double2 *Xi_out, *Fdrag_out;        // [nD][nC][6][nw]
    double *Bdrag_out;                  // [nD][nC][36]
    int *status;                        // [nD][nC][4]
};

// sum of 32 per-lane value arrays across the warp: on return lane l holds the warp total of v[l].
// Fixed butterfly order -> deterministic.  (V-1 shuffles instead of 5V.)
__device__ __forceinline__ double warp_multi_reduce32(double (&v)[32])
{
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int t = 0; t < half; t++) {
            const double keep = up ? v[t + half] : v[t];
            const double send = up ? v[t] : v[t + half];
            v[t] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// compile-time loop: indices are constants, so register arrays never fall back to local memory
// (ptxas/NVVM give up on "#pragma unroll" for the triple LU nest and then index dynamically).
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// 6x6 complex solve in registers: LU with partial pivoting (|re|+|im| metric, as LAPACK izamax),
// forward elimination applied to b on the fly, back substitution.  Returns false on a zero pivot.
__device__ __forceinline__ bool solve6(double (&ar)[6][6], double (&ai)[6][6], double (&br)[6], double (&bi)[6])
{
    double rr[6], ri[6];
    bool ok = true;
    static_for<0, 6>([&](auto K) {
        constexpr int k = decltype(K)::value;
        int p = k;
        double best = fabs(ar[k][k]) + fabs(ai[k][k]);
        static_for<k + 1, 6>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const double t = fabs(ar[i][k]) + fabs(ai[i][k]);
            if (t > best) { best = t; p = i; }
        });
        if (best == 0.0) ok = false;
        // (measured: guarding the swaps with a warp vote "does any lane pivot here?" is 4 % slower than always selecting)
        static_for<k + 1, 6>([&](auto I) {
            constexpr int i = decltype(I)::value;
            // row swap as register selects (a dynamic row index would push the matrix to local memory)
            const bool sw = (p == i);
            static_for<k, 6>([&](auto J) {
                constexpr int j = decltype(J)::value;
                const double r1 = ar[k][j], r2 = ar[i][j], i1 = ai[k][j], i2 = ai[i][j];
                ar[k][j] = sw ? r2 : r1; ar[i][j] = sw ? r1 : r2;
                ai[k][j] = sw ? i2 : i1; ai[i][j] = sw ? i1 : i2;
            });
            const double r1 = br[k], r2 = br[i], i1 = bi[k], i2 = bi[i];
            br[k] = sw ? r2 : r1; br[i] = sw ? r1 : r2;
            bi[k] = sw ? i2 : i1; bi[i] = sw ? i1 : i2;
        });
        const double pr = ar[k][k], pi = ai[k][k];
        const double inv = 1.0 / (pr * pr + pi * pi);
        rr[k] = pr * inv; ri[k] = -pi * inv;
        static_for<k + 1, 6>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const double lr = ar[i][k] * rr[k] - ai[i][k] * ri[k];
            const double li = ar[i][k] * ri[k] + ai[i][k] * rr[k];
            static_for<k + 1, 6>([&](auto J) {
                constexpr int j = decltype(J)::value;
                ar[i][j] -= lr * ar[k][j] - li * ai[k][j];
                ai[i][j] -= lr * ai[k][j] + li * ar[k][j];
            });
            br[i] -= lr * br[k] - li * bi[k];
            bi[i] -= lr * bi[k] + li * br[k];
        });
    });
    static_for<0, 6>([&](auto II) {
        constexpr int i = 5 - decltype(II)::value;
        double sr = br[i], si = bi[i];
        static_for<i + 1, 6>([&](auto J) {
            constexpr int j = decltype(J)::value;
            sr -= ar[i][j] * br[j] - ai[i][j] * bi[j];
            si -= ar[i][j] * bi[j] + ai[i][j] * br[j];
        });
        br[i] = sr * rr[i] - si * ri[i];
        bi[i] = sr * ri[i] + si * rr[i];
    });
    return ok;
}

// shared-memory carve-up (doubles unless noted); sizes depend on max_members / max_nodes / nwl
struct Smem {
    double *mem;        // [Nm][MEM_STRIDE]: q,p1,p2, axq, axp1, axp2, hq,h1,h2
    double *node;       // [7][NsP]: ls, cdq, cd1, cd2, bq, b1, b2
    double *msum;       // [Nm][8]: sum bq, sum b1, sum b1 ls, sum b1 ls^2, sum b2, sum b2 ls, sum b2 ls^2
    double *mat;        // [3][36]: M0, B0 + B_drag, C0
    double *warp_part;  // [nchunk][nwarps][32]
    double *sums;       // [2][nchunk*32 + 2]  (this CTA's partial sums + flags, double buffered)
    double *tot;        // [nchunk*32]
    double *xi;         // [12][nwl]
    int *imem;          // [Nm][3]: node start, node end (local), circ
};

__host__ __device__ inline size_t smem_doubles(int Nm, int NsP, int nchunk, int nwarps, int nwl)
{
    return (size_t)Nm * MEM_STRIDE + 7 * (size_t)NsP + (size_t)Nm * 8 + 108 + (size_t)nchunk * nwarps * 32
           + 2 * ((size_t)nchunk * 32 + 2) + (size_t)nchunk * 32 + 12 * (size_t)nwl;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 2)
k_drag_solve(DesignsDev D, CasesDev Cs, Work W, SolveParams P)
{
    extern __shared__ __align__(16) double smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = P.CS;
    const int rank = (CS > 1) ? (int)cluster.block_rank() : 0;
    const int unit_l = blockIdx.x / CS;                 // local unit in this chunk
    const int dl = unit_l / Cs.nC, c = unit_l % Cs.nC, d = W.d0 + dl;
    const int nw = D.nw, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = SOLVE_THREADS / 32;

    const int m0 = D.member_offset[d], Nm = D.member_offset[d + 1] - m0;
    const int nbase = D.mem_node_start[m0];
    const int Ns = D.mem_node_start[m0 + Nm] - nbase;
    const int NsP = D.max_nodes;
    const int nchunk = (D.max_nodes + CHUNK_NODES - 1) / CHUNK_NODES;
    const int nwl = P.nwl;
    const int f_begin = rank * nwl;
    const int nloc = max(0, min(nwl, nw - f_begin));     // frequencies owned by this CTA

    Smem S;
    {
        double *p = smem_raw;
        S.mem = p; p += (size_t)D.max_members * MEM_STRIDE;
        S.node = p; p += 7 * (size_t)NsP;
        S.msum = p; p += (size_t)D.max_members * 8;
        S.mat = p; p += 108;
        S.warp_part = p; p += (size_t)nchunk * nwarps * 32;
        S.sums = p; p += 2 * ((size_t)nchunk * 32 + 2);
        S.tot = p; p += (size_t)nchunk * 32;
        S.xi = p; p += 12 * (size_t)nwl;
        S.imem = reinterpret_cast<int *>(p);
    }
    const int sums_stride = nchunk * 32 + 2;

    // ---- stage design tables (members, nodes, matrices) ----------------------------------------
    const double beta = Cs.beta_deg[c] * (CUDART_PI / 180.0);
    double sb, cb;
    sincos(beta, &sb, &cb);
    for (int m = tid; m < Nm; m += SOLVE_THREADS) {
        const double *fr = D.mem_frame + 9 * (m0 + m);
        const double *arm = D.mem_arm + 3 * (m0 + m);
        double *o = S.mem + m * MEM_STRIDE;
        for (int t = 0; t < 9; t++) o[t] = fr[t];
        for (int v = 0; v < 3; v++) {                    // a x d for d = q, p1, p2
            const double d0_ = fr[3 * v], d1_ = fr[3 * v + 1], d2_ = fr[3 * v + 2];
            o[9 + 3 * v + 0] = arm[1] * d2_ - arm[2] * d1_;
            o[9 + 3 * v + 1] = arm[2] * d0_ - arm[0] * d2_;
            o[9 + 3 * v + 2] = arm[0] * d1_ - arm[1] * d0_;
            o[18 + v] = d0_ * cb + d1_ * sb;             // h_d
        }
        S.imem[3 * m + 0] = D.mem_node_start[m0 + m] - nbase;
        S.imem[3 * m + 1] = D.mem_node_start[m0 + m + 1] - nbase;
        S.imem[3 * m + 2] = D.mem_circ[m0 + m];
    }
    for (int j = tid; j < NsP; j += SOLVE_THREADS) {
        const bool in = j < Ns;
        S.node[0 * NsP + j] = in ? D.node_ls[nbase + j] : 0.0;
        S.node[1 * NsP + j] = in ? D.node_cd_q[nbase + j] : 0.0;
        S.node[2 * NsP + j] = in ? D.node_cd_p1[nbase + j] : 0.0;
        S.node[3 * NsP + j] = in ? D.node_cd_p2[nbase + j] : 0.0;
    }
    for (int t = tid; t < 36; t += SOLVE_THREADS) {
        S.mat[t] = D.M0[(size_t)d * 36 + t];
        S.mat[72 + t] = D.C0[(size_t)d * 36 + t];
    }
    // initial response guess (raft_model.py:999) or the given Xi (mode 1)
    const size_t ogl = ((size_t)d * Cs.nC + c) * 6 * nw;
    for (int t = tid; t < nloc; t += SOLVE_THREADS) {
        for (int a = 0; a < 6; a++) {
            double xr = P.xi_start, xi = 0.0;
            if (P.mode == 1) { const double2 v = P.Xi_in[ogl + (size_t)a * nw + f_begin + t]; xr = v.x; xi = v.y; }
            S.xi[(2 * a) * nwl + t] = xr; S.xi[(2 * a + 1) * nwl + t] = xi;
        }
    }
    __syncthreads();

    const size_t unit = (size_t)dl * Cs.nC + c;
    const double2 *ptab = W.phase_tab + unit * D.max_nodes * nw;
    const double2 *dtab = W.depth_tab + (size_t)dl * D.max_nodes * nw;
    const double2 *F0 = W.F0 + unit * 6 * nw;
    const double *Aw = D.A_w ? D.A_w + (size_t)d * 36 * nw : nullptr;
    const double *Bw = D.B_w ? D.B_w + (size_t)d * 36 * nw : nullptr;

    int passes = 0, converged = 0, flags = 0, par = 0;
    const int max_pass = (P.mode == 1) ? 1 : P.n_iter + 1;

    for (int it = 0; it < max_pass; it++) {
        // ================= pass part 1: sum_w |v_rel . d|^2 per node and direction =================
        for (int ch = 0; ch < nchunk; ch++) {
            double acc[32];
#pragma unroll
            for (int t = 0; t < 32; t++) acc[t] = 0.0;
            const int jc0 = ch * CHUNK_NODES;
            if (jc0 < Ns) {
                for (int t = tid; t < nloc; t += SOLVE_THREADS) {
                    const int i = f_begin + t;
                    const double w = D.w[i];
                    double xr[6], xi[6];
#pragma unroll
                    for (int a = 0; a < 6; a++) { xr[a] = S.xi[(2 * a) * nwl + t]; xi[a] = S.xi[(2 * a + 1) * nwl + t]; }
                    int mcur = -1, mend = 0;
                    double hq = 0, h1 = 0, h2 = 0, dzq = 0, dz1 = 0, dz2 = 0;
                    double mqr = 0, mqi = 0, m1r = 0, m1i = 0, m2r = 0, m2i = 0, t1r = 0, t1i = 0, t2r = 0, t2i = 0;
#pragma unroll
                    for (int jj = 0; jj < CHUNK_NODES; jj++) {
                        const int j = jc0 + jj;
                        if (j < Ns) {
                            if (j >= mend) {        // (uniform) entered a new member: member-level projections of the body velocity
                                do { mcur++; mend = S.imem[3 * mcur + 1]; } while (j >= mend);
                                const double *o = S.mem + mcur * MEM_STRIDE;
                                double sr, si;
                                // -i w (d . Xi_t + (a x d) . Xi_r)
                                sr = o[0] * xr[0] + o[1] * xr[1] + o[2] * xr[2] + o[9] * xr[3] + o[10] * xr[4] + o[11] * xr[5];
                                si = o[0] * xi[0] + o[1] * xi[1] + o[2] * xi[2] + o[9] * xi[3] + o[10] * xi[4] + o[11] * xi[5];
                                mqr = w * si; mqi = -w * sr;
                                sr = o[3] * xr[0] + o[4] * xr[1] + o[5] * xr[2] + o[12] * xr[3] + o[13] * xr[4] + o[14] * xr[5];
                                si = o[3] * xi[0] + o[4] * xi[1] + o[5] * xi[2] + o[12] * xi[3] + o[13] * xi[4] + o[14] * xi[5];
                                m1r = w * si; m1i = -w * sr;
                                sr = o[6] * xr[0] + o[7] * xr[1] + o[8] * xr[2] + o[15] * xr[3] + o[16] * xr[4] + o[17] * xr[5];
                                si = o[6] * xi[0] + o[7] * xi[1] + o[8] * xi[2] + o[15] * xi[3] + o[16] * xi[4] + o[17] * xi[5];
                                m2r = w * si; m2i = -w * sr;
                                sr = o[3] * xr[3] + o[4] * xr[4] + o[5] * xr[5];     // p1 . Xi_r
                                si = o[3] * xi[3] + o[4] * xi[4] + o[5] * xi[5];
                                t1r = w * si; t1i = -w * sr;
                                sr = o[6] * xr[3] + o[7] * xr[4] + o[8] * xr[5];     // p2 . Xi_r
                                si = o[6] * xi[3] + o[7] * xi[4] + o[8] * xi[5];
                                t2r = w * si; t2i = -w * sr;
                                hq = o[18]; h1 = o[19]; h2 = o[20]; dzq = o[2]; dz1 = o[5]; dz2 = o[8];
                            }
                            const double ls = S.node[j];
                            const double2 e = ptab[(size_t)j * nw + i];
                            const double2 cs = dtab[(size_t)j * nw + i];
                            double gr, gi, ar_, ai_;
                            gr = cs.x * hq; gi = cs.y * dzq;
                            ar_ = e.x * gr - e.y * gi + mqr; ai_ = e.x * gi + e.y * gr + mqi;
                            acc[3 * jj + 0] += ar_ * ar_ + ai_ * ai_;
                            gr = cs.x * h1; gi = cs.y * dz1;
                            ar_ = e.x * gr - e.y * gi + m1r + ls * t2r; ai_ = e.x * gi + e.y * gr + m1i + ls * t2i;
                            acc[3 * jj + 1] += ar_ * ar_ + ai_ * ai_;
                            gr = cs.x * h2; gi = cs.y * dz2;
                            ar_ = e.x * gr - e.y * gi + m2r - ls * t1r; ai_ = e.x * gi + e.y * gr + m2i - ls * t1i;
                            acc[3 * jj + 2] += ar_ * ar_ + ai_ * ai_;
                        }
                    }
                }
            }
            const double r = warp_multi_reduce32(acc);
            S.warp_part[((size_t)ch * nwarps + warp) * 32 + lane] = r;
        }
        __syncthreads();
        for (int t = tid; t < nchunk * 32; t += SOLVE_THREADS) {
            const int ch = t >> 5, l = t & 31;
            double s = 0.0;
            for (int wv = 0; wv < nwarps; wv++) s += S.warp_part[((size_t)ch * nwarps + wv) * 32 + l];
            S.sums[par * sums_stride + t] = s;
        }
        if (CS > 1) {
            cluster.sync();
            for (int t = tid; t < nchunk * 32; t += SOLVE_THREADS) {
                double s = 0.0;
                for (int r = 0; r < CS; r++) {
                    const double *rem = cluster.map_shared_rank(S.sums, r);
                    s += rem[par * sums_stride + t];
                }
                S.tot[t] = s;
            }
        } else {
            __syncthreads();
            for (int t = tid; t < nchunk * 32; t += SOLVE_THREADS) S.tot[t] = S.sums[par * sums_stride + t];
        }
        __syncthreads();

        // ================= linearised coefficients per node, member sums, B_drag ===================
        for (int j = tid; j < Ns; j += SOLVE_THREADS) {
            const int ch = j / CHUNK_NODES, jj = j - ch * CHUNK_NODES;
            const double sq = S.tot[ch * 32 + 3 * jj], s1 = S.tot[ch * 32 + 3 * jj + 1], s2 = S.tot[ch * 32 + 3 * jj + 2];
            int m = 0;
            while (j >= S.imem[3 * m + 1]) m++;
            const bool circ = S.imem[3 * m + 2] != 0;
            // getRMS (helpers.py:684): sqrt(0.5*sum |.|^2); circular members use the total transverse RMS
            const double vq = sqrt(0.5 * sq);
            const double v1 = circ ? sqrt(0.5 * (s1 + s2)) : sqrt(0.5 * s1);
            const double v2 = circ ? v1 : sqrt(0.5 * s2);
            S.node[4 * NsP + j] = S.node[1 * NsP + j] * vq;
            S.node[5 * NsP + j] = S.node[2 * NsP + j] * v1;
            S.node[6 * NsP + j] = S.node[3 * NsP + j] * v2;
        }
        __syncthreads();
        for (int m = tid; m < Nm; m += SOLVE_THREADS) {
            double bq = 0, b1 = 0, b1l = 0, b1ll = 0, b2 = 0, b2l = 0, b2ll = 0;
            for (int j = S.imem[3 * m]; j < S.imem[3 * m + 1]; j++) {
                const double ls = S.node[j], q_ = S.node[4 * NsP + j], p1_ = S.node[5 * NsP + j], p2_ = S.node[6 * NsP + j];
                bq += q_; b1 += p1_; b1l += p1_ * ls; b1ll += p1_ * ls * ls; b2 += p2_; b2l += p2_ * ls; b2ll += p2_ * ls * ls;
            }
            double *o = S.msum + m * 8;
            o[0] = bq; o[1] = b1; o[2] = b1l; o[3] = b1ll; o[4] = b2; o[5] = b2l; o[6] = b2ll;
        }
        __syncthreads();
        if (tid < 36) {
            const int a = tid / 6, b = tid % 6;
            double s = 0.0;
            for (int m = 0; m < Nm; m++) {
                const double *o = S.mem + m * MEM_STRIDE, *ms = S.msum + m * 8;
                // V_q = [q ; a x q]; V_1 = [p1 ; a x p1] + ls [0 ; p2]; V_2 = [p2 ; a x p2] - ls [0 ; p1]
                const double vqa = a < 3 ? o[a] : o[9 + a - 3], vqb = b < 3 ? o[b] : o[9 + b - 3];
                const double v1a = a < 3 ? o[3 + a] : o[12 + a - 3], v1b = b < 3 ? o[3 + b] : o[12 + b - 3];
                const double v2a = a < 3 ? o[6 + a] : o[15 + a - 3], v2b = b < 3 ? o[6 + b] : o[15 + b - 3];
                const double u1a = a < 3 ? 0.0 : o[6 + a - 3], u1b = b < 3 ? 0.0 : o[6 + b - 3];       // +p2
                const double u2a = a < 3 ? 0.0 : -o[3 + a - 3], u2b = b < 3 ? 0.0 : -o[3 + b - 3];     // -p1
                s += ms[0] * vqa * vqb;
                s += ms[1] * v1a * v1b + ms[2] * (v1a * u1b + u1a * v1b) + ms[3] * u1a * u1b;
                s += ms[4] * v2a * v2b + ms[5] * (v2a * u2b + u2a * v2b) + ms[6] * u2a * u2b;
            }
            S.mat[36 + tid] = D.B0[(size_t)d * 36 + tid] + s;
            if (P.Bdrag_out && rank == 0) P.Bdrag_out[((size_t)d * Cs.nC + c) * 36 + tid] = s;
        }
        __syncthreads();

        // ================= pass part 2: drag excitation, impedance, solve, convergence =============
        int conv_local = 1, nan_local = 0;
        for (int t = tid; t < nloc; t += SOLVE_THREADS) {
            const int i = f_begin + t;
            const double w = D.w[i];
            double br[6], bi[6];
#pragma unroll
            for (int a = 0; a < 6; a++) { br[a] = 0.0; bi[a] = 0.0; }
            for (int m = 0; m < Nm; m++) {
                const double *o = S.mem + m * MEM_STRIDE;
                const double hq = o[18], h1 = o[19], h2 = o[20], dzq = o[2], dz1 = o[5], dz2 = o[8];
                double Aqr = 0, Aqi = 0, A1r = 0, A1i = 0, A2r = 0, A2i = 0, L1r = 0, L1i = 0, L2r = 0, L2i = 0;
                const int j1 = S.imem[3 * m + 1];
#pragma unroll 4
                for (int j = S.imem[3 * m]; j < j1; j++) {
                    const double2 e = ptab[(size_t)j * nw + i];
                    const double2 cs = dtab[(size_t)j * nw + i];
                    const double ls = S.node[j], bq = S.node[4 * NsP + j], b1 = S.node[5 * NsP + j], b2 = S.node[6 * NsP + j];
                    double gr, gi, cr, ci;
                    gr = cs.x * hq; gi = cs.y * dzq; cr = e.x * gr - e.y * gi; ci = e.x * gi + e.y * gr;
                    Aqr += bq * cr; Aqi += bq * ci;
                    gr = cs.x * h1; gi = cs.y * dz1; cr = e.x * gr - e.y * gi; ci = e.x * gi + e.y * gr;
                    cr *= b1; ci *= b1; A1r += cr; A1i += ci; L1r += ls * cr; L1i += ls * ci;
                    gr = cs.x * h2; gi = cs.y * dz2; cr = e.x * gr - e.y * gi; ci = e.x * gi + e.y * gr;
                    cr *= b2; ci *= b2; A2r += cr; A2i += ci; L2r += ls * cr; L2i += ls * ci;
                }
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    br[a] += o[a] * Aqr + o[3 + a] * A1r + o[6 + a] * A2r;
                    bi[a] += o[a] * Aqi + o[3 + a] * A1i + o[6 + a] * A2i;
                    br[3 + a] += o[9 + a] * Aqr + o[12 + a] * A1r + o[15 + a] * A2r + o[6 + a] * L1r - o[3 + a] * L2r;
                    bi[3 + a] += o[9 + a] * Aqi + o[12 + a] * A1i + o[15 + a] * A2i + o[6 + a] * L1i - o[3 + a] * L2i;
                }
            }
            if (P.Fdrag_out)
                for (int a = 0; a < 6; a++) P.Fdrag_out[ogl + (size_t)a * nw + i] = make_double2(br[a], bi[a]);
            if (P.mode == 1) continue;

            // F_tot = F_lin + F_drag (raft_model.py:1081); Z = -w^2 M + i w B + C (:1086)
            double ar[6][6], ai[6][6];
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const double2 f0 = F0[(size_t)a * nw + i];
                br[a] += f0.x; bi[a] += f0.y;
            }
            const double w2 = w * w;
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    double M = S.mat[6 * a + b], B = S.mat[36 + 6 * a + b];
                    if (Aw) M += Aw[(size_t)(6 * a + b) * nw + i];
                    if (Bw) B += Bw[(size_t)(6 * a + b) * nw + i];
                    ar[a][b] = S.mat[72 + 6 * a + b] - w2 * M;
                    ai[a][b] = w * B;
                }
            const bool ok = solve6(ar, ai, br, bi);
            if (!ok) nan_local |= RAFTK_FLAG_SINGULAR;
            // convergence test (raft_model.py:1103-1104) and relaxation (:1133)
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const double lr = S.xi[(2 * a) * nwl + t], li = S.xi[(2 * a + 1) * nwl + t];
                if (isnan(br[a]) || isnan(bi[a])) nan_local |= RAFTK_FLAG_NAN;
                const double dr = br[a] - lr, di = bi[a] - li;
                const double tc = sqrt(dr * dr + di * di) / (sqrt(br[a] * br[a] + bi[a] * bi[a]) + P.tol);
                if (!(tc < P.tol)) conv_local = 0;
                S.xi[(2 * a) * nwl + t] = 0.2 * lr + 0.8 * br[a];
                S.xi[(2 * a + 1) * nwl + t] = 0.2 * li + 0.8 * bi[a];
                P.Xi_out[ogl + (size_t)a * nw + i] = make_double2(br[a], bi[a]);
            }
        }
        passes++;
        if (P.mode == 1) break;

        // ---- all-reduce of (converged, flags) over the CTA and the cluster ----
        int conv_all = __syncthreads_and(conv_local);
        // __syncthreads_or returns a boolean, so reduce the two flag bits separately
        int nan_all = (__syncthreads_or(nan_local & RAFTK_FLAG_NAN) ? RAFTK_FLAG_NAN : 0)
                      | (__syncthreads_or(nan_local & RAFTK_FLAG_SINGULAR) ? RAFTK_FLAG_SINGULAR : 0);
        if (CS > 1) {
            if (tid == 0) { S.sums[par * sums_stride + nchunk * 32] = (double)conv_all; S.sums[par * sums_stride + nchunk * 32 + 1] = (double)nan_all; }
            cluster.sync();
            int ca = 1, na = 0;
            for (int r = 0; r < CS; r++) {
                const double *rem = cluster.map_shared_rank(S.sums, r);
                ca &= (int)rem[par * sums_stride + nchunk * 32];
                na |= (int)rem[par * sums_stride + nchunk * 32 + 1];
            }
            conv_all = ca; nan_all = na;
        }
        par ^= 1;
        flags |= nan_all;
        if (nan_all & RAFTK_FLAG_NAN) break;              // raft_model.py:1098-1099 raises here
        if (conv_all) { converged = 1; break; }
    }
    if (P.status && rank == 0 && tid == 0) {
        int *st = P.status + ((size_t)d * Cs.nC + c) * 4;
        st[0] = passes; st[1] = converged; st[2] = flags; st[3] = 0;
    }
    if (CS > 1) cluster.sync();      // keep shared memory alive until every peer finished reading it
}

// ------------------------------------------------------------------------------------------------
// K2f: fused on-chip solver (v2).  One launch does excitation + the whole fixed-point loop; the
// wave-kinematics of the CTA's frequency slice live in SHARED MEMORY for the whole kernel, in a
// compact member-level form, so the iteration loop touches neither L2 nor HBM:
//   per member and frequency   : E0 = zeta w exp(-i k (x0 cos b + y0 sin b)) at the member's first node
//   per distinct first-node z  : A+-(z0) = (C0 +- S0)/2 from the accurate cosh/sinh ratios
//   per "step class" and freq. : W   = exp(-i k (q_x cos b + q_y sin b) step)   (phase factor)
//                                f+- = exp(+-k q_z step)                        (depth factors)
// and nodes are walked along the member with the geometric recurrences E <- E W, A+- <- A+- f+-,
// C = A+ + A-, S = A+ - A-  (the depth functions cosh/sinh(k(z+h))/sinh(kh) split into their growing
// and decaying exponentials, so there is no cancellation in either walking direction; rounding grows
// ~1 ulp per node).  Distinct steps are deduplicated per design (8 classes for VolturnUS-S).
// Directions that are exactly horizontal (d_z = 0) or vertical (d_x = d_y = 0) take 3-flop
// projections instead of the generic 6-flop complex product.
// ------------------------------------------------------------------------------------------------
struct FusedParams {
    int n_iter, CS, nwl, maxW, maxH, maxZ;
    double tol, xi_start;
    double2 *Xi_out, *Fdrag_out, *Finer_out, *Fbem_out;
    double *Bdrag_out, *zeta_out;
    int *status;
    double2 *F0g;            // [units][6][nw] linear excitation kept in global memory (frees 96 B/bin of smem), or NULL
    double *lin_g;           // [units][NCOEF*max_nodes + 36] linearisation hand-over primary -> secondary wave trains, or NULL
    int phase;               // -1: every case is its own primary; 0: run primaries only; 1: run secondaries only
};

#define IMEM_STRIDE 6      // ints per member: node start, node end, circular, direction kinds, z-class, spare
#define NCOEF 5            // per-node linearised coefficients: bq, b1, ls*b1, b2, ls*b2

struct FSmem {
    double *mem, *node, *coef, *msum, *mat, *warp_part, *sums, *tot, *xi, *f0, *ckpt, *wkey, *hkey, *zkey, *scr, *trans;
    double2 *ebase, *abase, *wtab, *htab;
    int *imem, *node_w, *node_h, *iscr, *cnt;
};

__host__ __device__ inline size_t fused_smem_bytes(int Nm, int NsP, int nchunk, int nwarps, int nwl, int maxW, int maxH, int maxZ, bool f0_smem)
{
    size_t dbl = (size_t)Nm * MEM_STRIDE + 4 * (size_t)NsP + 16 + NCOEF * (size_t)NsP + (size_t)Nm * 8 + 108
                 + (size_t)nchunk * nwarps * 32 + 2 * ((size_t)nchunk * 32 + 2) + (size_t)nchunk * 32 + (size_t)nwarps * 16 * 33
                 + (12 + (f0_smem ? 12 : 0) + 4) * (size_t)nwl + 2 * (size_t)maxW + (size_t)maxH + (size_t)maxZ + 3 * (size_t)NsP
                 + 2 * ((size_t)Nm + maxZ + (maxW + 1) + (maxH + 1)) * nwl;       // +1: identity rows of the factor tables
    size_t ints = (size_t)Nm * IMEM_STRIDE + 4 * (size_t)NsP + 40;
    return dbl * sizeof(double) + ints * sizeof(int) + 32;
}

// projection of the wave velocity on direction d plus a body-velocity term: a = E (C h + i S d_z) + m.
// (A 3-way specialisation on exactly horizontal / vertical directions was measured: ptxas if-converts it
// into predicated code that issues all variants, so the generic 6-flop form is kept.)
__device__ __forceinline__ void proj_add(int, double er, double ei, double Cc, double Sc, double h, double dz,
                                         double mr, double mi, double &ar, double &ai)
{
    const double gr = Cc * h, gi = Sc * dz;
    ar = fma(er, gr, fma(-ei, gi, mr));
    ai = fma(er, gi, fma(ei, gr, mi));
}
__device__ __forceinline__ void proj(int, double er, double ei, double Cc, double Sc, double h, double dz, double &cr, double &ci)
{
    const double gr = Cc * h, gi = Sc * dz;
    cr = fma(er, gr, -ei * gi);
    ci = fma(er, gi, ei * gr);
}

template <int T>
__global__ void __launch_bounds__(T, 256 / T)
k_rao_fused(DesignsDev D, CasesDev Cs, FusedParams P)
{
    extern __shared__ __align__(16) double smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = P.CS;
    const int rank = (CS > 1) ? (int)cluster.block_rank() : 0;
    const int unit = blockIdx.x / CS;
    const int d = unit / Cs.nC, c = unit % Cs.nC;
    const int nw = D.nw, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = T / 32;
    // wave trains: a secondary train reuses the linearisation (per-node coefficients, B_drag) of its primary case
    // (raft_model.py:1200-1236); primaries and secondaries run in two launches (cluster-uniform early exit)
    const int prim = (P.phase >= 0 && Cs.primary) ? Cs.primary[c] : c;
    const bool secondary = prim != c;
    if ((P.phase == 0 && secondary) || (P.phase == 1 && !secondary)) return;

    const int m0 = D.member_offset[d], Nm = D.member_offset[d + 1] - m0;
    const int nbase = D.mem_node_start[m0];
    const int Ns = D.mem_node_start[m0 + Nm] - nbase;
    const int NsP = D.max_nodes, NmP = D.max_members;
    const int nchunk = (D.max_nodes + CHUNK_NODES - 1) / CHUNK_NODES;
    const int nwl = P.nwl;
    const int f_begin = rank * nwl;
    const int nloc = max(0, min(nwl, nw - f_begin));

    FSmem S;
    {
        double *p = smem_raw;
        S.ebase = reinterpret_cast<double2 *>(p); p += 2 * (size_t)NmP * nwl;
        S.abase = reinterpret_cast<double2 *>(p); p += 2 * (size_t)P.maxZ * nwl;
        S.wtab = reinterpret_cast<double2 *>(p); p += 2 * (size_t)(P.maxW + 1) * nwl;
        S.htab = reinterpret_cast<double2 *>(p); p += 2 * (size_t)(P.maxH + 1) * nwl;
        S.mem = p; p += (size_t)NmP * MEM_STRIDE;
        S.node = p; p += 4 * (size_t)NsP + 16;
        S.coef = p; p += NCOEF * (size_t)NsP;
        S.msum = p; p += (size_t)NmP * 8;
        S.mat = p; p += 108;
        S.warp_part = p; p += (size_t)nchunk * nwarps * 32;
        S.sums = p; p += 2 * ((size_t)nchunk * 32 + 2);
        S.tot = p; p += (size_t)nchunk * 32;
        S.xi = p; p += 12 * (size_t)nwl;
        S.f0 = p; p += P.F0g ? 0 : 12 * (size_t)nwl;
        S.ckpt = p; p += 4 * (size_t)nwl;
        S.wkey = p; p += 2 * (size_t)P.maxW;
        S.hkey = p; p += (size_t)P.maxH;
        S.zkey = p; p += (size_t)P.maxZ;
        S.scr = p; p += 3 * (size_t)NsP;
        S.trans = p; p += (size_t)nwarps * 16 * 33;
        S.imem = reinterpret_cast<int *>(p);
        S.node_w = S.imem + (size_t)NmP * IMEM_STRIDE;     // per node: offset (class * nwl) of its step factors,
        S.node_h = S.node_w + NsP + 12;                     // identity row for a member's first node / zero steps
        S.iscr = S.node_h + NsP + 12;       // 2*NsP ints   (+12: the node loop prefetches up to 10 entries ahead)
        S.cnt = S.iscr + 2 * NsP;
    }
    const int sums_stride = nchunk * 32 + 2;

    // ---- stage design tables ---------------------------------------------------------------------
    const double beta = Cs.beta_deg[c] * (CUDART_PI / 180.0);
    double sb, cb;
    sincos(beta, &sb, &cb);
    if (tid < 4) S.cnt[tid] = 0;
    for (int m = tid; m < Nm; m += T) {
        const double *fr = D.mem_frame + 9 * (m0 + m);
        const double *arm = D.mem_arm + 3 * (m0 + m);
        double *o = S.mem + m * MEM_STRIDE;
        int kinds = 0;
        for (int t = 0; t < 9; t++) o[t] = fr[t];
        for (int v = 0; v < 3; v++) {
            const double d0_ = fr[3 * v], d1_ = fr[3 * v + 1], d2_ = fr[3 * v + 2];
            o[9 + 3 * v + 0] = arm[1] * d2_ - arm[2] * d1_;
            o[9 + 3 * v + 1] = arm[2] * d0_ - arm[0] * d2_;
            o[9 + 3 * v + 2] = arm[0] * d1_ - arm[1] * d0_;
            o[18 + v] = d0_ * cb + d1_ * sb;
            int kd = 0;
            if (fabs(d2_) < 1e-14) kd = 1;                               // horizontal direction: S d_z term vanishes
            else if (fabs(d0_) < 1e-14 && fabs(d1_) < 1e-14) kd = 2;     // vertical direction: C h term vanishes
            kinds |= kd << (2 * v);
        }
        const int js = D.mem_node_start[m0 + m] - nbase;
        S.imem[IMEM_STRIDE * m + 0] = js;
        S.imem[IMEM_STRIDE * m + 1] = D.mem_node_start[m0 + m + 1] - nbase;
        S.imem[IMEM_STRIDE * m + 2] = D.mem_circ[m0 + m];
        S.imem[IMEM_STRIDE * m + 3] = kinds;
        o[21] = D.mem_rA[3 * (m0 + m) + 2] + D.node_ls[nbase + js] * fr[2];    // z of the first submerged node
    }
    for (int j = tid; j < NsP; j += T) {
        const bool in = j < Ns;
        S.node[0 * NsP + j] = in ? D.node_ls[nbase + j] : 0.0;
        S.node[1 * NsP + j] = in ? D.node_cd_q[nbase + j] : 0.0;
        S.node[2 * NsP + j] = in ? D.node_cd_p1[nbase + j] : 0.0;
        S.node[3 * NsP + j] = in ? D.node_cd_p2[nbase + j] : 0.0;
    }
    for (int t = tid; t < 36; t += T) {
        S.mat[t] = D.M0[(size_t)d * 36 + t];
        S.mat[72 + t] = D.C0[(size_t)d * 36 + t];
    }
    __syncthreads();
    // ---- step classes, built in parallel (thread per node / member) -----------------------------------
    // A: keys per node
    for (int j = tid; j < Ns; j += T) {
        int m = 0;
        while (j >= S.imem[IMEM_STRIDE * m + 1]) m++;
        const double *o = S.mem + m * MEM_STRIDE;
        double kx = 0, ky = 0, kz = 0;
        if (j > S.imem[IMEM_STRIDE * m]) {
            const double step = S.node[j] - S.node[j - 1];
            kx = o[0] * step; ky = o[1] * step; kz = o[2] * step;
        }
        S.scr[j] = kx; S.scr[NsP + j] = ky; S.scr[2 * NsP + j] = kz;
    }
    __syncthreads();
    // B: representative (first node with the same key)
    for (int j = tid; j < Ns; j += T) {
        const double kx = S.scr[j], ky = S.scr[NsP + j], kz = S.scr[2 * NsP + j];
        int rw = -1, rh = -1;
        if (fabs(kx) > 1e-14 || fabs(ky) > 1e-14) {
            const double tol = 1e-11 * (fabs(kx) + fabs(ky));
            rw = j;
            for (int x = 0; x < j; x++)
                if (fabs(S.scr[x] - kx) <= tol && fabs(S.scr[NsP + x] - ky) <= tol) { rw = x; break; }
        }
        if (fabs(kz) > 1e-14) {
            const double tol = 1e-11 * fabs(kz);
            rh = j;
            for (int x = 0; x < j; x++)
                if (fabs(S.scr[2 * NsP + x] - kz) <= tol) { rh = x; break; }
        }
        S.iscr[j] = rw; S.iscr[NsP + j] = rh;
    }
    __syncthreads();
    // C: class id = rank of the representative among representatives
    for (int j = tid; j < Ns; j += T) {
        const int rw = S.iscr[j], rh = S.iscr[NsP + j];
        int wi = -1, hi = -1;
        if (rw >= 0) { wi = 0; for (int x = 0; x < rw; x++) wi += (S.iscr[x] == x); }
        if (rh >= 0) { hi = 0; for (int x = 0; x < rh; x++) hi += (S.iscr[NsP + x] == x); }
        if (wi >= P.maxW) { wi = 0; S.cnt[2] = 1; }
        if (hi >= P.maxH) { hi = 0; S.cnt[2] = 1; }
        if (rw == j && wi >= 0 && S.cnt[2] == 0) { S.wkey[2 * wi] = S.scr[j]; S.wkey[2 * wi + 1] = S.scr[NsP + j]; atomicMax(&S.cnt[0], wi + 1); }
        if (rh == j && hi >= 0 && S.cnt[2] == 0) { S.hkey[hi] = S.scr[2 * NsP + j]; atomicMax(&S.cnt[1], hi + 1); }
        S.node_w[j] = (wi >= 0 ? wi : P.maxW) * nwl;        // identity row when the phase / depth does not change
        S.node_h[j] = (hi >= 0 ? hi : P.maxH) * nwl;
    }
    for (int j = Ns + tid; j < NsP + 12; j += T) { S.node_w[j] = P.maxW * nwl; S.node_h[j] = P.maxH * nwl; }   // prefetch padding
    // z classes of the members' first nodes
    for (int m = tid; m < Nm; m += T) {
        const double z0 = S.mem[m * MEM_STRIDE + 21];
        int rep = m;
        for (int x = 0; x < m; x++) if (fabs(S.mem[x * MEM_STRIDE + 21] - z0) <= 1e-12 * fmax(1.0, fabs(z0))) { rep = x; break; }
        int zi = 0;
        for (int x = 0; x < rep; x++) {
            const double zx = S.mem[x * MEM_STRIDE + 21];
            bool first = true;
            for (int y = 0; y < x; y++) if (fabs(S.mem[y * MEM_STRIDE + 21] - zx) <= 1e-12 * fmax(1.0, fabs(zx))) { first = false; break; }
            zi += first;
        }
        if (zi >= P.maxZ) { zi = 0; S.cnt[2] = 1; }
        if (rep == m && S.cnt[2] == 0) { S.zkey[zi] = z0; atomicMax(&S.cnt[3], zi + 1); }
        S.imem[IMEM_STRIDE * m + 4] = zi;
    }
    __syncthreads();
    const int nW = S.cnt[0], nH = S.cnt[1], nZ = S.cnt[3];
    const bool plan_overflow = S.cnt[2] != 0;

    // ---- prologue per frequency: sea state, member bases, class factors, excitation F0 -----------
    const size_t ogl = ((size_t)d * Cs.nC + c) * 6 * nw;
    for (int t = tid; t < nloc && !plan_overflow; t += T) {
        const int i = f_begin + t;
        const double w = D.w[i], k = D.k[i];
        const double zeta = sea_state_zeta(Cs, c, i, nw, w, D.dw);
        if (P.zeta_out && d == 0) P.zeta_out[(size_t)c * nw + i] = zeta;
        const double zw = zeta * w;
        const bool deep = k * D.depth > 89.4;
        const double tanh_kh = tanh(k * D.depth);
        for (int x = 0; x < nW; x++) {
            double s_, c_;
            sincos(-(k * (S.wkey[2 * x] * cb + S.wkey[2 * x + 1] * sb)), &s_, &c_);
            S.wtab[x * nwl + t] = make_double2(c_, s_);
        }
        for (int x = 0; x < nH; x++) {
            const double a = k * S.hkey[x];
            S.htab[x * nwl + t] = make_double2(exp(a), exp(-a));
        }
        S.wtab[P.maxW * nwl + t] = make_double2(1.0, 0.0);
        S.htab[P.maxH * nwl + t] = make_double2(1.0, 1.0);
        for (int x = 0; x < nZ; x++) {
            double S_, C_, P_;
            depth_funcs(k, D.depth, S.zkey[x], S_, C_, P_);
            S.abase[x * nwl + t] = make_double2(0.5 * (C_ + S_), 0.5 * (C_ - S_));
        }
        double Fr[6] = {0, 0, 0, 0, 0, 0}, Fi[6] = {0, 0, 0, 0, 0, 0};
        for (int m = 0; m < Nm; m++) {
            const double *o = S.mem + m * MEM_STRIDE;
            const int j0 = S.imem[IMEM_STRIDE * m], j1 = S.imem[IMEM_STRIDE * m + 1];
            const int kinds = S.imem[IMEM_STRIDE * m + 3], kq = kinds & 3, k1 = (kinds >> 2) & 3, k2 = (kinds >> 4) & 3;
            const double *rA = D.mem_rA + 3 * (m0 + m);
            const double ls0 = S.node[j0];
            const double x0 = rA[0] + ls0 * o[0], y0 = rA[1] + ls0 * o[1];
            double se, ce;
            sincos(-(k * (cb * x0 + sb * y0)), &se, &ce);
            double er = zw * ce, ei = zw * se;
            S.ebase[m * nwl + t] = make_double2(er, ei);
            const double2 a0 = S.abase[S.imem[IMEM_STRIDE * m + 4] * nwl + t];
            double ap = a0.x, am = a0.y;
            const double hq = o[18], h1 = o[19], h2 = o[20];
            double Aqr = 0, Aqi = 0, A1r = 0, A1i = 0, A2r = 0, A2i = 0, L1r = 0, L1i = 0, L2r = 0, L2i = 0;
            for (int j = j0; j < j1; j++) {
                {   // step factors (identity at the member's first node)
                    const double2 W = S.wtab[S.node_w[j] + t], H = S.htab[S.node_h[j] + t];
                    const double tr = fma(er, W.x, -ei * W.y); ei = fma(er, W.y, ei * W.x); er = tr;
                    ap *= H.x; am *= H.y;
                }
                const int jg = nbase + j;
                const double inq = D.node_in_q[jg], pa = D.node_pa[jg];
                double in1 = D.node_in_p1[jg], in2 = D.node_in_p2[jg], in1i = 0.0, in2i = 0.0;
                if (D.node_in_p1_w) {
                    const double2 v1 = D.node_in_p1_w[(size_t)jg * nw + i], v2 = D.node_in_p2_w[(size_t)jg * nw + i];
                    in1 = v1.x; in1i = v1.y; in2 = v2.x; in2i = v2.y;
                }
                if (inq != 0.0 || in1 != 0.0 || in2 != 0.0 || in1i != 0.0 || in2i != 0.0 || pa != 0.0) {
                    const double ls = S.node[j], Cc = ap + am, Sc = ap - am;
                    double cr, ci;
                    proj(kq, er, ei, Cc, Sc, hq, o[2], cr, ci);
                    double fqr = -w * inq * ci, fqi = w * inq * cr;
                    proj(k1, er, ei, Cc, Sc, h1, o[5], cr, ci);
                    const double f1r = -w * (in1 * ci + in1i * cr), f1i = w * (in1 * cr - in1i * ci);
                    proj(k2, er, ei, Cc, Sc, h2, o[8], cr, ci);
                    const double f2r = -w * (in2 * ci + in2i * cr), f2i = w * (in2 * cr - in2i * ci);
                    if (pa != 0.0 && w != 0.0) {
                        // dynamic pressure: P = cosh(k(z+h))/cosh(kh) = C tanh(kh); deep-water branch of helpers.py:218
                        double Pd = Cc * tanh_kh;
                        if (deep) Pd = Cc + exp(-k * (rA[2] + ls * o[2] + 2.0 * D.depth));
                        const double sc = pa * Pd / w;
                        fqr = fma(sc, er, fqr); fqi = fma(sc, ei, fqi);
                    }
                    Aqr += fqr; Aqi += fqi; A1r += f1r; A1i += f1i; A2r += f2r; A2i += f2i;
                    L1r += ls * f1r; L1i += ls * f1i; L2r += ls * f2r; L2i += ls * f2i;
                }
            }
#pragma unroll
            for (int a = 0; a < 3; a++) {
                Fr[a] += o[a] * Aqr + o[3 + a] * A1r + o[6 + a] * A2r;
                Fi[a] += o[a] * Aqi + o[3 + a] * A1i + o[6 + a] * A2i;
                Fr[3 + a] += o[9 + a] * Aqr + o[12 + a] * A1r + o[15 + a] * A2r + o[6 + a] * L1r - o[3 + a] * L2r;
                Fi[3 + a] += o[9 + a] * Aqi + o[12 + a] * A1i + o[15 + a] * A2i + o[6 + a] * L1i - o[3 + a] * L2i;
            }
        }
        if (P.Finer_out)
            for (int a = 0; a < 6; a++) P.Finer_out[ogl + (size_t)a * nw + i] = make_double2(Fr[a], Fi[a]);
        if (D.n_bem_head > 0) {
            double Br[6], Bi[6];
            bem_excitation(D, d, i, k, beta, sb, cb, zeta, Br, Bi);
#pragma unroll
            for (int a = 0; a < 6; a++) {
                if (P.Fbem_out) P.Fbem_out[ogl + (size_t)a * nw + i] = make_double2(Br[a], Bi[a]);
                Fr[a] += Br[a]; Fi[a] += Bi[a];
            }
        } else if (P.Fbem_out) {
            for (int a = 0; a < 6; a++) P.Fbem_out[ogl + (size_t)a * nw + i] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int a = 0; a < 6; a++) {
            if (P.F0g) P.F0g[ogl + (size_t)a * nw + i] = make_double2(Fr[a], Fi[a]);
            else { S.f0[(2 * a) * nwl + t] = Fr[a]; S.f0[(2 * a + 1) * nwl + t] = Fi[a]; }
            S.xi[(2 * a) * nwl + t] = P.xi_start; S.xi[(2 * a + 1) * nwl + t] = 0.0;
        }
    }
    __syncthreads();

    const double *Aw = D.A_w ? D.A_w + (size_t)d * 36 * nw : nullptr;
    const double *Bw = D.B_w ? D.B_w + (size_t)d * 36 * nw : nullptr;
    int passes = 0, converged = 0, flags = plan_overflow ? RAFTK_FLAG_PLAN : 0, par = 0;
    const int max_pass = plan_overflow ? 0 : (secondary ? 1 : P.n_iter + 1);
    const size_t lin_stride = (size_t)NCOEF * NsP + 36;
    if (secondary && !plan_overflow) {          // frozen linearisation of the primary case
        const double *src = P.lin_g + ((size_t)d * Cs.nC + prim) * lin_stride;
        for (int t = tid; t < NCOEF * NsP; t += T) S.coef[t] = src[t];
        for (int t = tid; t < 36; t += T) S.mat[36 + t] = src[NCOEF * NsP + t];
        __syncthreads();
    }

    for (int it = 0; it < max_pass; it++) {
        if (!secondary) {
        // ================= pass part 1: sum_w |v_rel . d|^2 per node and direction =================
        for (int ch = 0; ch < nchunk; ch++) {
            double acc[32];
#pragma unroll
            for (int t = 0; t < 32; t++) acc[t] = 0.0;
            const int jc0 = ch * CHUNK_NODES;
            if (jc0 < Ns) {
                for (int t = tid; t < nloc; t += T) {
                    const double w = D.w[f_begin + t];
                    double xr[6], xi[6];
#pragma unroll
                    for (int a = 0; a < 6; a++) { xr[a] = S.xi[(2 * a) * nwl + t]; xi[a] = S.xi[(2 * a + 1) * nwl + t]; }
                    // walking state: restored from the checkpoint when the chunk starts inside a member
                    double er = S.ckpt[t], ei = S.ckpt[nwl + t], ap = S.ckpt[2 * nwl + t], am = S.ckpt[3 * nwl + t];
                    const double2 *wt_ = S.wtab + t, *ht_ = S.htab + t;
                    int mcur = -1, jj = 0;
                    while (jj < CHUNK_NODES && jc0 + jj < Ns) {
                        // (uniform) member entry: member-level projections of the body velocity, -i w (d . Xi_t + (a x d) . Xi_r)
                        const int jfirst = jc0 + jj;
                        do { mcur++; } while (jfirst >= S.imem[IMEM_STRIDE * mcur + 1]);
                        const int mstart = S.imem[IMEM_STRIDE * mcur], jlast = S.imem[IMEM_STRIDE * mcur + 1] - jc0;
                        const int kinds = S.imem[IMEM_STRIDE * mcur + 3], kq = kinds & 3, k1 = (kinds >> 2) & 3, k2 = (kinds >> 4) & 3;
                        const double *o = S.mem + mcur * MEM_STRIDE;
                        double sr, si;
                        sr = o[0] * xr[0] + o[1] * xr[1] + o[2] * xr[2] + o[9] * xr[3] + o[10] * xr[4] + o[11] * xr[5];
                        si = o[0] * xi[0] + o[1] * xi[1] + o[2] * xi[2] + o[9] * xi[3] + o[10] * xi[4] + o[11] * xi[5];
                        const double mqr = w * si, mqi = -w * sr;
                        sr = o[3] * xr[0] + o[4] * xr[1] + o[5] * xr[2] + o[12] * xr[3] + o[13] * xr[4] + o[14] * xr[5];
                        si = o[3] * xi[0] + o[4] * xi[1] + o[5] * xi[2] + o[12] * xi[3] + o[13] * xi[4] + o[14] * xi[5];
                        const double m1r = w * si, m1i = -w * sr;
                        sr = o[6] * xr[0] + o[7] * xr[1] + o[8] * xr[2] + o[15] * xr[3] + o[16] * xr[4] + o[17] * xr[5];
                        si = o[6] * xi[0] + o[7] * xi[1] + o[8] * xi[2] + o[15] * xi[3] + o[16] * xi[4] + o[17] * xi[5];
                        const double m2r = w * si, m2i = -w * sr;
                        sr = o[3] * xr[3] + o[4] * xr[4] + o[5] * xr[5];
                        si = o[3] * xi[3] + o[4] * xi[4] + o[5] * xi[5];
                        const double t1r = w * si, t1i = -w * sr;
                        sr = o[6] * xr[3] + o[7] * xr[4] + o[8] * xr[5];
                        si = o[6] * xi[3] + o[7] * xi[4] + o[8] * xi[5];
                        const double t2r = w * si, t2i = -w * sr;
                        const double hq = o[18], h1 = o[19], h2 = o[20], dzq = o[2], dz1 = o[5], dz2 = o[8];
                        if (jfirst == mstart) {
                            const double2 e0 = S.ebase[mcur * nwl + t], a0 = S.abase[S.imem[IMEM_STRIDE * mcur + 4] * nwl + t];
                            er = e0.x; ei = e0.y; ap = a0.x; am = a0.y;
                        }
                        // node body: branch-free; the step factors / ls of node JJ were loaded one node earlier
                        // (CUR set) and those of node JJ+1 are requested first (NXT set), so the shared-memory
                        // latency overlaps the arithmetic of this node.  Sets alternate with the parity of JJ.
#define P1_NODE(JJ, CW, CH, CL, NW, NH, NL)                                                                        \
    {                                                                                                                \
        const int jn = jc0 + JJ + 1;                                                                                 \
        NW = wt_[S.node_w[jn]]; NH = ht_[S.node_h[jn]]; NL = S.node[jn];                                             \
        { const double tr = fma(er, CW.x, -ei * CW.y); ei = fma(er, CW.y, ei * CW.x); er = tr; }                     \
        ap *= CH.x; am *= CH.y;                                                                                      \
        const double ls = CL, Cc = ap + am, Sc = ap - am;                                                            \
        double ar_, ai_;                                                                                             \
        proj_add(kq, er, ei, Cc, Sc, hq, dzq, mqr, mqi, ar_, ai_);                                                   \
        acc[3 * JJ + 0] = fma(ar_, ar_, fma(ai_, ai_, acc[3 * JJ + 0]));                                             \
        proj_add(k1, er, ei, Cc, Sc, h1, dz1, fma(ls, t2r, m1r), fma(ls, t2i, m1i), ar_, ai_);                       \
        acc[3 * JJ + 1] = fma(ar_, ar_, fma(ai_, ai_, acc[3 * JJ + 1]));                                             \
        proj_add(k2, er, ei, Cc, Sc, h2, dz2, fma(-ls, t1r, m2r), fma(-ls, t1i, m2i), ar_, ai_);                     \
        acc[3 * JJ + 2] = fma(ar_, ar_, fma(ai_, ai_, acc[3 * JJ + 2]));                                             \
    }
                        // Duff-style dispatch: one copy of each node body (static accumulator index), re-entered per member
                        double2 Wa, Ha, Wb, Hb; double La, Lb;
                        if (jj & 1) { Wb = wt_[S.node_w[jfirst]]; Hb = ht_[S.node_h[jfirst]]; Lb = S.node[jfirst]; Wa = Wb; Ha = Hb; La = Lb; }
                        else        { Wa = wt_[S.node_w[jfirst]]; Ha = ht_[S.node_h[jfirst]]; La = S.node[jfirst]; Wb = Wa; Hb = Ha; Lb = La; }
                        switch (jj) {
                        case 0: P1_NODE(0, Wa, Ha, La, Wb, Hb, Lb); jj = 1; if (jlast <= 1) break;
                        case 1: P1_NODE(1, Wb, Hb, Lb, Wa, Ha, La); jj = 2; if (jlast <= 2) break;
                        case 2: P1_NODE(2, Wa, Ha, La, Wb, Hb, Lb); jj = 3; if (jlast <= 3) break;
                        case 3: P1_NODE(3, Wb, Hb, Lb, Wa, Ha, La); jj = 4; if (jlast <= 4) break;
                        case 4: P1_NODE(4, Wa, Ha, La, Wb, Hb, Lb); jj = 5; if (jlast <= 5) break;
                        case 5: P1_NODE(5, Wb, Hb, Lb, Wa, Ha, La); jj = 6; if (jlast <= 6) break;
                        case 6: P1_NODE(6, Wa, Ha, La, Wb, Hb, Lb); jj = 7; if (jlast <= 7) break;
                        case 7: P1_NODE(7, Wb, Hb, Lb, Wa, Ha, La); jj = 8; if (jlast <= 8) break;
                        case 8: P1_NODE(8, Wa, Ha, La, Wb, Hb, Lb); jj = 9; if (jlast <= 9) break;
                        case 9: P1_NODE(9, Wb, Hb, Lb, Wa, Ha, La); jj = 10;
                        }
#undef P1_NODE
                    }
                    S.ckpt[t] = er; S.ckpt[nwl + t] = ei; S.ckpt[2 * nwl + t] = ap; S.ckpt[3 * nwl + t] = am;
                }
            }
            // warp sum of the 30 accumulators through a padded shared-memory transpose (16 values at a time):
            // every lane stores its values, then lane l adds 16 lanes' worth of value (l & 15) -- fixed order.
            {
                double *tr = S.trans + warp * (16 * 33);
                const int row = lane & 15, part = lane >> 4;
#pragma unroll
                for (int half = 0; half < 2; half++) {
#pragma unroll
                    for (int v = 0; v < 16; v++) tr[v * 33 + lane] = acc[half * 16 + v];
                    __syncwarp();
                    double sum = 0.0;
#pragma unroll
                    for (int x = 0; x < 16; x++) sum += tr[row * 33 + part * 16 + x];
                    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                    if (lane < 16) S.warp_part[(ch * nwarps + warp) * 32 + half * 16 + lane] = sum;
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        for (int t = tid; t < nchunk * 32; t += T) {
            const int ch = t >> 5, l = t & 31;
            double s = 0.0;
            for (int wv = 0; wv < nwarps; wv++) s += S.warp_part[(ch * nwarps + wv) * 32 + l];
            S.sums[par * sums_stride + t] = s;
        }
        if (CS > 1) {
            cluster.sync();
            for (int t = tid; t < nchunk * 32; t += T) {
                double s = 0.0;
                for (int r = 0; r < CS; r++) {
                    const double *rem = cluster.map_shared_rank(S.sums, r);
                    s += rem[par * sums_stride + t];
                }
                S.tot[t] = s;
            }
        } else {
            __syncthreads();
            for (int t = tid; t < nchunk * 32; t += T) S.tot[t] = S.sums[par * sums_stride + t];
        }
        __syncthreads();

        // ================= linearised coefficients per node, member sums, B_drag ===================
        for (int j = tid; j < Ns; j += T) {
            const int ch = j / CHUNK_NODES, jj = j - ch * CHUNK_NODES;
            const double sq = S.tot[ch * 32 + 3 * jj], s1 = S.tot[ch * 32 + 3 * jj + 1], s2 = S.tot[ch * 32 + 3 * jj + 2];
            int m = 0;
            while (j >= S.imem[IMEM_STRIDE * m + 1]) m++;
            const bool circ = S.imem[IMEM_STRIDE * m + 2] != 0;
            const double vq = sqrt(0.5 * sq);
            const double v1 = circ ? sqrt(0.5 * (s1 + s2)) : sqrt(0.5 * s1);
            const double v2 = circ ? v1 : sqrt(0.5 * s2);
            const double ls = S.node[j], b1 = S.node[2 * NsP + j] * v1, b2 = S.node[3 * NsP + j] * v2;
            S.coef[0 * NsP + j] = S.node[1 * NsP + j] * vq;
            S.coef[1 * NsP + j] = b1; S.coef[2 * NsP + j] = ls * b1;
            S.coef[3 * NsP + j] = b2; S.coef[4 * NsP + j] = ls * b2;
        }
        __syncthreads();
        for (int m = tid; m < Nm; m += T) {
            double bq = 0, b1 = 0, b1l = 0, b1ll = 0, b2 = 0, b2l = 0, b2ll = 0;
            for (int j = S.imem[IMEM_STRIDE * m]; j < S.imem[IMEM_STRIDE * m + 1]; j++) {
                const double ls = S.node[j], q_ = S.coef[j], p1_ = S.coef[NsP + j], p2_ = S.coef[3 * NsP + j];
                bq += q_; b1 += p1_; b1l += p1_ * ls; b1ll += p1_ * ls * ls; b2 += p2_; b2l += p2_ * ls; b2ll += p2_ * ls * ls;
            }
            double *o = S.msum + m * 8;
            o[0] = bq; o[1] = b1; o[2] = b1l; o[3] = b1ll; o[4] = b2; o[5] = b2l; o[6] = b2ll;
        }
        __syncthreads();
        if (tid < 36) {
            const int a = tid / 6, b = tid % 6;
            double s = 0.0;
            for (int m = 0; m < Nm; m++) {
                const double *o = S.mem + m * MEM_STRIDE, *ms = S.msum + m * 8;
                const double vqa = a < 3 ? o[a] : o[9 + a - 3], vqb = b < 3 ? o[b] : o[9 + b - 3];
                const double v1a = a < 3 ? o[3 + a] : o[12 + a - 3], v1b = b < 3 ? o[3 + b] : o[12 + b - 3];
                const double v2a = a < 3 ? o[6 + a] : o[15 + a - 3], v2b = b < 3 ? o[6 + b] : o[15 + b - 3];
                const double u1a = a < 3 ? 0.0 : o[6 + a - 3], u1b = b < 3 ? 0.0 : o[6 + b - 3];
                const double u2a = a < 3 ? 0.0 : -o[3 + a - 3], u2b = b < 3 ? 0.0 : -o[3 + b - 3];
                s += ms[0] * vqa * vqb;
                s += ms[1] * v1a * v1b + ms[2] * (v1a * u1b + u1a * v1b) + ms[3] * u1a * u1b;
                s += ms[4] * v2a * v2b + ms[5] * (v2a * u2b + u2a * v2b) + ms[6] * u2a * u2b;
            }
            S.mat[36 + tid] = D.B0[(size_t)d * 36 + tid] + s;
            if (P.Bdrag_out && rank == 0) P.Bdrag_out[((size_t)d * Cs.nC + c) * 36 + tid] = s;
        }
        __syncthreads();
        if (P.lin_g && P.phase == 0 && rank == 0) {      // hand the linearisation over to the secondary wave trains
            double *dst = P.lin_g + ((size_t)d * Cs.nC + c) * lin_stride;
            for (int t = tid; t < NCOEF * NsP; t += T) dst[t] = S.coef[t];
            for (int t = tid; t < 36; t += T) dst[NCOEF * NsP + t] = S.mat[36 + t];
        }
        }   // !secondary

        // ================= pass part 2: drag excitation, impedance, solve, convergence =============
        int conv_local = 1, nan_local = 0;
        const double *cq_ = S.coef, *c1_ = S.coef + NsP, *cl1_ = S.coef + 2 * NsP, *c2_ = S.coef + 3 * NsP, *cl2_ = S.coef + 4 * NsP;
        for (int t = tid; t < nloc; t += T) {
            const int i = f_begin + t;
            const double w = D.w[i];
            const double2 *wt_ = S.wtab + t, *ht_ = S.htab + t;
            double br[6], bi[6];
#pragma unroll
            for (int a = 0; a < 6; a++) { br[a] = 0.0; bi[a] = 0.0; }
            for (int m = 0; m < Nm; m++) {
                const double *o = S.mem + m * MEM_STRIDE;
                const double hq = o[18], h1 = o[19], h2 = o[20], dzq = o[2], dz1 = o[5], dz2 = o[8];
                const int kinds = S.imem[IMEM_STRIDE * m + 3], kq = kinds & 3, k1 = (kinds >> 2) & 3, k2 = (kinds >> 4) & 3;
                double Aqr = 0, Aqi = 0, A1r = 0, A1i = 0, A2r = 0, A2i = 0, L1r = 0, L1i = 0, L2r = 0, L2i = 0;
                const int j0 = S.imem[IMEM_STRIDE * m], j1 = S.imem[IMEM_STRIDE * m + 1];
                const double2 e0 = S.ebase[m * nwl + t], a0 = S.abase[S.imem[IMEM_STRIDE * m + 4] * nwl + t];
                double er = e0.x, ei = e0.y, ap = a0.x, am = a0.y;
#pragma unroll 2
                for (int j = j0; j < j1; j++) {
                    {
                        const double2 W = wt_[S.node_w[j]], H = ht_[S.node_h[j]];
                        const double tr = fma(er, W.x, -ei * W.y); ei = fma(er, W.y, ei * W.x); er = tr;
                        ap *= H.x; am *= H.y;
                    }
                    const double bq = cq_[j], b1 = c1_[j], lb1 = cl1_[j], b2 = c2_[j], lb2 = cl2_[j];
                    const double Cc = ap + am, Sc = ap - am;
                    double cr, ci;
                    proj(kq, er, ei, Cc, Sc, hq, dzq, cr, ci);
                    Aqr = fma(bq, cr, Aqr); Aqi = fma(bq, ci, Aqi);
                    proj(k1, er, ei, Cc, Sc, h1, dz1, cr, ci);
                    A1r = fma(b1, cr, A1r); A1i = fma(b1, ci, A1i); L1r = fma(lb1, cr, L1r); L1i = fma(lb1, ci, L1i);
                    proj(k2, er, ei, Cc, Sc, h2, dz2, cr, ci);
                    A2r = fma(b2, cr, A2r); A2i = fma(b2, ci, A2i); L2r = fma(lb2, cr, L2r); L2i = fma(lb2, ci, L2i);
                }
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    br[a] += o[a] * Aqr + o[3 + a] * A1r + o[6 + a] * A2r;
                    bi[a] += o[a] * Aqi + o[3 + a] * A1i + o[6 + a] * A2i;
                    br[3 + a] += o[9 + a] * Aqr + o[12 + a] * A1r + o[15 + a] * A2r + o[6 + a] * L1r - o[3 + a] * L2r;
                    bi[3 + a] += o[9 + a] * Aqi + o[12 + a] * A1i + o[15 + a] * A2i + o[6 + a] * L1i - o[3 + a] * L2i;
                }
            }
            if (P.Fdrag_out) {
#pragma unroll
                for (int a = 0; a < 6; a++) P.Fdrag_out[ogl + (size_t)a * nw + i] = make_double2(br[a], bi[a]);
            }
            double ar[6][6], ai[6][6];
            if (P.F0g) {
#pragma unroll
                for (int a = 0; a < 6; a++) { const double2 f = P.F0g[ogl + (size_t)a * nw + i]; br[a] += f.x; bi[a] += f.y; }
            } else {
#pragma unroll
                for (int a = 0; a < 6; a++) { br[a] += S.f0[(2 * a) * nwl + t]; bi[a] += S.f0[(2 * a + 1) * nwl + t]; }
            }
            const double w2 = w * w;
            if (Aw) {                      // frequency-dependent added mass / damping tables (BEM, aero)
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int b = 0; b < 6; b++) {
                        const double M = S.mat[6 * a + b] + Aw[(size_t)(6 * a + b) * nw + i];
                        const double B = S.mat[36 + 6 * a + b] + Bw[(size_t)(6 * a + b) * nw + i];
                        ar[a][b] = fma(-w2, M, S.mat[72 + 6 * a + b]);
                        ai[a][b] = w * B;
                    }
            } else {
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int b = 0; b < 6; b++) {
                        ar[a][b] = fma(-w2, S.mat[6 * a + b], S.mat[72 + 6 * a + b]);
                        ai[a][b] = w * S.mat[36 + 6 * a + b];
                    }
            }
            const bool ok = solve6(ar, ai, br, bi);
            if (!ok) nan_local |= RAFTK_FLAG_SINGULAR;
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const double lr = S.xi[(2 * a) * nwl + t], li = S.xi[(2 * a + 1) * nwl + t];
                if (isnan(br[a]) || isnan(bi[a])) nan_local |= RAFTK_FLAG_NAN;
                const double dr = br[a] - lr, di = bi[a] - li;
                // raft_model.py:1103: |d| / (|x| + tol) < tol  <=>  |d| < tol |x| + tol^2   (no division: 4 % faster;
                // same decision up to the last ulp)
                if (!(sqrt(dr * dr + di * di) < fma(P.tol, sqrt(br[a] * br[a] + bi[a] * bi[a]), P.tol * P.tol))) conv_local = 0;
                S.xi[(2 * a) * nwl + t] = 0.2 * lr + 0.8 * br[a];
                S.xi[(2 * a + 1) * nwl + t] = 0.2 * li + 0.8 * bi[a];
                P.Xi_out[ogl + (size_t)a * nw + i] = make_double2(br[a], bi[a]);
            }
        }
        passes++;
        int conv_all = __syncthreads_and(conv_local);
        // __syncthreads_or returns a boolean, so reduce the two flag bits separately
        int nan_all = (__syncthreads_or(nan_local & RAFTK_FLAG_NAN) ? RAFTK_FLAG_NAN : 0)
                      | (__syncthreads_or(nan_local & RAFTK_FLAG_SINGULAR) ? RAFTK_FLAG_SINGULAR : 0);
        if (CS > 1) {
            if (tid == 0) { S.sums[par * sums_stride + nchunk * 32] = (double)conv_all; S.sums[par * sums_stride + nchunk * 32 + 1] = (double)nan_all; }
            cluster.sync();
            int ca = 1, na = 0;
            for (int r = 0; r < CS; r++) {
                const double *rem = cluster.map_shared_rank(S.sums, r);
                ca &= (int)rem[par * sums_stride + nchunk * 32];
                na |= (int)rem[par * sums_stride + nchunk * 32 + 1];
            }
            conv_all = ca; nan_all = na;
        }
        par ^= 1;
        flags |= nan_all;
        if (nan_all & RAFTK_FLAG_NAN) break;
        if (conv_all) { converged = 1; break; }
    }
    if (P.status && rank == 0 && tid == 0) {
        int *st = P.status + ((size_t)d * Cs.nC + c) * 4;
        st[0] = secondary ? 0 : passes; st[1] = secondary ? 1 : converged; st[2] = flags; st[3] = secondary ? prim + 1 : 0;
    }
    if (CS > 1) cluster.sync();
}

// ------------------------------------------------------------------------------------------------
// K3: dense complex solve per frequency (farm system response).  One CTA per frequency, matrix in
// shared memory, LU with partial pivoting, nrhs right-hand sides.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_system_solve(int n, int nrhs, double2 *Z, double2 *F, int *info)
{
    extern __shared__ __align__(16) double smem_raw[];
    double2 *A = reinterpret_cast<double2 *>(smem_raw);              // [n][n+nrhs] augmented
    __shared__ int piv_s;
    __shared__ double2 rinv_s;
    const int iw = blockIdx.x, tid = threadIdx.x, nc = n + nrhs;
    double2 *Zg = Z + (size_t)iw * n * n, *Fg = F + (size_t)iw * n * nrhs;
    for (int t = tid; t < n * n; t += blockDim.x) A[(t / n) * nc + (t % n)] = Zg[t];
    for (int t = tid; t < n * nrhs; t += blockDim.x) A[(t / nrhs) * nc + n + (t % nrhs)] = Fg[t];
    __syncthreads();
    int bad = 0;
    for (int k = 0; k < n; k++) {
        if (tid < 32) {                                              // pivot search by warp 0
            double best = -1.0; int p = k;
            for (int r = k + tid; r < n; r += 32) {
                const double t = fabs(A[r * nc + k].x) + fabs(A[r * nc + k].y);
                if (t > best) { best = t; p = r; }
            }
            for (int o = 16; o >= 1; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int op = __shfl_xor_sync(0xffffffffu, p, o);
                if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
            }
            if (tid == 0) {
                piv_s = p;
                const double2 pv = A[p * nc + k];
                const double den = pv.x * pv.x + pv.y * pv.y;
                rinv_s = (den > 0.0) ? make_double2(pv.x / den, -pv.y / den) : make_double2(0.0, 0.0);
                if (!(den > 0.0)) bad = k + 1;
            }
        }
        __syncthreads();
        const int p = piv_s;
        if (p != k) for (int t = tid; t < nc; t += blockDim.x) { const double2 tmp = A[k * nc + t]; A[k * nc + t] = A[p * nc + t]; A[p * nc + t] = tmp; }
        __syncthreads();
        const double2 ri = rinv_s;
        for (int r = k + 1 + tid; r < n; r += blockDim.x) {
            const double2 v = A[r * nc + k];
            A[r * nc + k] = make_double2(v.x * ri.x - v.y * ri.y, v.x * ri.y + v.y * ri.x);
        }
        __syncthreads();
        const int rows = n - k - 1, cols = nc - k - 1;
        for (int t = tid; t < rows * cols; t += blockDim.x) {
            const int r = k + 1 + t / cols, cidx = k + 1 + t % cols;
            const double2 l = A[r * nc + k], u = A[k * nc + cidx];
            double2 v = A[r * nc + cidx];
            v.x -= l.x * u.x - l.y * u.y; v.y -= l.x * u.y + l.y * u.x;
            A[r * nc + cidx] = v;
        }
        __syncthreads();
    }
    // back substitution, one thread per right-hand side
    for (int rh = tid; rh < nrhs; rh += blockDim.x) {
        for (int r = n - 1; r >= 0; r--) {
            double2 s = A[r * nc + n + rh];
            for (int cidx = r + 1; cidx < n; cidx++) {
                const double2 a = A[r * nc + cidx], x = A[cidx * nc + n + rh];
                s.x -= a.x * x.x - a.y * x.y; s.y -= a.x * x.y + a.y * x.x;
            }
            const double2 pv = A[r * nc + r];
            const double den = pv.x * pv.x + pv.y * pv.y;
            A[r * nc + n + rh] = make_double2((s.x * pv.x + s.y * pv.y) / den, (s.y * pv.x - s.x * pv.y) / den);
        }
    }
    __syncthreads();
    for (int t = tid; t < n * nrhs; t += blockDim.x) Fg[t] = A[(t / nrhs) * nc + n + (t % nrhs)];
    if (tid == 0 && info) info[iw] = bad;
}

// ------------------------------------------------------------------------------------------------
// K4: response statistics (std, PSD) -- one CTA per (unit, dof), fixed-order block reduction over frequency
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_response_stats(int nw, double dw, int rot_deg, const double2 *Xi, double *sd, double *psd)
{
    __shared__ double part[4];
    const int row = blockIdx.x, dof = row % 6, tid = threadIdx.x;
    const double scale = (rot_deg && dof >= 3) ? (180.0 / CUDART_PI) : 1.0;      // np.rad2deg
    const double2 *x = Xi + (size_t)row * nw;
    double s = 0.0;
    for (int i = tid; i < nw; i += 128) {
        const double re = x[i].x * scale, im = x[i].y * scale;
        const double a2 = re * re + im * im;
        s += a2;
        if (psd) psd[(size_t)row * nw + i] = 0.5 * a2 / dw;
    }
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) part[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) sd[row] = sqrt(0.5 * (((part[0] + part[1]) + part[2]) + part[3]));
}

// ------------------------------------------------------------------------------------------------
// FP64 FMA peak micro-kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static DesignsDev to_dev(const raftk_designs *d, int max_nodes, int max_members)
{
    DesignsDev D;
    D.nD = d->n_designs; D.nw = d->nw; D.max_nodes = max_nodes; D.max_members = max_members; D.n_bem_head = d->n_bem_head;
    D.depth = d->depth; D.rho = d->rho; D.g = d->g; D.dw = d->dw;
    D.w = d->w; D.k = d->k; D.member_offset = d->member_offset;
    D.mem_frame = d->mem_frame; D.mem_rA = d->mem_rA; D.mem_arm = d->mem_arm;
    D.mem_node_start = d->mem_node_start; D.mem_circ = d->mem_circ;
    D.node_ls = d->node_ls; D.node_cd_q = d->node_cd_q; D.node_cd_p1 = d->node_cd_p1; D.node_cd_p2 = d->node_cd_p2;
    D.node_in_q = d->node_in_q; D.node_in_p1 = d->node_in_p1; D.node_in_p2 = d->node_in_p2; D.node_pa = d->node_pa;
    D.node_in_p1_w = reinterpret_cast<const double2 *>(d->node_in_p1_w);
    D.node_in_p2_w = reinterpret_cast<const double2 *>(d->node_in_p2_w);
    D.M0 = d->M0; D.B0 = d->B0; D.C0 = d->C0; D.A_w = d->A_w; D.B_w = d->B_w;
    D.bem_headings = d->bem_headings; D.X_BEM = d->X_BEM; D.bem_xyh = d->bem_xyh;
    return D;
}

static CasesDev to_dev(const raftk_cases *c)
{
    CasesDev C;
    C.nC = c->n_cases; C.Hs = c->Hs; C.Tp = c->Tp; C.gamma = c->gamma; C.beta_deg = c->beta_deg;
    C.zeta_in = c->zeta; C.spec = c->spec; C.primary = c->primary;
    return C;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t chunk_bytes(int nDc, int nC, int max_nodes, int nw)
{
    size_t b = 0;
    b += align_up((size_t)nDc * max_nodes * nw * sizeof(double2), 256);
    b += align_up((size_t)nDc * nC * max_nodes * nw * sizeof(double2), 256);
    b += align_up((size_t)nDc * nC * 6 * nw * sizeof(double2), 256);
    b += align_up((size_t)nC * nw * sizeof(double), 256);
    return b;
}

extern "C" size_t raftk_workspace_bytes(const raftk_designs *d, int32_t n_cases)
{
    if (!d || d->n_designs <= 0) return 0;
    const size_t cap = (size_t)8 << 30;       // plan at most 8 GiB; larger batches run in design chunks
    const int maxn = d->max_nodes > 0 ? d->max_nodes : 1;
    const size_t full = chunk_bytes(d->n_designs, n_cases, maxn, d->nw);
    const size_t one = chunk_bytes(1, n_cases, maxn, d->nw);
    return full <= cap ? full : std::max(cap, one);
}

struct FPlan { int CS, nwl, T, nchunk, maxW, maxH, maxZ; size_t smem; bool f0_global; };
static bool fused_plan(const raftk_designs *d, int units, int requested_cs, bool have_ws, FPlan &pl);

static int validate(const raftk_designs *d, const raftk_cases *c)
{
    if (!d || !c) return set_err(RAFTK_EINVAL, "null designs/cases");
    if (d->n_designs <= 0 || d->nw <= 0 || c->n_cases <= 0) return set_err(RAFTK_EINVAL, "empty batch (n_designs, nw, n_cases must be > 0)");
    if (d->max_nodes <= 0 || d->max_members <= 0) return set_err(RAFTK_EINVAL, "max_nodes/max_members must be > 0");
    if (d->max_members > 512 || d->max_nodes > 4096) return set_err(RAFTK_EINVAL, "design too large for the shared-memory tables");
    if ((d->node_in_p1_w == nullptr) != (d->node_in_p2_w == nullptr)) return set_err(RAFTK_EINVAL, "node_in_p1_w / node_in_p2_w must both be given or both NULL");
    return 0;
}

static int pick_cluster(int units, int nw, int requested)
{
    if (requested == 1 || requested == 2 || requested == 4 || requested == 8) {
        int cs = requested;
        while (cs > 1 && nw / cs < 32) cs >>= 1;
        return cs;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int cs = 1;
    // fill ~2 CTAs per SM, keep >= 128 frequencies per CTA, and keep 12*nwl doubles of state <= 48 KB
    while (cs < 8 && (units * cs < 2 * sms || nw / cs > 512) && nw / (cs * 2) >= 128) cs <<= 1;
    return cs;
}

struct Plan { int CS, nwl, nchunk; size_t smem; };

static int make_plan(const raftk_designs *d, int units_hint, int requested_cs, Plan &pl)
{
    pl.CS = pick_cluster(units_hint, d->nw, requested_cs);
    pl.nwl = (d->nw + pl.CS - 1) / pl.CS;
    pl.nchunk = (d->max_nodes + CHUNK_NODES - 1) / CHUNK_NODES;
    pl.smem = smem_doubles(d->max_members, d->max_nodes, pl.nchunk, SOLVE_THREADS / 32, pl.nwl) * sizeof(double)
              + (size_t)d->max_members * 3 * sizeof(int) + 16;
    if (pl.smem > 227 * 1024) return set_err(RAFTK_EINVAL, "shared-memory plan exceeds 227 KB (nw per CTA too large)");
    return 0;
}

// ---- fused (v2) planner / launcher -----------------------------------------------------------------

static bool fused_try(const raftk_designs *d, int cs, bool have_ws, FPlan &pl)
{
    pl.CS = cs;
    pl.nwl = (d->nw + cs - 1) / cs;
    pl.T = pl.nwl > 128 ? 256 : 128;
    pl.nchunk = (d->max_nodes + CHUNK_NODES - 1) / CHUNK_NODES;
    pl.maxW = d->max_w_classes > 0 ? d->max_w_classes : d->max_nodes;
    pl.maxH = d->max_h_classes > 0 ? d->max_h_classes : d->max_nodes;
    pl.maxZ = d->max_z_classes > 0 ? std::min(d->max_z_classes, d->max_members) : d->max_members;
    // 255 registers cap residency at 256 threads per SM (measured: 168 registers / 3 CTAs is slower, the LU
    // spills); shared memory must allow 2 CTAs of 128 threads or 1 of 256.  The linear excitation F0 lives in
    // shared memory when it fits, else in the caller's workspace.
    const size_t limit = (pl.T == 128) ? (size_t)112 * 1024 : (size_t)226 * 1024;
    pl.f0_global = false;
    pl.smem = fused_smem_bytes(d->max_members, d->max_nodes, pl.nchunk, pl.T / 32, pl.nwl, pl.maxW, pl.maxH, pl.maxZ, true);
    if (pl.smem > limit && have_ws) {
        pl.f0_global = true;
        pl.smem = fused_smem_bytes(d->max_members, d->max_nodes, pl.nchunk, pl.T / 32, pl.nwl, pl.maxW, pl.maxH, pl.maxZ, false);
    }
    return pl.smem <= limit && pl.nwl <= 2 * pl.T;
}

static bool fused_plan(const raftk_designs *d, int units, int requested_cs, bool have_ws, FPlan &pl)
{
    if (getenv("RAFTK_FORCE_V1")) return false;
    if (requested_cs == 1 || requested_cs == 2 || requested_cs == 4 || requested_cs == 8) {
        int cs = requested_cs;
        while (cs > 1 && d->nw / cs < 32) cs >>= 1;
        return fused_try(d, cs, have_ws, pl);
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    FPlan best; bool have = false;
    if (units >= 4 * sms) {                       // plenty of units: smallest cluster whose slice fits on chip
        for (int cs = 1; cs <= 8 && !have; cs <<= 1) { FPlan t; if (fused_try(d, cs, have_ws, t) && t.nwl <= t.T) { best = t; have = true; } }
        for (int cs = 1; cs <= 8 && !have; cs <<= 1) { FPlan t; if (fused_try(d, cs, have_ws, t)) { best = t; have = true; } }
    } else {                                      // few units: largest cluster that keeps >= 128 bins per CTA
        for (int cs = 8; cs >= 1 && !have; cs >>= 1) {
            if (cs > 1 && d->nw / cs < 128) continue;
            FPlan t; if (fused_try(d, cs, have_ws, t)) { best = t; have = true; }
        }
    }
    if (have) pl = best;
    return have;
}

template <int T>
static int fused_launch(const DesignsDev &D, const CasesDev &C, const FusedParams &P, const FPlan &pl, int units, cudaStream_t st)
{
    static std::mutex mu;
    static size_t smem_set = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (pl.smem > smem_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_rao_fused<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            smem_set = pl.smem;
        }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)((size_t)units * pl.CS), 1, 1);
    cfg.blockDim = dim3(T, 1, 1);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    {
        ProfScope ps(st, 2);
        CUDA_TRY(cudaLaunchKernelEx(&cfg, k_rao_fused<T>, D, C, P));
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return RAFTK_OK;
}

static int run_fused(const raftk_designs *d, const raftk_cases *c, const raftk_solve_opts *o, const raftk_outputs *out,
                     const FPlan &pl, void *workspace, size_t wbytes, cudaStream_t st)
{
    prof_begin_call();
    DesignsDev D = to_dev(d, d->max_nodes, d->max_members);
    CasesDev C = to_dev(c);
    FusedParams P;
    P.n_iter = o->n_iter; P.CS = pl.CS; P.nwl = pl.nwl; P.maxW = pl.maxW; P.maxH = pl.maxH; P.maxZ = pl.maxZ;
    P.tol = o->tol; P.xi_start = o->xi_start;
    P.Xi_out = reinterpret_cast<double2 *>(out->Xi);
    P.Fdrag_out = reinterpret_cast<double2 *>(out->F_drag);
    P.Finer_out = reinterpret_cast<double2 *>(out->F_iner);
    P.Fbem_out = reinterpret_cast<double2 *>(out->F_BEM);
    P.Bdrag_out = out->B_drag; P.zeta_out = out->zeta; P.status = out->status;
    P.F0g = pl.f0_global ? reinterpret_cast<double2 *>(workspace) : nullptr;
    const int units = d->n_designs * c->n_cases;
    P.lin_g = nullptr; P.phase = -1;
    if (c->primary) {                                  // wave trains: primaries first, then the trains that follow them
        const size_t f0b = align_up((size_t)units * 6 * d->nw * sizeof(double2), 256);
        const size_t need = f0b + (size_t)units * ((size_t)NCOEF * d->max_nodes + 36) * sizeof(double);
        if (!workspace || wbytes < need) return set_err(RAFTK_ENOMEM, "wave-train cases need raftk_solve_workspace_bytes() of workspace");
        P.lin_g = reinterpret_cast<double *>(static_cast<char *>(workspace) + f0b);
        for (int phase = 0; phase < 2; phase++) {
            P.phase = phase;
            const int rc = (pl.T == 128) ? fused_launch<128>(D, C, P, pl, units, st) : fused_launch<256>(D, C, P, pl, units, st);
            if (rc) return rc;
        }
        return RAFTK_OK;
    }
    if (pl.T == 128) return fused_launch<128>(D, C, P, pl, units, st);
    return fused_launch<256>(D, C, P, pl, units, st);
}

static int run(const raftk_designs *d, const raftk_cases *c, const raftk_solve_opts *o, const raftk_outputs *out,
               const double *Xi_in, int mode /*0 solve, 1 linearise, 2 excitation only*/, bool do_excitation,
               void *workspace, size_t wbytes, cudaStream_t st)
{
    int rc = validate(d, c);
    if (rc) return rc;
    const int nD = d->n_designs, nC = c->n_cases, nw = d->nw;
    if (mode == 0) {                                   // fused on-chip solver when the slice fits in shared memory
        FPlan fp;
        const bool have_ws = workspace && wbytes >= (size_t)nD * nC * 6 * nw * sizeof(double2);
        if (fused_plan(d, nD * nC, o ? o->cluster_size : 0, have_ws, fp)) return run_fused(d, c, o, out, fp, workspace, wbytes, st);
        if (c->primary) return set_err(RAFTK_EINVAL, "wave-train cases (cases.primary) need the fused solver; the design's frequency slice does not fit on chip");
    }
    if (c->primary && mode != 2) return set_err(RAFTK_EINVAL, "cases.primary is only supported by raftk_solve_dynamics_*");
    if (do_excitation) prof_begin_call();
    DesignsDev D = to_dev(d, d->max_nodes, d->max_members);
    CasesDev C = to_dev(c);
    const size_t one = chunk_bytes(1, nC, d->max_nodes, nw);
    if (!workspace || wbytes < one) return set_err(RAFTK_ENOMEM, "workspace smaller than one design's tables");
    int per = nD;
    while (per > 1 && (chunk_bytes(per, nC, d->max_nodes, nw) > wbytes || per > 65535)) per = (per + 1) / 2;
    if (mode == 1 && !do_excitation && per < nD)
        return set_err(RAFTK_ENOMEM, "linearization needs the whole batch's tables resident in the workspace");

    Plan pl;
    if (mode != 2) {
        rc = make_plan(d, std::min(per, nD) * nC, o ? o->cluster_size : 0, pl);
        if (rc) return rc;
        static std::mutex mu;
        static size_t smem_set = 0;
        std::lock_guard<std::mutex> lk(mu);
        if (pl.smem > smem_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_drag_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
            smem_set = pl.smem;
        }
    }

    for (int d0 = 0; d0 < nD; d0 += per) {
        const int nDc = std::min(per, nD - d0);
        Work W;
        W.d0 = d0; W.nDc = nDc;
        char *p = static_cast<char *>(workspace);
        W.depth_tab = reinterpret_cast<double2 *>(p); p += align_up((size_t)nDc * d->max_nodes * nw * sizeof(double2), 256);
        W.phase_tab = reinterpret_cast<double2 *>(p); p += align_up((size_t)nDc * nC * d->max_nodes * nw * sizeof(double2), 256);
        W.F0 = reinterpret_cast<double2 *>(p); p += align_up((size_t)nDc * nC * 6 * nw * sizeof(double2), 256);
        W.zeta = reinterpret_cast<double *>(p);

        if (do_excitation) {
            dim3 g0((nw + 127) / 128, nDc, 1);
            {
                ProfScope ps(st, 0);
                k_depth_table<<<g0, 128, 0, st>>>(D, W);
            }
            g_launches++;
            ExcOut EO;
            EO.F_iner = reinterpret_cast<double2 *>(out->F_iner);
            EO.F_BEM = reinterpret_cast<double2 *>(out->F_BEM);
            EO.zeta = out->zeta;
            dim3 g1((nw + 127) / 128, nC, nDc);
            {
                ProfScope ps(st, 1);
                k_excitation<<<g1, 128, 0, st>>>(D, C, W, EO);
            }
            g_launches++;
        }
        if (mode != 2) {
            SolveParams P;
            P.n_iter = o ? o->n_iter : 0; P.CS = pl.CS; P.nwl = pl.nwl; P.mode = mode;
            P.tol = o ? o->tol : 0.01; P.xi_start = o ? o->xi_start : 0.0;
            P.Xi_in = reinterpret_cast<const double2 *>(Xi_in);
            P.Xi_out = reinterpret_cast<double2 *>(out->Xi);
            P.Fdrag_out = reinterpret_cast<double2 *>(out->F_drag);
            P.Bdrag_out = out->B_drag;
            P.status = out->status;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)(nDc * nC * pl.CS), 1, 1);
            cfg.blockDim = dim3(SOLVE_THREADS, 1, 1);
            cfg.dynamicSmemBytes = pl.smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = pl.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            {
                ProfScope ps(st, 2);
                CUDA_TRY(cudaLaunchKernelEx(&cfg, k_drag_solve, D, C, W, P));
            }
            g_launches++;
        }
        CUDA_TRY(cudaGetLastError());
    }
    return RAFTK_OK;
}

extern "C" size_t raftk_solve_workspace_bytes(const raftk_designs *d, int32_t n_cases)
{
    if (!d || d->n_designs <= 0 || n_cases <= 0) return 0;
    FPlan fp;
    if (d->max_nodes > 0 && d->max_members > 0 && fused_plan(d, d->n_designs * n_cases, 0, true, fp))
        return align_up((size_t)d->n_designs * n_cases * 6 * d->nw * sizeof(double2), 256)
               + align_up((size_t)d->n_designs * n_cases * ((size_t)NCOEF * d->max_nodes + 36) * sizeof(double), 256);
    return raftk_workspace_bytes(d, n_cases);
}

extern "C" int raftk_hydro_excitation_dev(const raftk_designs *d, const raftk_cases *c, const raftk_outputs *out,
                                          void *workspace, size_t workspace_bytes, void *stream)
{
    if (!out) return set_err(RAFTK_EINVAL, "null outputs");
    if (d && c && chunk_bytes(d->n_designs, c->n_cases, d->max_nodes, d->nw) > workspace_bytes)
        return set_err(RAFTK_ENOMEM, "excitation needs the whole batch's tables in the workspace");
    return run(d, c, nullptr, out, nullptr, 2, true, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int raftk_hydro_linearization_dev(const raftk_designs *d, const raftk_cases *c, const double *Xi_in,
                                             const raftk_outputs *out, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!out || !Xi_in) return set_err(RAFTK_EINVAL, "null outputs / Xi_in");
    return run(d, c, nullptr, out, Xi_in, 1, false, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int raftk_solve_dynamics_dev(const raftk_designs *d, const raftk_cases *c, const raftk_solve_opts *o,
                                        const raftk_outputs *out, void *workspace, size_t workspace_bytes, void *stream)
{
    if (!out || !out->Xi || !out->status || !o) return set_err(RAFTK_EINVAL, "Xi, status and opts are required");
    return run(d, c, o, out, nullptr, 0, true, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---- farm system solve ----------------------------------------------------------------------------
extern "C" int raftk_system_solve_dev(int32_t n, int32_t nw, int32_t nrhs, double *Z, double *F, int32_t *info, void *stream)
{
    if (n <= 0 || nw <= 0 || nrhs <= 0 || !Z || !F) return set_err(RAFTK_EINVAL, "bad system-solve arguments");
    const size_t smem = (size_t)n * (n + nrhs) * sizeof(double2);
    if (smem > 227 * 1024) return set_err(RAFTK_EINVAL, "system too large for the shared-memory solver (n*(n+nrhs)*16 B > 227 KB)");
    static std::mutex mu;
    static size_t smem_set = 48 * 1024;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (smem > smem_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_system_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set = smem;
        }
    }
    k_system_solve<<<nw, 128, smem, (cudaStream_t)stream>>>(n, nrhs, reinterpret_cast<double2 *>(Z), reinterpret_cast<double2 *>(F), info);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return RAFTK_OK;
}

// ---- host-pointer front ends -------------------------------------------------------------------------
struct Arena {
    char *base = nullptr; size_t cap = 0, used = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        if (cudaMalloc(&base, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = bytes;
        return 0;
    }
    void *take(size_t bytes) { void *p = base + used; used += align_up(bytes, 256); return p; }
};
static Arena g_arena;
static std::mutex g_arena_mu;

// Small input arrays (grid, member/node tables, case table: ~30 arrays of a few KB) are gathered in one pinned
// staging block and sent with a single copy into a reserved region at the head of the arena; only large arrays
// (frequency tables, big sweeps) are copied one by one.  This trims ~100 us of per-copy launch overhead per call.
static const size_t SMALL_REGION = (size_t)1 << 20, SMALL_MAX = (size_t)64 << 10;
struct Stager {
    char *host = nullptr;            // pinned, SMALL_REGION bytes
    size_t used = 0;
    bool ensure()
    {
        if (host) return true;
        if (cudaHostAlloc(&host, SMALL_REGION, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); host = nullptr; return false; }
        return true;
    }
};
static Stager g_stage;

template <class T>
static const T *up(Arena &A, const T *h, size_t n, cudaStream_t st, cudaError_t &e)
{
    if (!h || n == 0) return nullptr;
    const size_t bytes = n * sizeof(T);
    if (bytes <= SMALL_MAX && g_stage.host && g_stage.used + align_up(bytes, 256) <= SMALL_REGION) {
        memcpy(g_stage.host + g_stage.used, h, bytes);                 // device twin: A.base + same offset
        const T *dptr = reinterpret_cast<const T *>(A.base + g_stage.used);
        g_stage.used += align_up(bytes, 256);
        return dptr;
    }
    T *dptr = static_cast<T *>(A.take(bytes));
    cudaError_t r = cudaMemcpyAsync(dptr, h, bytes, cudaMemcpyHostToDevice, st);
    if (r != cudaSuccess) e = r;
    return dptr;
}

static cudaError_t flush_small(Arena &A, cudaStream_t st)
{
    if (!g_stage.host || g_stage.used == 0) return cudaSuccess;
    return cudaMemcpyAsync(A.base, g_stage.host, g_stage.used, cudaMemcpyHostToDevice, st);
}

static size_t in_bytes(const raftk_designs *d, const raftk_cases *c)
{
    const size_t nD = d->n_designs, nw = d->nw, Nm = d->n_members_total, Ns = d->n_nodes_total, nC = c->n_cases;
    size_t b = 0;
    auto add = [&](size_t n) { b += align_up(n, 256); };
    add(nw * 8); add(nw * 8); add((nD + 1) * 4); add(Nm * 72); add(Nm * 24); add(Nm * 24); add((Nm + 1) * 4); add(Nm * 4);
    for (int t = 0; t < 8; t++) add(Ns * 8);
    if (d->node_in_p1_w) { add(Ns * nw * 16); add(Ns * nw * 16); }
    add(nD * 288); add(nD * 288); add(nD * 288);
    if (d->A_w) add(nD * 36 * nw * 8);
    if (d->B_w) add(nD * 36 * nw * 8);
    if (d->n_bem_head > 0) { add((size_t)d->n_bem_head * 8); add(nD * d->n_bem_head * 6 * nw * 16); add(nD * 24); }
    for (int t = 0; t < 4; t++) add(nC * 8);
    add(nC * 4); add(nC * 4);
    if (c->zeta) add(nC * nw * 8);
    return b;
}

static int host_run(const raftk_designs *d, const raftk_cases *c, const raftk_solve_opts *o, const raftk_outputs *out,
                    const double *Xi_in, int mode)
{
    int rc = validate(d, c);
    if (rc) return rc;
    if (!out) return set_err(RAFTK_EINVAL, "null outputs");
    std::lock_guard<std::mutex> lk(g_arena_mu);
    const size_t nD = d->n_designs, nw = d->nw, nC = c->n_cases, Nm = d->n_members_total, Ns = d->n_nodes_total;
    const size_t resp = nD * nC * 6 * nw * 16;
    size_t obytes = 0;
    auto oadd = [&](const void *p, size_t n) { if (p) obytes += align_up(n, 256); };
    oadd(out->Xi, resp); oadd(out->status, nD * nC * 16); oadd(out->B_drag, nD * nC * 288); oadd(out->F_drag, resp);
    oadd(out->F_iner, resp); oadd(out->F_BEM, resp); oadd(out->zeta, nC * nw * 8);
    if (Xi_in) obytes += align_up(resp, 256);
    size_t wb = raftk_workspace_bytes(d, (int32_t)nC);
    if (mode != 0) wb = chunk_bytes((int)nD, (int)nC, d->max_nodes, (int)nw);   // single chunk required
    else { FPlan fp; if (fused_plan(d, (int)(nD * nC), o ? o->cluster_size : 0, true, fp)) wb = raftk_solve_workspace_bytes(d, (int32_t)nC); }
    const size_t total = SMALL_REGION + in_bytes(d, c) + obytes + align_up(wb, 256) + 4096;
    if (g_arena.reserve(total)) return set_err(RAFTK_ENOMEM, "device arena allocation failed");
    Arena &A = g_arena;
    A.used = SMALL_REGION;                   // [0, SMALL_REGION) mirrors the pinned staging block
    g_stage.ensure();
    g_stage.used = 0;
    cudaStream_t st = 0;
    cudaError_t e = cudaSuccess;
    raftk_designs dd = *d;
    dd.w = up(A, d->w, nw, st, e); dd.k = up(A, d->k, nw, st, e);
    dd.member_offset = up(A, d->member_offset, nD + 1, st, e);
    dd.mem_frame = up(A, d->mem_frame, Nm * 9, st, e); dd.mem_rA = up(A, d->mem_rA, Nm * 3, st, e);
    dd.mem_arm = up(A, d->mem_arm, Nm * 3, st, e);
    dd.mem_node_start = up(A, d->mem_node_start, Nm + 1, st, e); dd.mem_circ = up(A, d->mem_circ, Nm, st, e);
    dd.node_ls = up(A, d->node_ls, Ns, st, e); dd.node_cd_q = up(A, d->node_cd_q, Ns, st, e);
    dd.node_cd_p1 = up(A, d->node_cd_p1, Ns, st, e); dd.node_cd_p2 = up(A, d->node_cd_p2, Ns, st, e);
    dd.node_in_q = up(A, d->node_in_q, Ns, st, e); dd.node_in_p1 = up(A, d->node_in_p1, Ns, st, e);
    dd.node_in_p2 = up(A, d->node_in_p2, Ns, st, e); dd.node_pa = up(A, d->node_pa, Ns, st, e);
    dd.node_in_p1_w = up(A, d->node_in_p1_w, d->node_in_p1_w ? Ns * nw * 2 : 0, st, e);
    dd.node_in_p2_w = up(A, d->node_in_p2_w, d->node_in_p2_w ? Ns * nw * 2 : 0, st, e);
    dd.M0 = up(A, d->M0, nD * 36, st, e); dd.B0 = up(A, d->B0, nD * 36, st, e); dd.C0 = up(A, d->C0, nD * 36, st, e);
    dd.A_w = up(A, d->A_w, nD * 36 * nw, st, e); dd.B_w = up(A, d->B_w, nD * 36 * nw, st, e);
    if (d->n_bem_head > 0) {
        dd.bem_headings = up(A, d->bem_headings, (size_t)d->n_bem_head, st, e);
        dd.X_BEM = up(A, d->X_BEM, nD * d->n_bem_head * 6 * nw * 2, st, e);
        dd.bem_xyh = up(A, d->bem_xyh, nD * 3, st, e);
    }
    raftk_cases cc = *c;
    cc.Hs = up(A, c->Hs, nC, st, e); cc.Tp = up(A, c->Tp, nC, st, e); cc.gamma = up(A, c->gamma, nC, st, e);
    cc.beta_deg = up(A, c->beta_deg, nC, st, e); cc.spec = up(A, c->spec, nC, st, e);
    cc.zeta = up(A, c->zeta, nC * nw, st, e);
    cc.primary = up(A, c->primary, c->primary ? nC : 0, st, e);
    const double *Xi_in_d = up(A, Xi_in, Xi_in ? nD * nC * 6 * nw * 2 : 0, st, e);
    {
        cudaError_t r = flush_small(A, st);
        if (r != cudaSuccess) e = r;
    }
    if (e != cudaSuccess) return set_err(RAFTK_ECUDA, "H2D copy: %s", cudaGetErrorString(e));
    raftk_outputs od;
    memset(&od, 0, sizeof(od));
    if (out->Xi) od.Xi = static_cast<double *>(A.take(resp));
    if (out->status) od.status = static_cast<int32_t *>(A.take(nD * nC * 16));
    if (out->B_drag) od.B_drag = static_cast<double *>(A.take(nD * nC * 288));
    if (out->F_drag) od.F_drag = static_cast<double *>(A.take(resp));
    if (out->F_iner) od.F_iner = static_cast<double *>(A.take(resp));
    if (out->F_BEM) od.F_BEM = static_cast<double *>(A.take(resp));
    if (out->zeta) od.zeta = static_cast<double *>(A.take(nC * nw * 8));
    void *ws = A.take(wb);
    if (mode == 0) rc = run(&dd, &cc, o, &od, nullptr, 0, true, ws, wb, st);
    else if (mode == 2) rc = run(&dd, &cc, nullptr, &od, nullptr, 2, true, ws, wb, st);
    else {
        rc = run(&dd, &cc, nullptr, &od, nullptr, 2, true, ws, wb, st);
        if (!rc) rc = run(&dd, &cc, nullptr, &od, Xi_in_d, 1, false, ws, wb, st);
    }
    if (rc) return rc;
    auto down = [&](void *h, const void *dv, size_t n) { if (h && dv) { cudaError_t r = cudaMemcpyAsync(h, dv, n, cudaMemcpyDeviceToHost, st); if (r != cudaSuccess) e = r; } };
    down(out->Xi, od.Xi, resp); down(out->status, od.status, nD * nC * 16); down(out->B_drag, od.B_drag, nD * nC * 288);
    down(out->F_drag, od.F_drag, resp); down(out->F_iner, od.F_iner, resp); down(out->F_BEM, od.F_BEM, resp);
    down(out->zeta, od.zeta, nC * nw * 8);
    cudaError_t se = cudaStreamSynchronize(st);
    if (e != cudaSuccess || se != cudaSuccess)
        return set_err(RAFTK_ECUDA, "kernel/D2H: %s", cudaGetErrorString(se != cudaSuccess ? se : e));
    return RAFTK_OK;
}

extern "C" int raftk_hydro_excitation_host(const raftk_designs *d, const raftk_cases *c, const raftk_outputs *out)
{
    return host_run(d, c, nullptr, out, nullptr, 2);
}
extern "C" int raftk_hydro_linearization_host(const raftk_designs *d, const raftk_cases *c, const double *Xi_in, const raftk_outputs *out)
{
    if (!Xi_in) return set_err(RAFTK_EINVAL, "null Xi_in");
    return host_run(d, c, nullptr, out, Xi_in, 1);
}
extern "C" int raftk_solve_dynamics_host(const raftk_designs *d, const raftk_cases *c, const raftk_solve_opts *o, const raftk_outputs *out)
{
    if (!out || !out->Xi || !out->status || !o) return set_err(RAFTK_EINVAL, "Xi, status and opts are required");
    return host_run(d, c, o, out, nullptr, 0);
}

extern "C" int raftk_system_solve_host(int32_t n, int32_t nw, int32_t nrhs, double *Z, double *F, int32_t *info)
{
    if (n <= 0 || nw <= 0 || nrhs <= 0 || !Z || !F) return set_err(RAFTK_EINVAL, "bad system-solve arguments");
    const size_t zb = (size_t)nw * n * n * 16, fb = (size_t)nw * n * nrhs * 16, ib = (size_t)nw * 4;
    double *dZ = nullptr, *dF = nullptr; int32_t *dI = nullptr;
    CUDA_TRY(cudaMalloc(&dZ, zb)); CUDA_TRY(cudaMalloc(&dF, fb)); CUDA_TRY(cudaMalloc(&dI, ib));
    CUDA_TRY(cudaMemcpy(dZ, Z, zb, cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(dF, F, fb, cudaMemcpyHostToDevice));
    int rc = raftk_system_solve_dev(n, nw, nrhs, dZ, dF, dI, nullptr);
    if (!rc) {
        CUDA_TRY(cudaMemcpy(F, dF, fb, cudaMemcpyDeviceToHost));
        if (info) CUDA_TRY(cudaMemcpy(info, dI, ib, cudaMemcpyDeviceToHost));
    }
    cudaFree(dZ); cudaFree(dF); cudaFree(dI);
    return rc;
}

extern "C" int raftk_response_stats_dev(int32_t n_units, int32_t nw, double dw, int32_t rot_deg, const double *Xi,
                                        double *sd, double *psd, void *stream)
{
    if (n_units <= 0 || nw <= 0 || !Xi || !sd || !(dw > 0.0)) return set_err(RAFTK_EINVAL, "bad response-stats arguments");
    k_response_stats<<<(unsigned)n_units * 6u, 128, 0, (cudaStream_t)stream>>>(nw, dw, rot_deg, reinterpret_cast<const double2 *>(Xi), sd, psd);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return RAFTK_OK;
}

extern "C" int raftk_response_stats_host(int32_t n_units, int32_t nw, double dw, int32_t rot_deg, const double *Xi,
                                         double *sd, double *psd)
{
    if (n_units <= 0 || nw <= 0 || !Xi || !sd || !(dw > 0.0)) return set_err(RAFTK_EINVAL, "bad response-stats arguments");
    const size_t xb = (size_t)n_units * 6 * nw * 16, sb = (size_t)n_units * 6 * 8, pb = (size_t)n_units * 6 * nw * 8;
    double *dX = nullptr, *dS = nullptr, *dP = nullptr;
    CUDA_TRY(cudaMalloc(&dX, xb)); CUDA_TRY(cudaMalloc(&dS, sb));
    if (psd) CUDA_TRY(cudaMalloc(&dP, pb));
    CUDA_TRY(cudaMemcpy(dX, Xi, xb, cudaMemcpyHostToDevice));
    int rc = raftk_response_stats_dev(n_units, nw, dw, rot_deg, dX, dS, dP, nullptr);
    if (!rc) {
        CUDA_TRY(cudaMemcpy(sd, dS, sb, cudaMemcpyDeviceToHost));
        if (psd) CUDA_TRY(cudaMemcpy(psd, dP, pb, cudaMemcpyDeviceToHost));
    }
    cudaFree(dX); cudaFree(dS); if (dP) cudaFree(dP);
    return rc;
}

extern "C" void *raftk_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void raftk_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" double raftk_fp64_peak_gflops(int iters)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    double *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * 8) != cudaSuccess) return -1.0;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_fp64_peak<<<blocks, threads>>>(out, 1000);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_fp64_peak<<<blocks, threads>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    g_launches += 2;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    return flops / (ms * 1e-3) * 1e-9;
}
