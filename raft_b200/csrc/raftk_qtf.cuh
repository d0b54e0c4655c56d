// raftk_qtf.cuh -- second-order (difference-frequency) wave forces from an external QTF table:
// FOWT.calcHydroForce_2ndOrd, interpMode 'qtf' (raft_fowt.py:2158-2253)  (included by raftk.cu only).
//
// Work decomposition.  For one (design, case) the reference interpolates the QTF onto the nw x nw model grid and
// sums every upper diagonal:  f(mu) = 4 dw sqrt(sum_i S_i S_{i+mu} |Q(w_i, w_{i+mu})|^2).  Nothing of the nw x nw
// grid is stored here: a warp owns a PAIR of diagonals (mu, nw - mu) -- nw elements together, so every warp has
// the same amount of work -- walks them with lanes on consecutive i, and evaluates the bilinear interpolation
// of the small QTF table on the fly.  Lanes of one step fall into the same table cell (the table is ~40x coarser
// than the grid), so the 4 corner loads (6 DOFs x complex = 96 contiguous bytes each) are L1 broadcasts.
// Per CTA the case's spectrum S, and every bin's table cell + fraction, are staged once in shared memory.
#pragma once

#ifndef QTF_THREADS
#define QTF_THREADS 256
#endif
#ifndef QTF_MIN_CTAS
#define QTF_MIN_CTAS 2          // <= 128 registers: 16 warps per SM hide the L1 latency of the corner loads
#endif
#define QTF_TASKS_PER_WARP 4

struct QtfParams {
    int nD;                 // designs written (F_2nd has a design axis)
    int shared;             // 0: table per design; 1: one table for all designs (computed once per case, stored nD times);
                            // 2: table per (design, case)
    int n2, nh, nw;
    double dw;
    const double *w;        // [nw] model grid
    const double *qw;       // [n2] table frequencies
    const double *qh;       // [nh] table headings [rad]
    const double2 *qtf;     // [nD|1][n2][n2][nh][6]
    double *F2;             // [nD][nC][6][nw]
    double *F2mean;         // [nD][nC][6] or NULL
};

// One lane's share of diagonal mu: sum over i = lane, lane+32, ... of S_i S_{i+mu} |Q(w_i, w_{i+mu})|^2 per DOF
// (MEAN: mu = 0 and the summand is S_i Re Q(w_i, w_i)).  MIX: blend two heading tables (offset dh) with weight hr.
template <bool MIX, bool MEAN>
__device__ __forceinline__ void qtf_diag(const double2 *__restrict__ Q, const double *S0, const double *tt, const int *cell,
                                         int nw, int n2, size_t sj, size_t si, int dh, double hr, int mu, int lane, double (&acc)[6])
{
    for (int i = lane; i < nw - mu; i += 32) {
        const int j = i + mu;
        const int ci = cell[i], cj = cell[j];
        if (ci < 0 || cj < 0) continue;
        const double ti = tt[i], tj = tt[j];
        const double w00 = (1.0 - ti) * (1.0 - tj), w01 = (1.0 - ti) * tj, w10 = ti * (1.0 - tj), w11 = ti * tj;
        const double ss = MEAN ? S0[i] : S0[i] * S0[j];
        const double2 *q00 = Q + ((size_t)ci * n2 + cj) * sj;
#pragma unroll
        for (int a = 0; a < 6; a++) {
            double2 v00 = __ldg(q00 + a), v01 = __ldg(q00 + sj + a), v10 = __ldg(q00 + si + a), v11 = __ldg(q00 + si + sj + a);
            if (MIX) {
                const double2 u00 = __ldg(q00 + dh + a), u01 = __ldg(q00 + sj + dh + a);
                const double2 u10 = __ldg(q00 + si + dh + a), u11 = __ldg(q00 + si + sj + dh + a);
                v00.x = fma(u00.x - v00.x, hr, v00.x); v00.y = fma(u00.y - v00.y, hr, v00.y);
                v01.x = fma(u01.x - v01.x, hr, v01.x); v01.y = fma(u01.y - v01.y, hr, v01.y);
                v10.x = fma(u10.x - v10.x, hr, v10.x); v10.y = fma(u10.y - v10.y, hr, v10.y);
                v11.x = fma(u11.x - v11.x, hr, v11.x); v11.y = fma(u11.y - v11.y, hr, v11.y);
            }
            const double re = ((v00.x * w00 + v01.x * w01) + v10.x * w10) + v11.x * w11;
            if (MEAN) acc[a] = fma(ss, re, acc[a]);
            else {
                const double im = ((v00.y * w00 + v01.y * w01) + v10.y * w10) + v11.y * w11;
                acc[a] = fma(ss, fma(re, re, im * im), acc[a]);
            }
        }
    }
}

template <bool MIX>
__global__ void __launch_bounds__(QTF_THREADS, QTF_MIN_CTAS) k_qtf_force(CasesDev Cs, QtfParams P)
{
    extern __shared__ __align__(16) double qsm[];
    const int nw = P.nw, n2 = P.n2, nh = P.nh;
    double *S0 = qsm, *tt = qsm + nw;
    int *cell = reinterpret_cast<int *>(qsm + 2 * (size_t)nw);
    const int c = blockIdx.y, dz = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- per-CTA tables: spectrum, table cell and fraction of every model bin (RegularGridInterpolator's
    //      find_indices: grid[c] <= x < grid[c+1], c clipped to n2-2; outside the table -> fill value 0) ----
    for (int i = tid; i < nw; i += QTF_THREADS) {
        const double x = P.w[i];
        S0[i] = sea_state_S(Cs, c, i, nw, x, P.dw);
        int ci = -1; double t = 0.0;
        if (x >= P.qw[0] && x <= P.qw[n2 - 1]) {
            int lo = 0, hi = n2 - 1;                     // invariant: qw[lo] <= x, and x < qw[hi] or hi == n2-1
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.qw[mid] <= x) lo = mid; else hi = mid; }
            ci = lo;
            t = (x - P.qw[ci]) / (P.qw[ci + 1] - P.qw[ci]);
        }
        cell[i] = ci; tt[i] = t;
    }
    // ---- heading bracket (scipy interp1d, linear, fill = first / last table outside the range; :2178-2187) ----
    int hl = 0, hh = 0; double hr = 0.0;
    if (MIX) {
        const double beta = Cs.beta_deg[c] * (CUDART_PI / 180.0);
        if (beta < P.qh[0]) { hl = hh = 0; }
        else if (beta > P.qh[nh - 1]) { hl = hh = nh - 1; }
        else {
            int idx = 0;
            while (idx < nh && P.qh[idx] < beta) idx++;      // searchsorted, side left
            idx = min(max(idx, 1), nh - 1);
            hl = idx - 1; hh = idx;
            hr = (beta - P.qh[hl]) / (P.qh[hh] - P.qh[hl]);
        }
    }
    __syncthreads();

    const size_t tab = (P.shared == 1) ? (size_t)0 : (P.shared == 2 ? (size_t)dz * Cs.nC + c : (size_t)dz);
    const double2 *Q = P.qtf + tab * n2 * n2 * nh * 6 + (size_t)hl * 6;
    const size_t sj = (size_t)nh * 6, si = (size_t)n2 * nh * 6;      // strides of the w2 / w1 axes (double2 units)
    const int dh = (hh - hl) * 6;                                    // 0 outside the table's heading range: blend of a table with itself
    const int ntasks = nw / 2 + 1;
    const int task0 = (blockIdx.x * (QTF_THREADS / 32) + warp) * QTF_TASKS_PER_WARP;

    for (int tk = task0; tk < min(task0 + QTF_TASKS_PER_WARP, ntasks); tk++) {
        // task 0: the main diagonal (mean drift); task t >= 1: diagonals t and nw - t
        const int ndiag = (tk == 0 || nw - tk == tk) ? 1 : 2;
        for (int which = 0; which < ndiag; which++) {
            const int mu = which == 0 ? tk : nw - tk;
            double acc[6] = {0, 0, 0, 0, 0, 0};
            if (mu == 0) qtf_diag<MIX, true>(Q, S0, tt, cell, nw, n2, sj, si, dh, hr, mu, lane, acc);
            else qtf_diag<MIX, false>(Q, S0, tt, cell, nw, n2, sj, si, dh, hr, mu, lane, acc);
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
            if (lane < 6) {
                double v = acc[0];
#pragma unroll
                for (int a = 1; a < 6; a++) if (lane == a) v = acc[a];
                // f[mu] lands in bin mu - 1 (the reference shifts by one bin, :2244) and the last bin is zero (:2245)
                const double outv = (mu == 0) ? 2.0 * v * P.dw : 4.0 * sqrt(v) * P.dw;
                const int d_lo = (P.shared == 1) ? 0 : dz, d_hi = (P.shared == 1) ? P.nD : dz + 1;
                for (int d = d_lo; d < d_hi; d++) {
                    const size_t unit = (size_t)d * Cs.nC + c;
                    if (mu == 0) {
                        if (P.F2mean) P.F2mean[unit * 6 + lane] = outv;
                        P.F2[(unit * 6 + lane) * nw + (nw - 1)] = 0.0;
                    } else {
                        P.F2[(unit * 6 + lane) * nw + (mu - 1)] = outv;
                    }
                }
            }
        }
    }
}
