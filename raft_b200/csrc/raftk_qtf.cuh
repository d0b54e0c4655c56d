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

// ------------------------------------------------------------------------------------------------
// Tile variant (k_qtf_tiles + k_qtf_finish).  ncu on k_qtf_force shows L1TEX at 94 % while the FP64 pipe idles at
// 35 %: every lane re-loads the four corner values of its table cell for every element (a 128-bit load costs four
// L1 wavefronts even when all lanes read the same address).  Here a warp works on ONE pair of table cells at a time:
// the corners (heading-blended) sit in registers for the whole tile, lanes are 32 consecutive columns j, the rows i of
// the cell are walked in order, and the per-diagonal accumulators are SYSTOLIC: the diagonal mu = j - i held by lane l
// at row i continues in lane l + 1 at row i + 1, so the accumulators shift up one lane per row (6 shuffles), lane 0
// starts a new diagonal and lane 31 retires one into the CTA's shared-memory sums.  Sums of a case are combined across
// its CTAs with atomic adds into the (zeroed) F_2nd buffer; k_qtf_finish turns them into amplitudes, applies the
// one-bin shift and computes the mean drift.  Floating-point addition order therefore varies run to run (1e-16).
// ------------------------------------------------------------------------------------------------
#define QT_THREADS 512
#ifndef QT_GROUPS
#define QT_GROUPS 4             // CTAs per (case, table)
#endif

struct QtfTileParams {
    QtfParams q;
    int ncell;                  // n2 - 1
};

__device__ __forceinline__ void qtf_heading(const CasesDev &Cs, const QtfParams &P, int c, int &hl, int &hh, double &hr)
{
    hl = 0; hh = 0; hr = 0.0;
    if (P.nh > 1) {
        const double beta = Cs.beta_deg[c] * (CUDART_PI / 180.0);
        if (beta < P.qh[0]) { hl = hh = 0; }
        else if (beta > P.qh[P.nh - 1]) { hl = hh = P.nh - 1; }
        else {
            int idx = 0;
            while (idx < P.nh && P.qh[idx] < beta) idx++;      // searchsorted, side left
            idx = min(max(idx, 1), P.nh - 1);
            hl = idx - 1; hh = idx;
            hr = (beta - P.qh[hl]) / (P.qh[hh] - P.qh[hl]);
        }
    }
}

__device__ __forceinline__ void qtf_bin_tables(const CasesDev &Cs, const QtfParams &P, int c, double *S0, double *tt, int *cell, int tid, int nthreads)
{
    for (int i = tid; i < P.nw; i += nthreads) {
        const double x = P.w[i];
        S0[i] = sea_state_S(Cs, c, i, P.nw, x, P.dw);
        int ci = -1; double t = 0.0;
        if (x >= P.qw[0] && x <= P.qw[P.n2 - 1]) {
            int lo = 0, hi = P.n2 - 1;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.qw[mid] <= x) lo = mid; else hi = mid; }
            ci = lo;
            t = (x - P.qw[ci]) / (P.qw[ci + 1] - P.qw[ci]);
        }
        cell[i] = ci; tt[i] = t;
    }
}

__global__ void __launch_bounds__(QT_THREADS, 1) k_qtf_tiles(CasesDev Cs, QtfTileParams TP)
{
    extern __shared__ __align__(16) double qsm[];
    __shared__ int maxw_s;
    const QtfParams &P = TP.q;
    const int nw = P.nw, n2 = P.n2, nh = P.nh, ncell = TP.ncell;
    double *S0 = qsm, *tt = qsm + nw, *f2s = qsm + 2 * (size_t)nw;            // f2s [6][nw]
    int *cell = reinterpret_cast<int *>(qsm + 8 * (size_t)nw);
    int *cstart = cell + nw, *cend = cstart + ncell;
    const int c = blockIdx.y, dz = blockIdx.z, grp = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    qtf_bin_tables(Cs, P, c, S0, tt, cell, tid, QT_THREADS);
    for (int t = tid; t < 6 * nw; t += QT_THREADS) f2s[t] = 0.0;
    for (int t = tid; t < 2 * ncell; t += QT_THREADS) cstart[t] = 0;           // cstart and cend are contiguous
    if (tid == 0) maxw_s = 0;
    __syncthreads();
    for (int i = tid; i < nw; i += QT_THREADS) {
        const int ci = cell[i];
        if (ci >= 0) {
            if (i == 0 || cell[i - 1] != ci) cstart[ci] = i;
            if (i == nw - 1 || cell[i + 1] != ci) cend[ci] = i + 1;
        }
    }
    int hl, hh; double hr;
    qtf_heading(Cs, P, c, hl, hh, hr);
    __syncthreads();
    for (int t = tid; t < ncell; t += QT_THREADS) atomicMax(&maxw_s, cend[t] - cstart[t]);
    __syncthreads();
    const int smax = (maxw_s + 31) / 32;                     // column strips (32 bins) of the widest cell

    const size_t tab = (P.shared == 1) ? (size_t)0 : (P.shared == 2 ? (size_t)dz * Cs.nC + c : (size_t)dz);
    const double2 *Q = P.qtf + tab * n2 * n2 * nh * 6;
    const size_t sj = (size_t)nh * 6, si = (size_t)n2 * nh * 6;
    const bool mix = hh != hl;

    // work items: (ci <= cj, strip of 32 columns of cell cj), dealt round-robin to the warps of the case's CTAs
    const int npair = ncell * (ncell + 1) / 2;
    const int nitems = npair * smax;
    const int nwarps = (QT_THREADS / 32) * QT_GROUPS;
    for (int item = grp * (QT_THREADS / 32) + warp; item < nitems; item += nwarps) {
        const int pr = item / smax, s = item % smax;
        int ci = 0, rem = pr;
        while (rem >= ncell - ci) { rem -= ncell - ci; ci++; }
        const int cjx = ci + rem;
        const int a0 = cstart[ci], a1 = cend[ci], b0 = cstart[cjx], b1 = cend[cjx];
        const int jb = b0 + 32 * s;
        if (a1 <= a0 || jb >= b1) continue;
        if (jb + 31 <= a0) continue;                                    // every column at or below the first row: no j > i
        const int a1e = min(a1, jb + 31);                              // rows i >= jb + 31 have no column j > i in this strip
        const int j = jb + lane;
        const bool jin = j < b1;
        const double sjv = jin ? S0[j] : 0.0, tj = jin ? tt[j] : 0.0;
        const double2 *q00 = Q + ((size_t)ci * n2 + cjx) * sj + (size_t)hl * 6;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            double2 v00[3], v01[3], v10[3], v11[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int ad = 3 * half + a;
                v00[a] = __ldg(q00 + ad); v01[a] = __ldg(q00 + sj + ad); v10[a] = __ldg(q00 + si + ad); v11[a] = __ldg(q00 + si + sj + ad);
                if (mix) {
                    const int dh = (hh - hl) * 6;
                    const double2 u00 = __ldg(q00 + dh + ad), u01 = __ldg(q00 + sj + dh + ad);
                    const double2 u10 = __ldg(q00 + si + dh + ad), u11 = __ldg(q00 + si + sj + dh + ad);
                    v00[a].x = fma(u00.x - v00[a].x, hr, v00[a].x); v00[a].y = fma(u00.y - v00[a].y, hr, v00[a].y);
                    v01[a].x = fma(u01.x - v01[a].x, hr, v01[a].x); v01[a].y = fma(u01.y - v01[a].y, hr, v01[a].y);
                    v10[a].x = fma(u10.x - v10[a].x, hr, v10[a].x); v10[a].y = fma(u10.y - v10[a].y, hr, v10[a].y);
                    v11[a].x = fma(u11.x - v11[a].x, hr, v11[a].x); v11[a].y = fma(u11.y - v11[a].y, hr, v11[a].y);
                }
            }
            double acc[3] = {0.0, 0.0, 0.0}, ret[3] = {0.0, 0.0, 0.0};
            int held = 0;                                               // retired diagonals parked in lanes 0 .. held-1
            for (int i = a0; i < a1e; i++) {
                const double ti = tt[i];
                const double ss = (jin && j > i) ? S0[i] * sjv : 0.0;
                const double w00 = (1.0 - ti) * (1.0 - tj), w01 = (1.0 - ti) * tj, w10 = ti * (1.0 - tj), w11 = ti * tj;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const double re = ((v00[a].x * w00 + v01[a].x * w01) + v10[a].x * w10) + v11[a].x * w11;
                    const double im = ((v00[a].y * w00 + v01[a].y * w01) + v10[a].y * w10) + v11[a].y * w11;
                    acc[a] = fma(ss, fma(re, re, im * im), acc[a]);
                }
                // lane 31's diagonal (mu = jb + 31 - i) is complete for this tile: park it in lane `held` (a single-lane
                // shared-memory atomic per row would serialise the warp), shift the others up one lane, open a new one in lane 0
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const double out = __shfl_sync(0xffffffffu, acc[a], 31);
                    if (lane == held) ret[a] = out;
                    const double up = __shfl_up_sync(0xffffffffu, acc[a], 1);
                    acc[a] = lane == 0 ? 0.0 : up;
                }
                held++;
                if (held == 32 || i == a1e - 1) {                       // lane l parked the diagonal retired at row i - held + 1 + l
                    const int mur = jb + 31 - (i - held + 1 + lane);
                    if (lane < held && mur >= 1 && mur < nw) {
#pragma unroll
                        for (int a = 0; a < 3; a++) if (ret[a] != 0.0) atomicAdd(&f2s[(size_t)(3 * half + a) * nw + mur], ret[a]);
                    }
                    held = 0;
                }
            }
            // after the last row (a1e - 1) and one more shift, lane l holds the diagonal of lane l - 1: mu = jb + l - 1 - (a1e - 1)
            const int mu = jb + lane - a1e;
            if (lane >= 1 && mu >= 1 && mu < nw) {
#pragma unroll
                for (int a = 0; a < 3; a++) if (acc[a] != 0.0) atomicAdd(&f2s[(size_t)(3 * half + a) * nw + mu], acc[a]);
            }
        }
    }
    __syncthreads();
    // combine the CTAs of this (case, table): atomic adds into the zeroed F_2nd rows (design dz, or design 0 when shared)
    const size_t unit = (size_t)((P.shared == 1) ? 0 : dz) * Cs.nC + c;
    for (int t = tid; t < 6 * nw; t += QT_THREADS) {
        const double v = f2s[t];
        if (v != 0.0) atomicAdd(&P.F2[unit * 6 * nw + t], v);
    }
}

// F_2nd rows hold sum_i S_i S_{i+mu} |Q|^2 at index mu: -> 4 sqrt(.) dw shifted by one bin (raft_fowt.py:2236, :2244-2245),
// plus the mean drift 2 dw sum_i S_i Re Q(w_i, w_i) (:2239).  grid (6, nC, nD|1), block 256.
__global__ void __launch_bounds__(256) k_qtf_finish(CasesDev Cs, QtfParams P)
{
    __shared__ double red[8];
    const int a = blockIdx.x, c = blockIdx.y, dz = blockIdx.z, tid = threadIdx.x, nw = P.nw, n2 = P.n2, nh = P.nh;
    const size_t unit = (size_t)((P.shared == 1) ? 0 : dz) * Cs.nC + c;
    double *row = P.F2 + (unit * 6 + a) * nw;
    int hl, hh; double hr;
    qtf_heading(Cs, P, c, hl, hh, hr);
    const size_t tab = (P.shared == 1) ? (size_t)0 : (P.shared == 2 ? (size_t)dz * Cs.nC + c : (size_t)dz);
    const double2 *Q = P.qtf + tab * n2 * n2 * nh * 6 + (size_t)hl * 6 + a;
    const size_t sj = (size_t)nh * 6, si = (size_t)n2 * nh * 6;
    const int dh = (hh - hl) * 6;
    double sm = 0.0;
    for (int i = tid; i < nw; i += 256) {                  // mean drift: diagonal of the interpolated table
        const double x = P.w[i];
        if (x >= P.qw[0] && x <= P.qw[n2 - 1]) {
            int lo = 0, hi = n2 - 1;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.qw[mid] <= x) lo = mid; else hi = mid; }
            const double t = (x - P.qw[lo]) / (P.qw[lo + 1] - P.qw[lo]);
            const double2 *q00 = Q + ((size_t)lo * n2 + lo) * sj;
            double r00 = __ldg(q00).x, r01 = __ldg(q00 + sj).x, r10 = __ldg(q00 + si).x, r11 = __ldg(q00 + si + sj).x;
            if (dh) {
                r00 = fma(__ldg(q00 + dh).x - r00, hr, r00); r01 = fma(__ldg(q00 + sj + dh).x - r01, hr, r01);
                r10 = fma(__ldg(q00 + si + dh).x - r10, hr, r10); r11 = fma(__ldg(q00 + si + sj + dh).x - r11, hr, r11);
            }
            const double re = ((r00 * ((1.0 - t) * (1.0 - t)) + r01 * ((1.0 - t) * t)) + r10 * (t * (1.0 - t))) + r11 * (t * t);
            sm = fma(sea_state_S(Cs, c, i, nw, x, P.dw), re, sm);
        }
    }
    for (int o = 16; o >= 1; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
    if ((tid & 31) == 0) red[tid >> 5] = sm;
    // amplitudes, shifted by one bin: read everything before anything is written (in-place)
    const int per = (nw + 255) / 256;
    double vals[16];
    for (int e = 0; e < per && e < 16; e++) {
        const int m = tid + 256 * e;
        vals[e] = (m < nw - 1) ? row[m + 1] : 0.0;
    }
    __syncthreads();
    double mean = 0.0;
    if (tid == 0) { for (int t = 0; t < 8; t++) mean += red[t]; mean = 2.0 * mean * P.dw; }
    const int d_lo = (P.shared == 1) ? 0 : dz, d_hi = (P.shared == 1) ? P.nD : dz + 1;
    for (int d = d_lo; d < d_hi; d++) {
        const size_t u = (size_t)d * Cs.nC + c;
        double *orow = P.F2 + (u * 6 + a) * nw;
        for (int e = 0; e < per && e < 16; e++) {
            const int m = tid + 256 * e;
            if (m < nw) orow[m] = (m < nw - 1) ? 4.0 * sqrt(vals[e]) * P.dw : 0.0;
        }
        if (tid == 0 && P.F2mean) P.F2mean[u * 6 + a] = mean;
    }
}
