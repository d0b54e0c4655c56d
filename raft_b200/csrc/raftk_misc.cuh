// raftk_misc.cuh -- k_system_solve (farm 6N system), k_response_stats, k_fp64_peak (included by raftk.cu only).
#pragma once

// ------------------------------------------------------------------------------------------------
// K3: dense complex solve per frequency (farm system response).  One CTA per frequency, matrix in
// shared memory, LU with partial pivoting, nrhs right-hand sides.
// ------------------------------------------------------------------------------------------------
// LU with partial pivoting (|re| + |im| metric, as LAPACK izamax) of the augmented system A [n][nc] (nc = n + nrhs) held in
// shared memory, then back substitution of every right-hand side in place.  All threads of the CTA call it; returns k + 1
// of the first zero pivot (0 = none) to thread 0.
__device__ __forceinline__ int lu_solve_smem(double2 *A, int n, int nc, int nrhs, int *piv_s, double2 *rinv_s)
{
    const int tid = threadIdx.x;
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
                *piv_s = p;
                const double2 pv = A[p * nc + k];
                const double den = pv.x * pv.x + pv.y * pv.y;
                *rinv_s = (den > 0.0) ? make_double2(pv.x / den, -pv.y / den) : make_double2(0.0, 0.0);
                if (!(den > 0.0) && bad == 0) bad = k + 1;
            }
        }
        __syncthreads();
        const int p = *piv_s;
        if (p != k) for (int t = tid; t < nc; t += blockDim.x) { const double2 tmp = A[k * nc + t]; A[k * nc + t] = A[p * nc + t]; A[p * nc + t] = tmp; }
        __syncthreads();
        const double2 ri = *rinv_s;
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
    return bad;
}

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
    const int bad = lu_solve_smem(A, n, nc, nrhs, &piv_s, &rinv_s);
    for (int t = tid; t < n * nrhs; t += blockDim.x) Fg[t] = A[(t / nrhs) * nc + n + (t % nrhs)];
    if (tid == 0 && info) info[iw] = bad;
}

// ------------------------------------------------------------------------------------------------
// K3b: farm system response straight from the per-FOWT solves (raft_model.py:1164-1236), one CTA per (frequency, case):
// Z_sys = blockdiag_i(-w^2 (M0_i + A_w,i) + i w (B0_i + B_drag_i + B_w,i) + C0_i) + (-w^2 M_arr + i w B_arr + C_arr),
// F = F_BEM_i + F_iner_i + F_drag_i (+ F_2nd_i) stacked, Xi_sys = Z_sys^-1 F.  Everything is read from device-resident
// outputs of the drag-linearisation solve: no host assembly of Z, no per-case transfer of nw n^2 complex numbers.
// ------------------------------------------------------------------------------------------------
struct FarmParams {
    int N, nC, nw;
    const double *B_drag;                       // [N][nC][36]
    const double2 *F_drag, *F_iner, *F_BEM;     // [N][nC][6][nw]; F_BEM may be NULL
    const double *M_arr, *B_arr, *C_arr;        // [6N][6N] or NULL
    double2 *Xi;                                // [nC][6N][nw]
    int *info;                                  // [nC][nw] or NULL
};

__global__ void __launch_bounds__(128) k_farm_response(DesignsDev D, CasesDev Cs, FarmParams P)
{
    extern __shared__ __align__(16) double smem_raw[];
    double2 *A = reinterpret_cast<double2 *>(smem_raw);              // [n][n+1]
    __shared__ int piv_s;
    __shared__ double2 rinv_s;
    const int iw = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int n = 6 * P.N, nc = n + 1, nw = P.nw;
    const double w = D.w[iw], w2 = w * w;
    const int cp = Cs.primary ? Cs.primary[c] : c;                   // secondary wave trains use their primary's damping
    for (int t = tid; t < n * n; t += blockDim.x) {
        const int a = t / n, b = t % n, i = a / 6, j = b / 6;
        double zr = 0.0, zi = 0.0;
        if (i == j) {
            const int e = 6 * (a - 6 * i) + (b - 6 * j);
            double M = D.M0[(size_t)i * 36 + e], B = D.B0[(size_t)i * 36 + e] + P.B_drag[((size_t)i * P.nC + cp) * 36 + e];
            if (D.A_w) { M += D.A_w[((size_t)i * 36 + e) * nw + iw]; B += D.B_w[((size_t)i * 36 + e) * nw + iw]; }
            zr = fma(-w2, M, D.C0[(size_t)i * 36 + e]);
            zi = w * B;
        }
        if (P.C_arr) zr += P.C_arr[t];
        if (P.M_arr) zr -= w2 * P.M_arr[t];
        if (P.B_arr) zi += w * P.B_arr[t];
        A[a * nc + b] = make_double2(zr, zi);
    }
    for (int a = tid; a < n; a += blockDim.x) {
        const int i = a / 6, e = a - 6 * i;
        const size_t o = (((size_t)i * P.nC + c) * 6 + e) * nw + iw;
        double2 f = P.F_drag[o];
        const double2 g = P.F_iner[o];
        f.x += g.x; f.y += g.y;
        if (P.F_BEM) { const double2 h = P.F_BEM[o]; f.x += h.x; f.y += h.y; }
        if (Cs.F_2nd) f.x += Cs.F_2nd[o];
        A[a * nc + n] = f;
    }
    __syncthreads();
    const int bad = lu_solve_smem(A, n, nc, 1, &piv_s, &rinv_s);
    for (int a = tid; a < n; a += blockDim.x) P.Xi[((size_t)c * n + a) * nw + iw] = A[a * nc + n];
    if (tid == 0 && P.info) P.info[(size_t)c * nw + iw] = bad;
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
// K5: output-channel statistics -- every channel of saveTurbineOutputs beyond the platform DOFs (nacelle accelerations,
// tower-base moment, raft_fowt.py:2401-2444, 2504-2538) is a linear functional of the response:
// Y(w) = sum_dof coef[design][ch][dof][w] * Xi[design][case][dof][w].  One CTA per (design, case, channel).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_channel_stats(int nC, int nch, int nw, double dw, const double2 *coef, const double2 *Xi,
                                                       double *sd, double *psd, double2 *amp)
{
    __shared__ double part[4];
    const int row = blockIdx.x, ch = row % nch, unit = row / nch, d = unit / nC, tid = threadIdx.x;
    const double2 *cf = coef + ((size_t)d * nch + ch) * 6 * nw;
    const double2 *x = Xi + (size_t)unit * 6 * nw;
    double s = 0.0;
    for (int i = tid; i < nw; i += 128) {
        double yr = 0.0, yi = 0.0;
#pragma unroll
        for (int a = 0; a < 6; a++) {
            const double2 c = cf[(size_t)a * nw + i], v = x[(size_t)a * nw + i];
            yr += c.x * v.x - c.y * v.y; yi += c.x * v.y + c.y * v.x;
        }
        const double a2 = yr * yr + yi * yi;
        s += a2;
        if (psd) psd[(size_t)row * nw + i] = 0.5 * a2 / dw;
        if (amp) amp[(size_t)row * nw + i] = make_double2(yr, yi);
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
// K6: cross-GPU arrival barrier of the fused exchange (raftk_peer_barrier_dev).  Thread p tells rank p that this
// rank's stores of `epoch` are complete (they were issued by earlier kernels of this stream, so they are performed
// before this kernel starts; the release store orders the flag behind them at system scope), then waits for
// rank p's flag in the local copy.  Bounded spin: a dead peer sets *timeout_flag instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------
struct PeerFlags { int n, rank; unsigned epoch; unsigned *flags[RAFTK_MAX_PEERS]; };

__global__ void k_peer_barrier(PeerFlags F, int *timeout_flag)
{
    const int p = threadIdx.x;
    if (p >= F.n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(F.flags[p] + F.rank), "r"(F.epoch) : "memory");
    const unsigned *mine = F.flags[F.rank] + p;
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - F.epoch) >= 0) break;
        if (clock64() - t0 > 8000000000LL) { if (timeout_flag) *timeout_flag = 1; break; }     // ~4 s at 1.9 GHz
        __nanosleep(200);
    }
}
