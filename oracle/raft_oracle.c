/*
 * raft_oracle.c -- plain-C CPU restatement of the reference's RAO-solve hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product path (raft_b200/)
 * never links or calls it and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 *   - the reference's own golden pickles (tests/test_data/ *_true_hydroExcitation.pkl,
 *     *_true_hydroLinearization.pkl), re-exported as tests/golden/ *.npz by
 *     tests/golden/make_golden.py, and
 *   - outputs of the unmodified reference run in the build container under the stub harness
 *     (oracle/ref_harness.py), including full Model.solveDynamics responses + iteration counts.
 *
 * Each function cites the reference lines (relative to /root/reference/raft/) it follows.
 * The loop structure, operation order and two-step (member node -> reduced DOF) translation of
 * the reference are kept on purpose: this is the checker for the restructured CUDA kernels.
 *
 * Scope: rigid 6-DOF FOWT (MacCamy-Fuchs Imat as an input table), no underwater rotor; second-order forces only
 * from an external QTF table (potSecOrder 2; raft_fowt.py:2158-2253), not the slender-body QTF (potSecOrder 1)
 * (BASELINE.json configs 1-4; SURVEY.md section 8a rows a1-a11 + 8f row 3).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

/* ---- flat tables (mirrors raft_b200/packer.py; one FOWT design) ------------------------ */
typedef struct {
    int n_nodes, n_members, nw, n_bem_head;
    double depth, rho, g, dw, x_ref, y_ref, heading_adjust;
    const double *prp;        /* [3]   reduced-DOF reference point (global)                 */
    const double *w, *k;      /* [nw]                                                         */
    const double *mem_q, *mem_p1, *mem_p2, *mem_rA; /* [Nm,3]                                 */
    const int *mem_circ;      /* [Nm]                                                         */
    const double *node_r;     /* [Ns,3] global node positions                                 */
    const int *node_mem;      /* [Ns]                                                         */
    const double *node_Imat;  /* [Ns,3,3]                                                     */
    const double *node_a_i;   /* [Ns] signed end area                                         */
    const double *a_q, *a_p1, *a_p2, *a_End;       /* [Ns] drag areas                         */
    const double *Cd_q, *Cd_p1, *Cd_p2, *Cd_End;   /* [Ns] interpolated coefficients          */
    const double *M0, *B0, *C0;                    /* [6,6] row-major                         */
    const double *A_w, *B_w;                       /* [6,6,nw] or NULL (reference layout)     */
    const cplx *X_BEM;                             /* [nhead,6,nw] or NULL (reference layout) */
    const double *bem_headings;                    /* [nhead] deg                             */
    const cplx *node_Imat_w;                       /* [Ns,3,3,nw] MacCamy-Fuchs Imat_MCF or NULL */
    /* external difference-frequency QTF (fowt.qtf after readQTF, raft_fowt.py:2081-2128), or qtf == NULL */
    int n_qtf_w, n_qtf_head;
    const double *qtf_w;                           /* [n_qtf_w] rad/s ascending (w1_2nd == w2_2nd)   */
    const double *qtf_heads;                       /* [n_qtf_head] rad ascending (heads_2nd)        */
    const cplx *qtf;                               /* [n_qtf_w,n_qtf_w,n_qtf_head,6] (reference layout) */
} ro_design;

/* helpers.py:377-392 waveNumber(omega, h, e=0.001) */
double ro_wave_number(double omega, double h)
{
    const double g = 9.81, e = 0.001;
    double k1 = omega * omega / g;
    double k2 = omega * omega / (tanh(k1 * h) * g);
    while (fabs(k2 - k1) / k1 > e) {
        k1 = k2;
        k2 = omega * omega / (tanh(k1 * h) * g);
    }
    return k2;
}

/* helpers.py:703-760 JONSWAP(ws, Hs, Tp, Gamma) */
void ro_jonswap(const double *ws, int nw, double Hs, double Tp, double Gamma, double *S)
{
    if (!(Gamma != 0.0)) {                    /* "if not Gamma" :733 */
        double TpOvrSqrtHs = Tp / sqrt(Hs);
        if (TpOvrSqrtHs <= 3.6) Gamma = 5.0;
        else if (TpOvrSqrtHs >= 5.0) Gamma = 1.0;
        else Gamma = exp(5.75 - 1.15 * TpOvrSqrtHs);
    }
    for (int i = 0; i < nw; i++) {
        double f = 0.5 / M_PI * ws[i];
        double fpOvrf4 = pow(Tp * f, -4.0);
        double C = 1.0 - (0.287 * log(Gamma));
        double Sigma = (f <= 1.0 / Tp) ? 0.07 : 0.09;
        double t = (f * Tp - 1.0) / Sigma;
        double Alpha = exp(-0.5 * t * t);
        S[i] = 0.5 / M_PI * C * 0.3125 * Hs * Hs * fpOvrf4 / f * exp(-1.25 * fpOvrf4) * pow(Gamma, Alpha);
    }
}

/* raft_fowt.py:1759-1774: spectrum -> S, zeta = sqrt(2 S dw).  spec: 0 JONSWAP 1 unit 2 constant 3 none */
int ro_sea_state(const double *w, int nw, double dw, int spec, double Hs, double Tp, double gamma,
                 double *S, double *zeta)
{
    if (spec == 0) ro_jonswap(w, nw, Hs, Tp, gamma, S);
    else if (spec == 1) for (int i = 0; i < nw; i++) S[i] = 1.0;
    else if (spec == 2) for (int i = 0; i < nw; i++) S[i] = Hs;
    else if (spec == 3) for (int i = 0; i < nw; i++) S[i] = 0.0;
    else return -1;
    for (int i = 0; i < nw; i++) zeta[i] = sqrt(2.0 * S[i] * dw);
    return 0;
}

/* helpers.py:188-236 getWaveKin; u, ud are [3][nw], pDyn [nw] */
void ro_wave_kin(const double *zeta0, double beta, const double *w, const double *k, double h,
                 const double *r, int nw, double rho, double g, cplx *u, cplx *ud, cplx *pDyn)
{
    for (int i = 0; i < nw; i++) {
        cplx zeta = zeta0[i] * cexp(-I * (k[i] * (cos(beta) * r[0] + sin(beta) * r[1])));
        double z = r[2];
        u[i] = u[nw + i] = u[2 * nw + i] = 0; ud[i] = ud[nw + i] = ud[2 * nw + i] = 0; pDyn[i] = 0;
        if (z <= 0) {
            double S_, C_, P_;
            if (k[i] == 0.0) { S_ = 1.0; C_ = 99999.0; P_ = 99999.0; }
            else if (k[i] * h > 89.4) {
                S_ = exp(k[i] * z); C_ = exp(k[i] * z);
                P_ = exp(k[i] * z) + exp(-k[i] * (z + 2.0 * h));
            } else {
                S_ = sinh(k[i] * (z + h)) / sinh(k[i] * h);
                C_ = cosh(k[i] * (z + h)) / sinh(k[i] * h);
                P_ = cosh(k[i] * (z + h)) / cosh(k[i] * h);
            }
            u[i]          = w[i] * zeta * C_ * cos(beta);
            u[nw + i]     = w[i] * zeta * C_ * sin(beta);
            u[2 * nw + i] = I * w[i] * zeta * S_;
            for (int c = 0; c < 3; c++) ud[c * nw + i] = I * w[i] * u[c * nw + i];
            pDyn[i] = rho * g * zeta * P_;
        }
    }
}

/* helpers.py:468-483 translateForce3to6DOF */
static void translate_force(const cplx *f, const double *r, cplx *out)
{
    out[0] = f[0]; out[1] = f[1]; out[2] = f[2];
    out[3] = r[1] * f[2] - r[2] * f[1];
    out[4] = r[2] * f[0] - r[0] * f[2];
    out[5] = r[0] * f[1] - r[1] * f[0];
}

/* helpers.py:428-437 getH */
static void get_H(const double *r, double H[3][3])
{
    H[0][0] = 0;     H[0][1] = r[2];  H[0][2] = -r[1];
    H[1][0] = -r[2]; H[1][1] = 0;     H[1][2] = r[0];
    H[2][0] = r[1];  H[2][1] = -r[0]; H[2][2] = 0;
}

/* helpers.py:537-560 translateMatrix3to6DOF */
static void translate_matrix(const double Min[3][3], const double *r, double Mout[6][6])
{
    double H[3][3], MH[3][3], HM[3][3];
    get_H(r, H);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double a = 0, b = 0;
        for (int l = 0; l < 3; l++) { a += Min[i][l] * H[l][j]; b += H[i][l] * Min[l][j]; }
        MH[i][j] = a; HM[i][j] = b;
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        Mout[i][j] = Min[i][j];
        Mout[i][3 + j] = MH[i][j];
        Mout[3 + j][i] = MH[i][j];
        double a = 0;
        for (int l = 0; l < 3; l++) a += HM[i][l] * H[j][l];   /* (H m) H^T */
        Mout[3 + i][3 + j] = a;
    }
}

/* node.T for a rigid link from the reference node to a point offset d (raft_node.py:262-290):
 * T = [[I, H(d)],[0, I]] with H = getH(d). */
static void node_T(const double *d, double T[6][6])
{
    double H[3][3];
    get_H(d, H);
    memset(T, 0, 36 * sizeof(double));
    for (int i = 0; i < 6; i++) T[i][i] = 1.0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i][3 + j] = H[i][j];
}

/* exported wrappers so the helper-level known answers of the reference's tests/test_helpers.py
 * (getKinematics :26-38, translateForce3to6DOF :88-94, translateMatrix3to6DOF :123-136) can be
 * checked against this file's building blocks */
void ro_translate_force(const cplx *f, const double *r, cplx *out) { translate_force(f, r, out); }
void ro_translate_matrix(const double *Min, const double *r, double *Mout)
{
    translate_matrix((const double (*)[3])Min, r, (double (*)[6])Mout);
}
/* helpers.py:149-184 getKinematics(r, Xi, ws): Xi [6][nw] -> dr, v, a [3][nw] */
void ro_get_kinematics(const double *r, const cplx *Xi, const double *ws, int nw, cplx *dr, cplx *v, cplx *a)
{
    for (int i = 0; i < nw; i++) {
        const cplx th0 = Xi[3 * nw + i], th1 = Xi[4 * nw + i], th2 = Xi[5 * nw + i];
        dr[0 * nw + i] = Xi[0 * nw + i] + (-th2 * r[1] + th1 * r[2]);
        dr[1 * nw + i] = Xi[1 * nw + i] + ( th2 * r[0] - th0 * r[2]);
        dr[2 * nw + i] = Xi[2 * nw + i] + (-th1 * r[0] + th0 * r[1]);
        for (int c = 0; c < 3; c++) { v[c * nw + i] = I * ws[i] * dr[c * nw + i]; a[c * nw + i] = I * ws[i] * v[c * nw + i]; }
    }
}

/* ---- excitation -------------------------------------------------------------------- */

/* raft_fowt.py:1796-1849 BEM excitation for one wave train: F_BEM[6][nw] */
static void bem_excitation(const ro_design *d, const double *zeta, double beta_deg, cplx *F_BEM)
{
    int nw = d->nw, nhs = d->n_bem_head;
    double beta_rad = beta_deg * (M_PI / 180.0);         /* deg2rad, :1755 */
    for (int i = 0; i < 6 * nw; i++) F_BEM[i] = 0;
    if (!d->X_BEM || nhs <= 0) return;
    double beta = fmod(beta_rad * (180.0 / M_PI) - d->heading_adjust, 360.0);
    if (beta < 0) beta += 360.0;                          /* python % is non-negative */
    const double *hd = d->bem_headings;
    int i1 = 0, i2 = 0; double f2 = 0;
    if (beta <= hd[0]) {
        double hlast = hd[nhs - 1] - 360.0;
        i1 = nhs - 1; i2 = 0; f2 = (beta - hlast) / (hd[0] - hlast);
    } else if (beta >= hd[nhs - 1]) {
        double hfirst = hd[0] + 360.0;
        i1 = nhs - 1; i2 = 0; f2 = (beta - hd[nhs - 1]) / (hfirst - hd[nhs - 1]);
    } else {
        for (int i = 0; i < nhs - 1; i++) if (hd[i + 1] > beta) {
            i1 = i; i2 = i + 1; f2 = (beta - hd[i]) / (hd[i + 1] - hd[i]); break;
        }
    }
    double f1 = 1.0 - f2, sb = sin(beta_rad), cb = cos(beta_rad);
    for (int iw = 0; iw < nw; iw++) {
        cplx Xp[6], X[6];
        for (int j = 0; j < 6; j++)
            Xp[j] = d->X_BEM[((size_t)i1 * 6 + j) * nw + iw] * f1 + d->X_BEM[((size_t)i2 * 6 + j) * nw + iw] * f2;
        X[0] = Xp[0] * cb - Xp[1] * sb;  X[1] = Xp[0] * sb + Xp[1] * cb;  X[2] = Xp[2];
        X[3] = Xp[3] * cb - Xp[4] * sb;  X[4] = Xp[3] * sb + Xp[4] * cb;  X[5] = Xp[5];
        cplx ph = cexp(-I * d->k[iw] * (d->x_ref * cos(beta_rad) + d->y_ref * sin(beta_rad)));
        for (int j = 0; j < 6; j++) F_BEM[j * nw + iw] = X[j] * zeta[iw] * ph;
    }
}

/* raft_member.py:1940-1992 (+ raft_fowt.py:1854-1857, 1888).
 * u_out [Ns][3][nw] wave velocities kept for the drag passes.  F_iner [6][nw]. */
static void hydro_excitation(const ro_design *d, const double *zeta, double beta, cplx *u_out, cplx *F_iner)
{
    int nw = d->nw, Ns = d->n_nodes, Nm = d->n_members;
    cplx *ud = malloc(sizeof(cplx) * 3 * nw), *pD = malloc(sizeof(cplx) * nw);
    cplx *Fm = malloc(sizeof(cplx) * 6 * nw);
    for (int i = 0; i < 6 * nw; i++) F_iner[i] = 0;
    for (int m = 0; m < Nm; m++) {
        const double *q = d->mem_q + 3 * m, *rn = d->mem_rA + 3 * m;  /* member structural node */
        for (int i = 0; i < 6 * nw; i++) Fm[i] = 0;
        for (int il = 0; il < Ns; il++) {
            if (d->node_mem[il] != m) continue;
            const double *r = d->node_r + 3 * il;
            cplx *u = u_out + (size_t)il * 3 * nw;
            ro_wave_kin(zeta, beta, d->w, d->k, d->depth, r, nw, 1025.0, 9.81, u, ud, pD); /* defaults, fowt:1857 */
            const double *Im = d->node_Imat + 9 * il;
            double rr[3] = { r[0] - rn[0], r[1] - rn[1], r[2] - rn[2] };
            for (int i = 0; i < nw; i++) {
                cplx f[3], f6[6];
                for (int a = 0; a < 3; a++) {
                    if (d->node_Imat_w) {                             /* member:1984-1985 Imat_MCF[il,:,:,i] */
                        const cplx *Iw = d->node_Imat_w + (size_t)il * 9 * nw;
                        f[a] = Iw[(3 * a) * nw + i] * ud[i] + Iw[(3 * a + 1) * nw + i] * ud[nw + i]
                             + Iw[(3 * a + 2) * nw + i] * ud[2 * nw + i] + pD[i] * d->node_a_i[il] * q[a];
                    } else
                    f[a] = Im[3 * a] * ud[i] + Im[3 * a + 1] * ud[nw + i] + Im[3 * a + 2] * ud[2 * nw + i]
                         + pD[i] * d->node_a_i[il] * q[a];            /* member:1988 */
                }
                translate_force(f, rr, f6);
                for (int a = 0; a < 6; a++) Fm[a * nw + i] += f6[a];
            }
        }
        /* T^T reduction (fowt:1888) for the rigid link member node -> reference node */
        double dd[3] = { rn[0] - d->prp[0], rn[1] - d->prp[1], rn[2] - d->prp[2] }, T[6][6];
        node_T(dd, T);
        for (int i = 0; i < nw; i++)
            for (int a = 0; a < 6; a++) {
                cplx s = 0;
                for (int b = 0; b < 6; b++) s += T[b][a] * Fm[b * nw + i];
                F_iner[a * nw + i] += s;
            }
    }
    free(ud); free(pD); free(Fm);
}

/* raft_member.py:1995-2126 + raft_fowt.py:1891-1936.  One drag-linearisation pass.
 * Xi [6][nw] in; Bmat [Ns][3][3], B_drag [6][6], F_drag [6][nw] out. */
static void hydro_linearization(const ro_design *d, const cplx *u_all, const cplx *Xi,
                                double *Bmat_all, double B_drag[6][6], cplx *F_drag)
{
    int nw = d->nw, Ns = d->n_nodes, Nm = d->n_members;
    cplx *Xin = malloc(sizeof(cplx) * 6 * nw), *Fm = malloc(sizeof(cplx) * 6 * nw);
    cplx *vrel = malloc(sizeof(cplx) * 3 * nw);
    memset(B_drag, 0, 36 * sizeof(double));
    for (int i = 0; i < 6 * nw; i++) F_drag[i] = 0;
    for (int m = 0; m < Nm; m++) {
        const double *q = d->mem_q + 3 * m, *p1 = d->mem_p1 + 3 * m, *p2 = d->mem_p2 + 3 * m;
        const double *rn = d->mem_rA + 3 * m;
        int circ = d->mem_circ[m];
        double dd[3] = { rn[0] - d->prp[0], rn[1] - d->prp[1], rn[2] - d->prp[2] }, T[6][6];
        node_T(dd, T);
        for (int i = 0; i < nw; i++)                      /* Xi_nodes = node.T @ Xi, fowt:1921 */
            for (int a = 0; a < 6; a++) {
                cplx s = 0;
                for (int b = 0; b < 6; b++) s += T[a][b] * Xi[b * nw + i];
                Xin[a * nw + i] = s;
            }
        double Bm[6][6]; memset(Bm, 0, sizeof(Bm));
        for (int i = 0; i < 6 * nw; i++) Fm[i] = 0;
        for (int il = 0; il < Ns; il++) {
            if (d->node_mem[il] != m) continue;
            const double *r = d->node_r + 3 * il;
            const cplx *u = u_all + (size_t)il * 3 * nw;
            double rr[3] = { r[0] - rn[0], r[1] - rn[1], r[2] - rn[2] };
            double sq = 0, sp = 0, sp1 = 0, sp2 = 0;
            for (int i = 0; i < nw; i++) {
                /* getKinematics helpers.py:178-181 with SmallRotate :396-408 */
                const cplx th0 = Xin[3 * nw + i], th1 = Xin[4 * nw + i], th2 = Xin[5 * nw + i];
                cplx dr[3], v[3];
                dr[0] = Xin[0 * nw + i] + (-th2 * rr[1] + th1 * rr[2]);
                dr[1] = Xin[1 * nw + i] + ( th2 * rr[0] - th0 * rr[2]);
                dr[2] = Xin[2 * nw + i] + (-th1 * rr[0] + th0 * rr[1]);
                for (int a = 0; a < 3; a++) { v[a] = I * d->w[i] * dr[a]; vrel[a * nw + i] = u[a * nw + i] - v[a]; }
                cplx aq = 0, a1 = 0, a2 = 0;
                for (int a = 0; a < 3; a++) { aq += vrel[a * nw + i] * q[a]; a1 += vrel[a * nw + i] * p1[a]; a2 += vrel[a * nw + i] * p2[a]; }
                for (int a = 0; a < 3; a++) {              /* member:2078-2081 + getRMS helpers:684 */
                    cplx vq = aq * q[a], vp = vrel[a * nw + i] - vq, v1 = a1 * p1[a], v2 = a2 * p2[a];
                    double t;
                    t = cabs(vq); sq += t * t;  t = cabs(vp); sp += t * t;
                    t = cabs(v1); sp1 += t * t; t = cabs(v2); sp2 += t * t;
                }
            }
            double vRMS_q = sqrt(0.5 * sq), vRMS_p1, vRMS_p2;
            if (circ) { vRMS_p1 = sqrt(0.5 * sp); vRMS_p2 = vRMS_p1; }
            else { vRMS_p1 = sqrt(0.5 * sp1); vRMS_p2 = sqrt(0.5 * sp2); }
            double c = sqrt(8.0 / M_PI), rho = d->rho;
            double Bq  = c * vRMS_q  * 0.5 * rho * d->a_q[il]  * d->Cd_q[il];
            double Bp1 = c * vRMS_p1 * 0.5 * rho * d->a_p1[il] * d->Cd_p1[il];
            double Bp2 = c * vRMS_p2 * 0.5 * rho * d->a_p2[il] * d->Cd_p2[il];
            double Be  = c * vRMS_q  * 0.5 * rho * d->a_End[il] * d->Cd_End[il];
            double Bmat[3][3], B6[6][6];
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
                Bmat[a][b] = (Bq * (q[a] * q[b]) + Bp1 * (p1[a] * p1[b]) + Bp2 * (p2[a] * p2[b])) + Be * (q[a] * q[b]);
                Bmat_all[9 * il + 3 * a + b] = Bmat[a][b];
            }
            translate_matrix(Bmat, rr, B6);
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Bm[a][b] += B6[a][b];
            for (int i = 0; i < nw; i++) {                 /* member:2122-2124 */
                cplx f[3], f6[6];
                for (int a = 0; a < 3; a++)
                    f[a] = Bmat[a][0] * u[i] + Bmat[a][1] * u[nw + i] + Bmat[a][2] * u[2 * nw + i];
                translate_force(f, rr, f6);
                for (int a = 0; a < 6; a++) Fm[a * nw + i] += f6[a];
            }
        }
        /* B_drag += T^T Bm T ; F_drag += T^T Fm   (fowt:1927-1929) */
        double TB[6][6];
        for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
            double s = 0; for (int l = 0; l < 6; l++) s += T[l][a] * Bm[l][b]; TB[a][b] = s;
        }
        for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
            double s = 0; for (int l = 0; l < 6; l++) s += TB[a][l] * T[l][b]; B_drag[a][b] += s;
        }
        for (int i = 0; i < nw; i++)
            for (int a = 0; a < 6; a++) {
                cplx s = 0;
                for (int b = 0; b < 6; b++) s += T[b][a] * Fm[b * nw + i];
                F_drag[a * nw + i] += s;
            }
    }
    free(Xin); free(Fm); free(vrel);
}

/* ---- dense complex linear algebra (stands in for LAPACK zgesv / zgetrf+zgetri, which the
 * reference reaches through numpy.linalg.solve / inv at raft_model.py:1089, 1191) ---------- */
static double cabs1(cplx z) { return fabs(creal(z)) + fabs(cimag(z)); }

/* LU with partial pivoting (izamax metric |re|+|im|); returns 0 or k+1 for a zero pivot */
static int lu_factor(int n, cplx *A, int *piv)
{
    for (int k = 0; k < n; k++) {
        int p = k; double best = cabs1(A[k * n + k]);
        for (int i = k + 1; i < n; i++) { double t = cabs1(A[i * n + k]); if (t > best) { best = t; p = i; } }
        piv[k] = p;
        if (best == 0.0) return k + 1;
        if (p != k) for (int j = 0; j < n; j++) { cplx t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
        cplx rinv = 1.0 / A[k * n + k];
        for (int i = k + 1; i < n; i++) {
            A[i * n + k] *= rinv;
            cplx l = A[i * n + k];
            for (int j = k + 1; j < n; j++) A[i * n + j] -= l * A[k * n + j];
        }
    }
    return 0;
}

static void lu_solve(int n, const cplx *A, const int *piv, cplx *b)
{
    for (int k = 0; k < n; k++) { if (piv[k] != k) { cplx t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; } }
    for (int i = 1; i < n; i++) { cplx s = b[i]; for (int j = 0; j < i; j++) s -= A[i * n + j] * b[j]; b[i] = s; }
    for (int i = n - 1; i >= 0; i--) { cplx s = b[i]; for (int j = i + 1; j < n; j++) s -= A[i * n + j] * b[j]; b[i] = s / A[i * n + i]; }
}

/* solve A x = b for n<=NMAX, A row-major (destroyed) */
int ro_zgesv(int n, cplx *A, cplx *b)
{
    int *piv = malloc(sizeof(int) * n);
    int info = lu_factor(n, A, piv);
    if (!info) lu_solve(n, A, piv, b);
    free(piv);
    return info;
}

/* explicit inverse (numpy.linalg.inv) via LU and n unit right-hand sides */
int ro_zinv(int n, cplx *A, cplx *Ainv)
{
    int *piv = malloc(sizeof(int) * n);
    cplx *col = malloc(sizeof(cplx) * n);
    int info = lu_factor(n, A, piv);
    if (!info)
        for (int j = 0; j < n; j++) {
            for (int i = 0; i < n; i++) col[i] = (i == j);
            lu_solve(n, A, piv, col);
            for (int i = 0; i < n; i++) Ainv[i * n + j] = col[i];
        }
    free(piv); free(col);
    return info;
}

/* ---- public entry points --------------------------------------------------------------- */

/* FOWT.calcHydroForce_2ndOrd(beta, S0), interpMode 'qtf' (raft_fowt.py:2158-2253): difference-frequency force
 * amplitudes from the QTF table.  S0 [nw] wave spectrum, beta [rad] -> f_mean [6], f [6][nw] (real amplitudes,
 * already shifted by one bin, :2244-2245).
 *   :2178-2187  heading: single table as is, else scipy interp1d(kind linear) along the heading axis with the first /
 *               last table as fill value outside the range (slope form y = (y_hi-y_lo)/(x_hi-x_lo)*(x-x_lo)+y_lo)
 *   :2221-2229  RegularGridInterpolator(linear, bounds_error False, fill_value 0) of Re and Im onto (w_i, w_j):
 *               v00(1-ti)(1-tj) + v01(1-ti)tj + v10 ti(1-tj) + v11 ti tj, zero outside the table
 *   :2231-2236  f[imu] = 4 sqrt(sum_i S0[i] S0[i+imu] |Q(w_i, w_{i+imu})|^2) dw, imu = 1..nw-1
 *   :2239       f_mean = 2 sum_i S0[i] Re Q(w_i, w_i) dw */
void ro_hydro_force_2nd(const ro_design *d, double beta, const double *S0, double *f_mean, double *f)
{
    int nw = d->nw, n2 = d->n_qtf_w, nh = d->n_qtf_head;
    cplx *qb = malloc(sizeof(cplx) * (size_t)n2 * n2 * 6);
    for (size_t e = 0; e < (size_t)n2 * n2; e++)
        for (int a = 0; a < 6; a++) {
            const cplx *row = d->qtf + (e * nh) * 6;
            if (nh == 1) qb[e * 6 + a] = row[a];
            else if (beta < d->qtf_heads[0]) qb[e * 6 + a] = row[a];
            else if (beta > d->qtf_heads[nh - 1]) qb[e * 6 + a] = row[(size_t)(nh - 1) * 6 + a];
            else {
                int idx = 0;
                while (idx < nh && d->qtf_heads[idx] < beta) idx++;          /* searchsorted, side left */
                if (idx < 1) idx = 1;
                if (idx > nh - 1) idx = nh - 1;
                double xl = d->qtf_heads[idx - 1], xh = d->qtf_heads[idx];
                cplx yl = row[(size_t)(idx - 1) * 6 + a], yh = row[(size_t)idx * 6 + a];
                double re = (creal(yh) - creal(yl)) / (xh - xl) * (beta - xl) + creal(yl);
                double im = (cimag(yh) - cimag(yl)) / (xh - xl) * (beta - xl) + cimag(yl);
                qb[e * 6 + a] = re + I * im;
            }
        }
    int *cell = malloc(sizeof(int) * nw);
    double *t = malloc(sizeof(double) * nw);
    for (int i = 0; i < nw; i++) {
        double x = d->w[i];
        if (x < d->qtf_w[0] || x > d->qtf_w[n2 - 1]) { cell[i] = -1; t[i] = 0; continue; }
        int c = 0;
        while (c < n2 - 2 && d->qtf_w[c + 1] <= x) c++;
        cell[i] = c;
        t[i] = (x - d->qtf_w[c]) / (d->qtf_w[c + 1] - d->qtf_w[c]);
    }
#define QTF_AT(i_, j_, a_, re_, im_) do {                                                            \
        re_ = 0; im_ = 0;                                                                            \
        if (cell[i_] >= 0 && cell[j_] >= 0) {                                                        \
            int ci = cell[i_], cj = cell[j_];                                                        \
            double ti = t[i_], tj = t[j_];                                                           \
            cplx v00 = qb[((size_t)ci * n2 + cj) * 6 + a_], v01 = qb[((size_t)ci * n2 + cj + 1) * 6 + a_]; \
            cplx v10 = qb[((size_t)(ci + 1) * n2 + cj) * 6 + a_], v11 = qb[((size_t)(ci + 1) * n2 + cj + 1) * 6 + a_]; \
            double w00 = (1 - ti) * (1 - tj), w01 = (1 - ti) * tj, w10 = ti * (1 - tj), w11 = ti * tj; \
            re_ = ((creal(v00) * w00 + creal(v01) * w01) + creal(v10) * w10) + creal(v11) * w11;     \
            im_ = ((cimag(v00) * w00 + cimag(v01) * w01) + cimag(v10) * w10) + cimag(v11) * w11;     \
        } } while (0)
    for (int a = 0; a < 6; a++) {
        double *fa = f + (size_t)a * nw;
        fa[0] = 0;
        for (int imu = 1; imu < nw; imu++) {
            double s = 0;
            for (int i = 0; i < nw - imu; i++) {
                double re, im;
                QTF_AT(i, i + imu, a, re, im);
                s += S0[i] * S0[i + imu] * (re * re + im * im);
            }
            fa[imu] = 4 * sqrt(s) * d->dw;
        }
        double sm = 0;
        for (int i = 0; i < nw; i++) {
            double re, im;
            QTF_AT(i, i, a, re, im);
            (void)im;
            sm += S0[i] * re;
        }
        f_mean[a] = 2 * sm * d->dw;
        for (int i = 0; i < nw - 1; i++) fa[i] = fa[i + 1];                   /* :2244 */
        fa[nw - 1] = 0;                                                       /* :2245 */
    }
#undef QTF_AT
    free(qb); free(cell); free(t);
}

/* second-order force of one wave train as the solver adds it (raft_model.py:1035-1038, :1210-1211): zero unless
 * the design carries a QTF table.  F2 [6][nw] real, F2_mean [6] (may be NULL). */
static int second_order_force(const ro_design *d, int spec, double Hs, double Tp, double gamma, double beta_deg,
                              double *F2, double *F2_mean)
{
    int nw = d->nw;
    double fm[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6 * nw; i++) F2[i] = 0;
    if (d->qtf) {
        double *S = malloc(sizeof(double) * nw), *z = malloc(sizeof(double) * nw);
        int rc = ro_sea_state(d->w, nw, d->dw, spec, Hs, Tp, gamma, S, z);
        if (!rc) ro_hydro_force_2nd(d, beta_deg * (M_PI / 180.0), S, fm, F2);
        free(S); free(z);
        if (rc) return rc;
    }
    if (F2_mean) memcpy(F2_mean, fm, sizeof(fm));
    return 0;
}

/* FOWT.calcHydroExcitation for one single-train case: zeta[nw], F_BEM[6][nw], F_iner[6][nw], u[Ns][3][nw] */
int ro_calc_hydro_excitation(const ro_design *d, int spec, double Hs, double Tp, double gamma, double beta_deg,
                             double *zeta, cplx *F_BEM, cplx *F_iner, cplx *u)
{
    double *S = malloc(sizeof(double) * d->nw);
    int rc = ro_sea_state(d->w, d->nw, d->dw, spec, Hs, Tp, gamma, S, zeta);
    free(S);
    if (rc) return rc;
    double beta = beta_deg * (M_PI / 180.0);
    bem_excitation(d, zeta, beta_deg, F_BEM);
    hydro_excitation(d, zeta, beta, u, F_iner);
    return 0;
}

/* FOWT.calcHydroLinearization(Xi) given wave velocities u from ro_calc_hydro_excitation */
void ro_calc_hydro_linearization(const ro_design *d, const cplx *u, const cplx *Xi,
                                 double *Bmat, double *B_drag, cplx *F_drag)
{
    hydro_linearization(d, u, Xi, Bmat, (double (*)[6])B_drag, F_drag);
}

/* Model.solveDynamics for one FOWT and one single-train case (raft_model.py:994-1156, 1189-1216).
 * Xi_out [6][nw].  status[0] = passes executed, status[1] = converged flag, status[2] = NaN flag.
 * If Z_out != NULL it receives the last impedance matrices [nw][6][6]; B_drag_out [6][6] optional. */
int ro_solve_dynamics(const ro_design *d, int spec, double Hs, double Tp, double gamma, double beta_deg,
                      int nIter, double tol, double XiStart, cplx *Xi_out, int *status,
                      cplx *Z_out, double *B_drag_out)
{
    int nw = d->nw, Ns = d->n_nodes;
    double *zeta = malloc(sizeof(double) * nw), *Bmat = malloc(sizeof(double) * 9 * (Ns > 0 ? Ns : 1));
    cplx *F_BEM = malloc(sizeof(cplx) * 6 * nw), *F_iner = malloc(sizeof(cplx) * 6 * nw);
    cplx *u = malloc(sizeof(cplx) * (size_t)(Ns > 0 ? Ns : 1) * 3 * nw);
    cplx *XiLast = malloc(sizeof(cplx) * 6 * nw), *Xi = malloc(sizeof(cplx) * 6 * nw);
    cplx *F_drag = malloc(sizeof(cplx) * 6 * nw), *Z = malloc(sizeof(cplx) * 36 * (size_t)nw);
    double B_drag[6][6];
    double *F2 = malloc(sizeof(double) * 6 * nw);
    int rc = ro_calc_hydro_excitation(d, spec, Hs, Tp, gamma, beta_deg, zeta, F_BEM, F_iner, u);
    status[0] = status[1] = status[2] = 0;
    if (rc) goto done;
    second_order_force(d, spec, Hs, Tp, gamma, beta_deg, F2, NULL);             /* :1035-1038 */
    for (int i = 0; i < 6 * nw; i++) XiLast[i] = XiStart;
    int passes = 0, conv = 0;
    for (int iiter = 0; iiter < nIter + 1; iiter++) {                        /* :977, :1052 */
        hydro_linearization(d, u, XiLast, Bmat, B_drag, F_drag);             /* :1063-1064 */
        passes++;
        int nan = 0;
        for (int ii = 0; ii < nw; ii++) {
            cplx A[36], b[6];
            double wv = d->w[ii];
            for (int a = 0; a < 6; a++) {
                for (int c = 0; c < 6; c++) {
                    double M = d->M0[6 * a + c], B = d->B0[6 * a + c];
                    if (d->A_w) M += d->A_w[(size_t)(6 * a + c) * nw + ii];
                    if (d->B_w) B += d->B_w[(size_t)(6 * a + c) * nw + ii];
                    B += B_drag[a][c];
                    A[6 * a + c] = -wv * wv * M + I * wv * B + d->C0[6 * a + c];   /* :1086 */
                }
                b[a] = ((F_BEM[a * nw + ii] + F_iner[a * nw + ii]) + F2[a * nw + ii]) + F_drag[a * nw + ii]; /* :1048,:1081 */
            }
            memcpy(Z + (size_t)ii * 36, A, sizeof(A));
            ro_zgesv(6, A, b);                                               /* :1089 */
            for (int a = 0; a < 6; a++) { Xi[a * nw + ii] = b[a]; if (isnan(creal(b[a])) || isnan(cimag(b[a]))) nan = 1; }
        }
        if (nan) { status[2] = 1; break; }                                   /* :1098 */
        int all = 1;                                                          /* :1103-1104 */
        for (int i = 0; i < 6 * nw; i++) {
            double tc = cabs(Xi[i] - XiLast[i]) / (cabs(Xi[i]) + tol);
            if (!(tc < tol)) { all = 0; break; }
        }
        if (all) { conv = 1; break; }
        for (int i = 0; i < 6 * nw; i++) XiLast[i] = 0.2 * XiLast[i] + 0.8 * Xi[i];   /* :1133 */
    }
    status[0] = passes; status[1] = conv;
    /* system response with the explicit inverse of the last Z (raft_model.py:1189-1216);
       F_wave = F_BEM + F_iner + F_drag(Bmat_last, u[0]) which equals the last pass's F_drag */
    for (int ii = 0; ii < nw; ii++) {
        cplx A[36], Ai[36];
        memcpy(A, Z + (size_t)ii * 36, sizeof(A));
        if (ro_zinv(6, A, Ai)) { for (int a = 0; a < 6; a++) Xi_out[a * nw + ii] = NAN; continue; }
        for (int a = 0; a < 6; a++) {
            cplx s = 0;
            for (int c = 0; c < 6; c++)
                s += Ai[6 * a + c] * (((F_BEM[c * nw + ii] + F_iner[c * nw + ii]) + F_drag[c * nw + ii]) + F2[c * nw + ii]);  /* :1212 */
            Xi_out[a * nw + ii] = s;
        }
    }
    if (Z_out) memcpy(Z_out, Z, sizeof(cplx) * 36 * (size_t)nw);
    if (B_drag_out) memcpy(B_drag_out, B_drag, sizeof(B_drag));
done:
    free(zeta); free(Bmat); free(F_BEM); free(F_iner); free(u); free(XiLast); free(Xi); free(F_drag); free(Z); free(F2);
    return rc;
}

/* raft_member.py:2128-2152 + raft_fowt.py:1940-1957: drag excitation of one wave train with the Bmat left
 * behind by the last calcHydroLinearization.  u [Ns][3][nw], Bmat [Ns][3][3] -> F_drag [6][nw]. */
static void drag_excitation(const ro_design *d, const cplx *u_all, const double *Bmat_all, cplx *F_drag)
{
    int nw = d->nw, Ns = d->n_nodes, Nm = d->n_members;
    cplx *Fm = malloc(sizeof(cplx) * 6 * nw);
    for (int i = 0; i < 6 * nw; i++) F_drag[i] = 0;
    for (int m = 0; m < Nm; m++) {
        const double *rn = d->mem_rA + 3 * m;
        double dd[3] = { rn[0] - d->prp[0], rn[1] - d->prp[1], rn[2] - d->prp[2] }, T[6][6];
        node_T(dd, T);
        for (int i = 0; i < 6 * nw; i++) Fm[i] = 0;
        for (int il = 0; il < Ns; il++) {
            if (d->node_mem[il] != m) continue;
            const double *r = d->node_r + 3 * il, *B = Bmat_all + 9 * il;
            const cplx *u = u_all + (size_t)il * 3 * nw;
            double rr[3] = { r[0] - rn[0], r[1] - rn[1], r[2] - rn[2] };
            for (int i = 0; i < nw; i++) {
                cplx f[3], f6[6];
                for (int a = 0; a < 3; a++) f[a] = B[3 * a] * u[i] + B[3 * a + 1] * u[nw + i] + B[3 * a + 2] * u[2 * nw + i];
                translate_force(f, rr, f6);
                for (int a = 0; a < 6; a++) Fm[a * nw + i] += f6[a];
            }
        }
        for (int i = 0; i < nw; i++)
            for (int a = 0; a < 6; a++) {
                cplx s = 0;
                for (int b = 0; b < 6; b++) s += T[b][a] * Fm[b * nw + i];
                F_drag[a * nw + i] += s;
            }
    }
    free(Fm);
}

/* Model.solveDynamics for a case with nH wave trains (raft_fowt.py:1742-1752 lists; raft_model.py:1200-1236):
 * the drag linearisation iterates on train 0 only (raft_fowt.py:1910); every train's response is then
 * inv(Z_last) (F_BEM[ih] + F_iner[ih] + F_drag(Bmat_last, u[ih])).  Xi_out [nH][6][nw]. */
int ro_solve_dynamics_trains(const ro_design *d, int nH, const int *spec, const double *Hs, const double *Tp,
                             const double *gamma, const double *beta_deg, int nIter, double tol, double XiStart,
                             cplx *Xi_out, int *status)
{
    int nw = d->nw, Ns = d->n_nodes > 0 ? d->n_nodes : 1;
    cplx *Z = malloc(sizeof(cplx) * 36 * (size_t)nw);
    double Bd[36];
    int rc = ro_solve_dynamics(d, spec[0], Hs[0], Tp[0], gamma[0], beta_deg[0], nIter, tol, XiStart, Xi_out, status, Z, Bd);
    if (rc || nH == 1) { free(Z); return rc; }
    /* recover Bmat of the last pass: one more linearisation with the XiLast that produced Z is not available here, so
       redo the loop bookkeeping explicitly: run the loop again keeping Bmat (cheap; this is test infrastructure) */
    double *zeta = malloc(sizeof(double) * nw), *Bmat = malloc(sizeof(double) * 9 * Ns);
    cplx *F_BEM = malloc(sizeof(cplx) * 6 * nw), *F_iner = malloc(sizeof(cplx) * 6 * nw), *u = malloc(sizeof(cplx) * (size_t)Ns * 3 * nw);
    cplx *XiLast = malloc(sizeof(cplx) * 6 * nw), *Xi = malloc(sizeof(cplx) * 6 * nw), *F_drag = malloc(sizeof(cplx) * 6 * nw);
    double B_drag[6][6];
    double *F2 = malloc(sizeof(double) * 6 * nw);
    ro_calc_hydro_excitation(d, spec[0], Hs[0], Tp[0], gamma[0], beta_deg[0], zeta, F_BEM, F_iner, u);
    second_order_force(d, spec[0], Hs[0], Tp[0], gamma[0], beta_deg[0], F2, NULL);
    for (int i = 0; i < 6 * nw; i++) XiLast[i] = XiStart;
    for (int iiter = 0; iiter < nIter + 1; iiter++) {
        hydro_linearization(d, u, XiLast, Bmat, B_drag, F_drag);
        for (int ii = 0; ii < nw; ii++) {
            cplx A[36], b[6];
            double wv = d->w[ii];
            for (int a = 0; a < 6; a++) {
                for (int c = 0; c < 6; c++) {
                    double M = d->M0[6 * a + c], B = d->B0[6 * a + c];
                    if (d->A_w) M += d->A_w[(size_t)(6 * a + c) * nw + ii];
                    if (d->B_w) B += d->B_w[(size_t)(6 * a + c) * nw + ii];
                    B += B_drag[a][c];
                    A[6 * a + c] = -wv * wv * M + I * wv * B + d->C0[6 * a + c];
                }
                b[a] = ((F_BEM[a * nw + ii] + F_iner[a * nw + ii]) + F2[a * nw + ii]) + F_drag[a * nw + ii];
            }
            ro_zgesv(6, A, b);
            for (int a = 0; a < 6; a++) Xi[a * nw + ii] = b[a];
        }
        int all = 1;
        for (int i = 0; i < 6 * nw; i++) { double tc = cabs(Xi[i] - XiLast[i]) / (cabs(Xi[i]) + tol); if (!(tc < tol)) { all = 0; break; } }
        if (all) break;
        for (int i = 0; i < 6 * nw; i++) XiLast[i] = 0.2 * XiLast[i] + 0.8 * Xi[i];
    }
    for (int ih = 1; ih < nH; ih++) {
        rc = ro_calc_hydro_excitation(d, spec[ih], Hs[ih], Tp[ih], gamma[ih], beta_deg[ih], zeta, F_BEM, F_iner, u);
        if (rc) break;
        second_order_force(d, spec[ih], Hs[ih], Tp[ih], gamma[ih], beta_deg[ih], F2, NULL);   /* :1210-1211 */
        drag_excitation(d, u, Bmat, F_drag);
        cplx *Xo = Xi_out + (size_t)ih * 6 * nw;
        for (int ii = 0; ii < nw; ii++) {
            cplx A[36], Ai[36];
            memcpy(A, Z + (size_t)ii * 36, sizeof(A));
            if (ro_zinv(6, A, Ai)) { for (int a = 0; a < 6; a++) Xo[a * nw + ii] = NAN; continue; }
            for (int a = 0; a < 6; a++) {
                cplx s = 0;
                for (int c = 0; c < 6; c++) s += Ai[6 * a + c] * (((F_BEM[c * nw + ii] + F_iner[c * nw + ii]) + F_drag[c * nw + ii]) + F2[c * nw + ii]);
                Xo[a * nw + ii] = s;
            }
        }
    }
    free(Z); free(zeta); free(Bmat); free(F_BEM); free(F_iner); free(u); free(XiLast); free(Xi); free(F_drag); free(F2);
    return rc;
}

/* Batched driver used for parity sweeps and the CPU baseline: nC cases of one design,
 * OpenMP over cases.  Xi_out [nC][6][nw], status [nC][3].  Returns the thread count used. */
int ro_solve_cases(const ro_design *d, int nC, const int *spec, const double *Hs, const double *Tp,
                   const double *gamma, const double *beta_deg, int nIter, double tol, double XiStart,
                   cplx *Xi_out, int *status, int nthreads)
{
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int c = 0; c < nC; c++)
        ro_solve_dynamics(d, spec[c], Hs[c], Tp[c], gamma[c], beta_deg[c], nIter, tol, XiStart,
                          Xi_out + (size_t)c * 6 * d->nw, status + 3 * c, NULL, NULL);
    return used;
}

/* Farm: dense n x n (n = 6N) system response per frequency, raft_model.py:1164-1216.
 * Z_sys [nw][n][n] (already assembled incl. array mooring), F [nw][n] -> Xi [nw][n] via inverse. */
int ro_system_response(int n, int nw, const cplx *Z_sys, const cplx *F, cplx *Xi)
{
    int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+:bad)
#endif
    for (int iw = 0; iw < nw; iw++) {
        cplx *A = malloc(sizeof(cplx) * n * n), *Ai = malloc(sizeof(cplx) * n * n);
        memcpy(A, Z_sys + (size_t)iw * n * n, sizeof(cplx) * n * n);
        if (ro_zinv(n, A, Ai)) bad++;
        else for (int a = 0; a < n; a++) {
            cplx s = 0;
            for (int c = 0; c < n; c++) s += Ai[a * n + c] * F[(size_t)iw * n + c];
            Xi[(size_t)iw * n + a] = s;
        }
        free(A); free(Ai);
    }
    return bad;
}
