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
 * from an external QTF table (potSecOrder 2; raft_fowt.py:2158-2253) or the slender-body QTF (potSecOrder 1;
 * raft_fowt.py:1988-2078, raft_member.py:1488-1792, helpers.py:239-373), single wave train for the latter
 * (BASELINE.json configs 1-4; SURVEY.md section 8a rows a1-a11 + 8f rows 3-4).
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
    /* slender-body QTF (potSecOrder 1): member tables + second-order grid, or qs == NULL */
    const struct ro_qtf_design_s *qs;
    int qs_nw;
    const double *qs_w, *qs_k;                     /* [qs_nw] w1_2nd, k1_2nd                              */
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


/* ======================================================================================================
 * Slender-body QTF (potSecOrder 1): FOWT.calcQTF_slenderBody (raft_fowt.py:1988-2078),
 * Member.calcQTF_slenderBody + correction_KAY (raft_member.py:1488-1792) and the second-order wave kinematics
 * helpers (helpers.py:239-373).  Loop order and the reference's quirks are kept:
 *   - getWaveKin_grad_u1 / grad_pres1st / pot2ndOrd apply deg2rad to a heading that is already in radians for the
 *     amplitude factors (cosBeta, sinBeta) while the phase of grad_u1 uses the heading itself (helpers.py:244-245, 262);
 *   - getWaveKin_axdivAcc removes the axial component of the node velocities IN PLACE (helpers.py:320-321, the
 *     arguments are views of nodeV), so every later use of nodeV in the pair loop sees the transverse part only,
 *     while nodeV_axial_rel was computed before from the full velocity (raft_member.py:1517).
 * ====================================================================================================== */
typedef struct ro_qtf_design_s {
    int n_nodes, n_members, n_seg;
    double depth, rho, g;
    const double *mem_q, *mem_p1, *mem_p2;   /* [Nm,3]                                                      */
    const int *mem_mcf;                      /* [Nm] 1: Kim & Yue correction applies (mem.MCF and rA_z * rB_z < 0) */
    const int *mem_wl;                       /* [Nm] 1: the member crosses the mean waterline                */
    const double *mem_r_int;                 /* [Nm,3] intersection with z = 0 (raft_member.py:1528)         */
    const double *mem_a_wl;                  /* [Nm] cross-section area at the waterline (:1660-1674)        */
    const double *mem_rwl;                   /* [Nm,3] KAY: waterline point from rA, rB (:1723)              */
    const double *mem_R_wl;                  /* [Nm] KAY: radius at z = 0 (:1725)                            */
    const int *node_mem;                     /* [Ns] submerged strip nodes (r_z < 0), grouped by member      */
    const double *node_r;                    /* [Ns,3]                                                       */
    const double *node_v_side, *node_Ca_p1, *node_Ca_p2, *node_Ca_End, *node_v_end, *node_a_i;   /* [Ns]   */
    const int *seg_mem;                      /* [n_seg] KAY integration segments (:1741-1760)                */
    const double *seg_z1, *seg_z2, *seg_R, *seg_rmid;  /* [n_seg], [n_seg], [n_seg], [n_seg,3]               */
    const double *M_struc;                   /* [6,6] fowt.M_struc (Pinkster IV term, raft_fowt.py:2044)     */
} ro_qtf_design;

typedef struct { cplx v[3]; } c3;

static c3 c3_zero(void) { c3 r; r.v[0] = r.v[1] = r.v[2] = 0; return r; }
static c3 c3_add(c3 a, c3 b) { c3 r; for (int i = 0; i < 3; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
static c3 c3_sub(c3 a, c3 b) { c3 r; for (int i = 0; i < 3; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
static c3 c3_scale(c3 a, cplx s) { c3 r; for (int i = 0; i < 3; i++) r.v[i] = a.v[i] * s; return r; }
static c3 c3_conj(c3 a) { c3 r; for (int i = 0; i < 3; i++) r.v[i] = conj(a.v[i]); return r; }
static cplx c3_dotr(c3 a, const double *d) { return a.v[0] * d[0] + a.v[1] * d[1] + a.v[2] * d[2]; }
static cplx c3_dot(c3 a, c3 b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }   /* np.dot: no conjugation */
static c3 c3_vecr(const double *d, cplx s) { c3 r; for (int i = 0; i < 3; i++) r.v[i] = d[i] * s; return r; }
static c3 m33_mul(cplx M[3][3], c3 x) { c3 r; for (int i = 0; i < 3; i++) r.v[i] = M[i][0] * x.v[0] + M[i][1] * x.v[1] + M[i][2] * x.v[2]; return r; }
static c3 m33c_mul(cplx M[3][3], c3 x) { c3 r; for (int i = 0; i < 3; i++) r.v[i] = conj(M[i][0]) * x.v[0] + conj(M[i][1]) * x.v[1] + conj(M[i][2]) * x.v[2]; return r; }
/* (a p1 p1' + b p2 p2') v */
static c3 proj_p(const double *p1, const double *p2, double a, double b, c3 v)
{
    return c3_add(c3_vecr(p1, a * c3_dotr(v, p1)), c3_vecr(p2, b * c3_dotr(v, p2)));
}
static c3 proj_q(const double *q, c3 v) { return c3_vecr(q, c3_dotr(v, q)); }

/* helpers.py:239-278 getWaveKin_grad_u1 (beta in radians, see the quirk note above) */
static void grad_u1(double w, double k, double beta, double h, const double *r, cplx G[3][3])
{
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) G[i][j] = 0;
    double z = r[2];
    double cosBeta = cos(beta * 0.017453292519943295), sinBeta = sin(beta * 0.017453292519943295);
    if (z <= 0 && k > 0) {
        double khz_xy, khz_z;
        if (k * h >= 10) { khz_xy = exp(k * z); khz_z = khz_xy; }
        else { khz_xy = cosh(k * (z + h)) / sinh(k * h); khz_z = sinh(k * (z + h)) / sinh(k * h); }
        cplx ph = cexp(-I * (k * (cos(beta) * r[0] + sin(beta) * r[1])));
        cplx aux = w * cosBeta * ph;
        G[0][0] = -I * aux * khz_xy * k * cosBeta;
        G[0][1] = -I * aux * khz_xy * k * sinBeta;
        G[0][2] = aux * k * khz_z;
        aux = w * sinBeta * ph;
        G[1][0] = G[0][1];
        G[1][1] = -I * aux * khz_xy * k * sinBeta;
        G[1][2] = aux * k * khz_z;
        aux = I * w * ph;
        G[2][0] = G[0][2];
        G[2][1] = G[0][1];
        G[2][2] = aux * k * khz_xy;
    }
}

/* helpers.py:285-308 getWaveKin_grad_pres1st */
static c3 grad_pres1st(double k, double beta, double h, const double *r, double rho, double g)
{
    c3 gr = c3_zero();
    double z = r[2];
    double cosBeta = cos(beta * 0.017453292519943295), sinBeta = sin(beta * 0.017453292519943295);
    if (z <= 0 && k > 0) {
        double khz_xy, khz_z;
        if (k * h >= 10) { khz_xy = exp(k * z); khz_z = khz_xy; }
        else { khz_xy = cosh(k * (z + h)) / cosh(k * h); khz_z = sinh(k * (z + h)) / cosh(k * h); }
        cplx ph = cexp(-I * (k * (cosBeta * r[0] + sinBeta * r[1])));
        gr.v[0] = rho * g * khz_xy * ph * (-I * k * cosBeta);
        gr.v[1] = rho * g * khz_xy * ph * (-I * k * sinBeta);
        gr.v[2] = rho * g * khz_z * ph * k;
    }
    return gr;
}

/* getWaveKin for one frequency and unit amplitude -> u, ud, pDyn (helpers.py:188-236) */
static void wave_kin1(double beta, double w, double k, double h, const double *r, double rho, double g, c3 *u, c3 *ud, cplx *pDyn)
{
    double one = 1.0;
    cplx uu[3], uud[3], p;
    ro_wave_kin(&one, beta, &w, &k, h, r, 1, rho, g, uu, uud, &p);
    for (int i = 0; i < 3; i++) { if (u) u->v[i] = uu[i]; if (ud) ud->v[i] = uud[i]; }
    if (pDyn) *pDyn = p;
}

/* helpers.py:311-334 getWaveKin_axdivAcc; vel1/vel2 are already the transverse node velocities */
static c3 axdiv_acc(double w1, double w2, double k1, double k2, double beta, double h, const double *r, c3 vel1, c3 vel2,
                    const double *q, double g)
{
    cplx G[3][3];
    c3 qv = c3_vecr(q, 1.0);
    grad_u1(w1, k1, beta, h, r, G);
    cplx dwdz1 = c3_dotr(m33_mul(G, qv), q);
    c3 u1, u2;
    wave_kin1(beta, w1, k1, h, r, 1025.0, g, &u1, NULL, NULL);
    grad_u1(w2, k2, beta, h, r, G);
    cplx dwdz2 = c3_dotr(m33_mul(G, qv), q);
    wave_kin1(beta, w2, k2, h, r, 1025.0, g, &u2, NULL, NULL);
    vel1 = c3_sub(vel1, proj_q(q, vel1));
    vel2 = c3_sub(vel2, proj_q(q, vel2));
    u1 = c3_sub(u1, proj_q(q, u1));
    u2 = c3_sub(u2, proj_q(q, u2));
    c3 acc = c3_add(c3_scale(c3_conj(c3_sub(u2, vel2)), 0.25 * dwdz1), c3_scale(c3_sub(u1, vel1), 0.25 * conj(dwdz2)));
    acc = c3_sub(acc, proj_q(q, acc));
    return acc;
}

/* helpers.py:337-373 getWaveKin_pot2ndOrd (beta1 == beta2 == beta) */
static void pot_2nd(double w1, double w2, double k1, double k2, double beta, double h, const double *r, double g, double rho,
                    c3 *acc, cplx *p)
{
    *acc = c3_zero(); *p = 0;
    if (w1 == w2) return;
    double b = beta * 0.017453292519943295, cosB = cos(b), sinB = sin(b), z = r[2];
    if (z <= 0 && k1 > 0 && k2 > 0) {
        double kk[3] = { k1 * cosB - k2 * cosB, k1 * sinB - k2 * sinB, 0 };
        double nk = sqrt(kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2]);
        double t1 = tanh(k1 * h), t2 = tanh(k2 * h);
        cplx g12 = (-I * g / (2 * w1)) * ((k1 * k1) * (1 - t1 * t1) - 2 * k1 * k2 * (1 + t1 * t2)) / ((w1 - w2) * (w1 - w2) / g - nk * tanh(nk * h));
        cplx g21 = (-I * g / (2 * w2)) * ((k2 * k2) * (1 - t2 * t2) - 2 * k2 * k1 * (1 + t2 * t1)) / ((w2 - w1) * (w2 - w1) / g - nk * tanh(nk * h));
        cplx aux = 0.5 * (g21 + conj(g12));
        double khz_xy = cosh(nk * (z + h)) / cosh(nk * h), khz_z = sinh(nk * (z + h)) / cosh(nk * h);
        cplx ph = cexp(-I * (kk[0] * r[0] + kk[1] * r[1] + kk[2] * r[2]));
        acc->v[0] = aux * khz_xy * ph; acc->v[1] = acc->v[0];
        acc->v[0] *= (w1 - w2) * (k1 * cosB - k2 * cosB);
        acc->v[1] *= (w1 - w2) * (k1 * sinB - k2 * sinB);
        acc->v[2] = aux * khz_z * ph;
        acc->v[2] *= I * (w1 - w2) * nk;
        *p = aux * khz_xy * ph;
        *p *= -I * rho * (w1 - w2);
    }
}

static cplx hankel1(int n, double x)
{
    if (n < 0) { cplx hv = jn(-n, x) + I * yn(-n, x); return ((-n) & 1) ? -hv : hv; }   /* H_{-n} = (-1)^n H_n */
    return jn(n, x) + I * yn(n, x);
}
static cplx kay_omega(double k1R, double k2R, int n)
{
    cplx H_N_ii = 0.5 * (hankel1(n - 1, k1R) - hankel1(n + 1, k1R));
    cplx H_N_jj = 0.5 * conj(hankel1(n - 1, k2R) - hankel1(n + 1, k2R));
    cplx H_Nm1_ii = 0.5 * (hankel1(n, k1R) - hankel1(n + 2, k1R));
    cplx H_Nm1_jj = 0.5 * conj(hankel1(n, k2R) - hankel1(n + 2, k2R));
    return 1 / (H_Nm1_ii * H_N_jj) - 1 / (H_N_ii * H_Nm1_jj);
}

/* raft_member.py:1680-1792 correction_KAY for member m, Nm = 10 */
static void correction_kay(const ro_qtf_design *d, int m, double w1, double w2, double k1, double k2, double beta, cplx F[6])
{
    for (int a = 0; a < 6; a++) F[a] = 0;
    if (!d->mem_mcf[m]) return;
    double h = d->depth, rho = d->rho, g = d->g;
    const double *p1 = d->mem_p1 + 3 * m, *p2 = d->mem_p2 + 3 * m;
    double cosB = cos(beta), sinB = sin(beta);
    double kk[3] = { k1 * cosB - k2 * cosB, k1 * sinB - k2 * sinB, 0 };
    double bv[3] = { cosB, sinB, 0 };
    double d1 = bv[0] * p1[0] + bv[1] * p1[1] + bv[2] * p1[2], d2 = bv[0] * p2[0] + bv[1] * p2[1] + bv[2] * p2[2];
    double pf[3] = { d1 * p1[0] + d2 * p2[0], d1 * p1[1] + d2 * p2[1], d1 * p1[2] + d2 * p2[2] };
    double nrm = sqrt(pf[0] * pf[0] + pf[1] * pf[1] + pf[2] * pf[2]);
    for (int i = 0; i < 3; i++) pf[i] /= nrm;
    {   /* mem_mcf already folds in rA_z * rB_z < 0 (:1721, :1741): without a waterline crossing the correction is zero */
        const double *rwl = d->mem_rwl + 3 * m;
        double R = d->mem_R_wl[m], k1R = k1 * R, k2R = k2 * R;
        cplx Fwl = 0;
        for (int nn = 0; nn <= 10; nn++) Fwl += -rho * g * R * 2 * I / M_PI / (k1R * k2R) * kay_omega(k1R, k2R, nn);
        cplx ph = cexp(-I * (kk[0] * rwl[0] + kk[1] * rwl[1] + kk[2] * rwl[2]));
        cplx Fs = creal(Fwl) * ph;
        cplx f3[3] = { Fs * pf[0], Fs * pf[1], Fs * pf[2] }, f6[6];
        translate_force(f3, rwl, f6);
        for (int a = 0; a < 6; a++) F[a] += f6[a];
        for (int s = 0; s < d->n_seg; s++) {
            if (d->seg_mem[s] != m) continue;
            double z1 = d->seg_z1[s], z2 = d->seg_z2[s];
            R = d->seg_R[s]; k1R = k1 * R; k2R = k2 * R;
            double H = h / R, k1h = k1R * H, k2h = k2R * H, Im, Ip;
            if (w1 == w2) {
                Im = 0.5 * (sinh((k1 + k2) * (z2 + h)) / (k1h + k2h) - (z2 + h) / h - sinh((k1 + k2) * (z1 + h)) / (k1h + k2h) + (z1 + h) / h);
                Ip = 0.5 * (sinh((k1 + k2) * (z2 + h)) / (k1h + k2h) + (z2 + h) / h - sinh((k1 + k2) * (z1 + h)) / (k1h + k2h) - (z1 + h) / h);
            } else {
                Im = 0.5 * (sinh((k1 + k2) * (z2 + h)) / (k1h + k2h) - sinh((k1 - k2) * (z2 + h)) / (k1h - k2h) - sinh((k1 + k2) * (z1 + h)) / (k1h + k2h) + sinh((k1 - k2) * (z1 + h)) / (k1h - k2h));
                Ip = 0.5 * (sinh((k1 + k2) * (z2 + h)) / (k1h + k2h) + sinh((k1 - k2) * (z2 + h)) / (k1h - k2h) - sinh((k1 + k2) * (z1 + h)) / (k1h + k2h) - sinh((k1 - k2) * (z1 + h)) / (k1h - k2h));
            }
            double c1 = cosh(k1h), c2 = cosh(k2h);
            cplx dF = 0;
            for (int nn = 0; nn <= 10; nn++)
                dF += rho * g * R * 2 * I / M_PI / (k1R * k2R) * kay_omega(k1R, k2R, nn)
                      * (k1h * k2h / sqrt(k1h * tanh(k1h)) / sqrt(k2h * tanh(k2h)) * (Im + Ip * nn * (nn + 1) / k1R / k2R) / c1 / c2);
            cplx dFs = creal(dF) * ph;
            cplx g3[3] = { dFs * pf[0], dFs * pf[1], dFs * pf[2] }, g6[6];
            translate_force(g3, d->seg_rmid + 3 * s, g6);
            for (int a = 0; a < 6; a++) F[a] += g6[a];
        }
    }
    if (k1 < k2) for (int a = 0; a < 6; a++) F[a] = conj(F[a]);
}

/* Member.calcQTF_slenderBody for member m, accumulated into qtf [nw][nw][6] (upper triangle w2 >= w1) */
static void member_qtf(const ro_qtf_design *d, int m, int nw, const double *w, const double *k, double beta, const cplx *Xi, cplx *qtf)
{
    double h = d->depth, rho = d->rho, g = d->g;
    const double *q = d->mem_q + 3 * m, *p1 = d->mem_p1 + 3 * m, *p2 = d->mem_p2 + 3 * m;
    int j0 = -1, j1 = -1;
    for (int j = 0; j < d->n_nodes; j++) if (d->node_mem[j] == m) { if (j0 < 0) j0 = j; j1 = j + 1; }
    int ns = j0 < 0 ? 0 : j1 - j0;
    /* per node, per frequency kinematics (raft_member.py:1501-1521) */
    c3 *dr = malloc(sizeof(c3) * (size_t)(ns + 1) * nw), *nodeV = malloc(sizeof(c3) * (size_t)(ns + 1) * nw), *u = malloc(sizeof(c3) * (size_t)(ns + 1) * nw);
    c3 *gp = malloc(sizeof(c3) * (size_t)(ns + 1) * nw);
    cplx (*G)[3][3] = malloc(sizeof(cplx[3][3]) * (size_t)(ns + 1) * nw), (*Gd)[3][3] = malloc(sizeof(cplx[3][3]) * (size_t)(ns + 1) * nw);
    cplx *vax = malloc(sizeof(cplx) * (size_t)(ns + 1) * nw);
    cplx *t3 = malloc(sizeof(cplx) * 3 * nw), *t3b = malloc(sizeof(cplx) * 3 * nw), *t3c = malloc(sizeof(cplx) * 3 * nw);
    for (int n = 0; n < ns; n++) {
        const double *r = d->node_r + 3 * (j0 + n);
        ro_get_kinematics(r, Xi, w, nw, t3, t3b, t3c);
        for (int i = 0; i < nw; i++) {
            size_t e = (size_t)n * nw + i;
            for (int a = 0; a < 3; a++) { dr[e].v[a] = t3[a * nw + i]; nodeV[e].v[a] = t3b[a * nw + i]; }
            wave_kin1(beta, w[i], k[i], h, r, rho, g, &u[e], NULL, NULL);
            grad_u1(w[i], k[i], beta, h, r, G[e]);
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Gd[e][a][b] = I * w[i] * G[e][a][b];
            vax[e] = c3_dotr(c3_sub(u[e], nodeV[e]), q);
            gp[e] = grad_pres1st(k[i], beta, h, r, rho, g);
            nodeV[e] = c3_sub(nodeV[e], proj_q(q, nodeV[e]));        /* the in-place side effect of getWaveKin_axdivAcc */
        }
    }
    /* waterline quantities (:1524-1539) */
    cplx *eta_r = calloc(nw, sizeof(cplx));
    c3 *ud_wl = calloc(nw, sizeof(c3)), *a_wl = calloc(nw, sizeof(c3)), *g_e1 = calloc(nw, sizeof(c3));
    const double *r_int = d->mem_r_int + 3 * m;
    if (d->mem_wl[m]) {
        ro_get_kinematics(r_int, Xi, w, nw, t3, t3b, t3c);
        for (int i = 0; i < nw; i++) {
            cplx eta;
            wave_kin1(beta, w[i], k[i], h, r_int, 1.0, 1.0, NULL, &ud_wl[i], &eta);
            for (int a = 0; a < 3; a++) a_wl[i].v[a] = t3c[a * nw + i];
            eta_r[i] = eta - t3[2 * nw + i];
        }
    }
    for (int i = 0; i < nw; i++) {
        const cplx *th = Xi + 3 * nw;      /* Xi[3:, i] = th[0*nw+i], th[nw+i], th[2nw+i] */
        cplx a = th[i], b = th[nw + i];
        cplx c1z = a * p1[1] - b * p1[0], c2z = a * p2[1] - b * p2[0];       /* cross(theta, p)[2] */
        for (int x = 0; x < 3; x++) g_e1[i].v[x] = -g * (c1z * p1[x] + c2z * p2[x]);
    }
    for (int i1 = 0; i1 < nw; i1++) {
        for (int i2 = 0; i2 < nw; i2++) {
            double w1 = w[i1], w2 = w[i2], k1 = k[i1], k2 = k[i2];
            if (w2 < w1) continue;
            cplx F[6] = {0, 0, 0, 0, 0, 0};
            cplx O1[3][3], O2[3][3];
            {   /* OMEGA = -getH(1j w Xi[3:]) */
                cplx a1[3] = { I * w1 * Xi[3 * nw + i1], I * w1 * Xi[4 * nw + i1], I * w1 * Xi[5 * nw + i1] };
                cplx a2[3] = { I * w2 * Xi[3 * nw + i2], I * w2 * Xi[4 * nw + i2], I * w2 * Xi[5 * nw + i2] };
                cplx H1[3][3] = { { 0, a1[2], -a1[1] }, { -a1[2], 0, a1[0] }, { a1[1], -a1[0], 0 } };
                cplx H2[3][3] = { { 0, a2[2], -a2[1] }, { -a2[2], 0, a2[0] }, { a2[1], -a2[0], 0 } };
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { O1[a][b] = -H1[a][b]; O2[a][b] = -H2[a][b]; }
            }
            for (int n = 0; n < ns; n++) {
                int jg = j0 + n;
                const double *r = d->node_r + 3 * jg;
                size_t e1 = (size_t)n * nw + i1, e2 = (size_t)n * nw + i2;
                double Ca1 = d->node_Ca_p1[jg], Ca2 = d->node_Ca_p2[jg], CaE = d->node_Ca_End[jg];
                double v_i = d->node_v_side[jg], v_e = d->node_v_end[jg], a_i = d->node_a_i[jg];
                c3 acc2; cplx p2nd;
                pot_2nd(w1, w2, k1, k2, beta, h, r, g, rho, &acc2, &p2nd);
                c3 f_2ndPot = c3_scale(proj_p(p1, p2, 1. + Ca1, 1. + Ca2, acc2), rho * v_i);
                c3 conv_acc = c3_scale(c3_add(m33_mul(G[e1], c3_conj(u[e2])), m33c_mul(G[e2], u[e1])), 0.25);
                c3 f_conv = c3_scale(proj_p(p1, p2, 1. + Ca1, 1. + Ca2, conv_acc), rho * v_i);
                c3 f_axdv = c3_scale(proj_p(p1, p2, Ca1, Ca2, axdiv_acc(w1, w2, k1, k2, beta, h, r, nodeV[e1], nodeV[e2], q, g)), rho * v_i);
                c3 acc_nabla = c3_add(c3_scale(m33_mul(Gd[e1], c3_conj(dr[e2])), 0.25), c3_scale(m33c_mul(Gd[e2], dr[e1]), 0.25));
                c3 f_nabla = c3_scale(proj_p(p1, p2, 1. + Ca1, 1. + Ca2, acc_nabla), rho * v_i);
                c3 t = c3_add(m33_mul(O1, c3_conj(c3_vecr(q, vax[e2]))), m33c_mul(O2, c3_vecr(q, vax[e1])));
                c3 f_rslb = c3_scale(proj_p(p1, p2, Ca1, Ca2, t), -0.25 * 2);
                f_rslb = c3_scale(f_rslb, rho * v_i);
                c3 u1a = c3_sub(u[e1], nodeV[e1]), u2a = c3_sub(u[e2], nodeV[e2]);
                cplx V1[3][3], V2[3][3];
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { V1[a][b] = G[e1][a][b] + O1[a][b]; V2[a][b] = G[e2][a][b] + O2[a][b]; }
                c3 aux = c3_scale(c3_add(m33_mul(V1, c3_conj(proj_p(p1, p2, Ca1, Ca2, u2a))), m33c_mul(V2, proj_p(p1, p2, Ca1, Ca2, u1a))), 0.25);
                aux = c3_sub(aux, proj_q(q, aux));
                f_rslb = c3_add(f_rslb, c3_scale(aux, rho * v_i));
                u1a = c3_sub(u1a, proj_q(q, u1a));
                u2a = c3_sub(u2a, proj_q(q, u2a));
                aux = c3_scale(c3_add(proj_p(p1, p2, Ca1, Ca2, m33_mul(V1, c3_conj(u2a))), proj_p(p1, p2, Ca1, Ca2, m33c_mul(V2, u1a))), 0.25);
                f_rslb = c3_add(f_rslb, c3_scale(aux, -rho * v_i));
                /* axial / end effects */
                f_2ndPot = c3_add(f_2ndPot, c3_vecr(q, a_i * p2nd));
                f_2ndPot = c3_add(f_2ndPot, c3_scale(proj_q(q, acc2), rho * v_e * CaE));
                f_conv = c3_add(f_conv, c3_scale(proj_q(q, conv_acc), rho * v_e * CaE));
                f_nabla = c3_add(f_nabla, c3_scale(proj_q(q, acc_nabla), rho * v_e * CaE));
                cplx p_nabla = 0.25 * c3_dot(gp[e1], c3_conj(dr[e2])) + 0.25 * c3_dot(c3_conj(gp[e2]), dr[e1]);
                f_nabla = c3_add(f_nabla, c3_vecr(q, a_i * p_nabla));
                cplx p_drop = -2 * 0.25 * 0.5 * rho * c3_dot(proj_p(p1, p2, 1.0, 1.0, c3_sub(u[e1], nodeV[e1])),
                                                               c3_conj(proj_p(p1, p2, Ca1, Ca2, c3_sub(u[e2], nodeV[e2]))));
                f_conv = c3_add(f_conv, c3_vecr(q, a_i * p_drop));
                u1a = proj_p(p1, p2, Ca1, Ca2, u1a);
                u2a = proj_p(p1, p2, Ca1, Ca2, u2a);
                c3 f_transv = c3_scale(c3_add(c3_scale(c3_conj(u1a), vax[e2]), c3_scale(u2a, conj(vax[e1]))), 0.25 * a_i * rho);
                f_conv = c3_add(f_conv, f_transv);
                c3 parts[5] = { f_2ndPot, f_conv, f_axdv, f_nabla, f_rslb };
                for (int pz = 0; pz < 5; pz++) {
                    cplx f6[6];
                    translate_force(parts[pz].v, r, f6);
                    for (int a = 0; a < 6; a++) F[a] += f6[a];
                }
            }
            if (d->mem_wl[m]) {          /* relative-wave-elevation force at the waterline (:1655-1683) */
                double a_i = d->mem_a_wl[m];
                double Ca1 = 0, Ca2 = 0;
                if (ns > 0) { Ca1 = d->node_Ca_p1[j1 - 1]; Ca2 = d->node_Ca_p2[j1 - 1]; }   /* the loop variables keep the last submerged node's values */
                c3 fe = c3_scale(c3_add(c3_scale(ud_wl[i1], conj(eta_r[i2])), c3_scale(c3_conj(ud_wl[i2]), eta_r[i1])), 0.25);
                fe = c3_scale(proj_p(p1, p2, 1. + Ca1, 1. + Ca2, fe), rho * a_i);
                c3 ae = c3_scale(c3_add(c3_scale(a_wl[i1], conj(eta_r[i2])), c3_scale(c3_conj(a_wl[i2]), eta_r[i1])), 0.25);
                fe = c3_sub(fe, c3_scale(proj_p(p1, p2, Ca1, Ca2, ae), rho * a_i));
                fe = c3_sub(fe, c3_scale(c3_add(c3_scale(g_e1[i1], conj(eta_r[i2])), c3_scale(c3_conj(g_e1[i2]), eta_r[i1])), 0.25 * rho * a_i));
                cplx f6[6];
                translate_force(fe.v, r_int, f6);
                for (int a = 0; a < 6; a++) F[a] += f6[a];
            }
            cplx K[6];
            correction_kay(d, m, w1, w2, k1, k2, beta, K);
            for (int a = 0; a < 6; a++) qtf[((size_t)i1 * nw + i2) * 6 + a] += F[a] + K[a];
        }
    }
    free(dr); free(nodeV); free(u); free(gp); free(G); free(Gd); free(vax); free(t3); free(t3b); free(t3c);
    free(eta_r); free(ud_wl); free(a_wl); free(g_e1);
}

/* FOWT.calcQTF_slenderBody (raft_fowt.py:1988-2078): Xi [6][nw] RAOs already on the second-order grid w (:2021-2023).
 * qtf [nw][nw][6], Hermitian-filled (:2068-2070). */
void ro_qtf_slender(const ro_qtf_design *d, int nw, const double *w, const double *k, double beta, const cplx *Xi, cplx *qtf)
{
    for (size_t e = 0; e < (size_t)nw * nw * 6; e++) qtf[e] = 0;
    /* Pinkster IV: rotation of the first-order forces, F1st = M_struc (-w^2 Xi) (:2044-2058) */
    cplx *F1 = malloc(sizeof(cplx) * 6 * nw);
    for (int a = 0; a < 6; a++)
        for (int i = 0; i < nw; i++) {
            cplx s = 0;
            for (int b = 0; b < 6; b++) s += d->M_struc[6 * a + b] * (-w[i] * w[i] * Xi[b * nw + i]);
            F1[a * nw + i] = s;
        }
#define CROSS(o, a0, a1, a2, b0, b1, b2) do { o[0] = (a1) * (b2) - (a2) * (b1); o[1] = (a2) * (b0) - (a0) * (b2); o[2] = (a0) * (b1) - (a1) * (b0); } while (0)
    for (int i1 = 0; i1 < nw; i1++)
        for (int i2 = i1; i2 < nw; i2++) {
            if (w[i2] < w[i1]) continue;
            cplx t1[3], t2[3], x1[3] = { Xi[3 * nw + i1], Xi[4 * nw + i1], Xi[5 * nw + i1] };
            cplx x2c[3] = { conj(Xi[3 * nw + i2]), conj(Xi[4 * nw + i2]), conj(Xi[5 * nw + i2]) };
            cplx *o = qtf + ((size_t)i1 * nw + i2) * 6;
            CROSS(t1, x1[0], x1[1], x1[2], conj(F1[0 * nw + i2]), conj(F1[1 * nw + i2]), conj(F1[2 * nw + i2]));
            CROSS(t2, x2c[0], x2c[1], x2c[2], F1[0 * nw + i1], F1[1 * nw + i1], F1[2 * nw + i1]);
            for (int a = 0; a < 3; a++) o[a] = 0.25 * (t1[a] + t2[a]);
            CROSS(t1, x1[0], x1[1], x1[2], conj(F1[3 * nw + i2]), conj(F1[4 * nw + i2]), conj(F1[5 * nw + i2]));
            CROSS(t2, x2c[0], x2c[1], x2c[2], F1[3 * nw + i1], F1[4 * nw + i1], F1[5 * nw + i1]);
            for (int a = 0; a < 3; a++) o[3 + a] = 0.25 * (t1[a] + t2[a]);
        }
#undef CROSS
    free(F1);
    for (int m = 0; m < d->n_members; m++) member_qtf(d, m, nw, w, k, beta, Xi, qtf);
    for (int i1 = 0; i1 < nw; i1++)
        for (int i2 = i1 + 1; i2 < nw; i2++)
            for (int a = 0; a < 6; a++) {
                cplx up = qtf[((size_t)i1 * nw + i2) * 6 + a], lo = qtf[((size_t)i2 * nw + i1) * 6 + a];
                qtf[((size_t)i1 * nw + i2) * 6 + a] = up + conj(lo);
                qtf[((size_t)i2 * nw + i1) * 6 + a] = lo + conj(up);
            }
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
    int passes = 0, conv = 0, flagQTF = 0;
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
        if (all) {
            if (!d->qs || flagQTF) { conv = 1; break; }                       /* :1106-1107 */
            /* potSecOrder 1 (:1108-1131): QTF from the motions just found, second-order force added to the excitation,
               and the loop goes on from the SAME XiLast with its counter reset (iiter = 0, then += 1) */
            iiter = 0;
            int n2 = d->qs_nw;
            cplx *rao = malloc(sizeof(cplx) * 6 * nw), *Xi2 = malloc(sizeof(cplx) * 6 * n2), *qtf = malloc(sizeof(cplx) * (size_t)n2 * n2 * 6);
            for (int a = 0; a < 6; a++)
                for (int i = 0; i < nw; i++) rao[a * nw + i] = (fabs(zeta[i]) > 1e-6) ? Xi[a * nw + i] / zeta[i] : 0;   /* helpers.getRAO */
            for (int a = 0; a < 6; a++)                                      /* np.interp(w1_2nd, w, Xi0, left=0, right=0)  raft_fowt.py:2021-2023 */
                for (int j = 0; j < n2; j++) {
                    double x = d->qs_w[j];
                    cplx v = 0;
                    if (x >= d->w[0] && x <= d->w[nw - 1]) {
                        if (x == d->w[nw - 1]) v = rao[a * nw + nw - 1];
                        else {
                            int i0 = 0;
                            while (i0 < nw - 2 && d->w[i0 + 1] <= x) i0++;
                            cplx y0 = rao[a * nw + i0], y1 = rao[a * nw + i0 + 1];
                            cplx slope = (y1 - y0) / (d->w[i0 + 1] - d->w[i0]);
                            v = slope * (x - d->w[i0]) + y0;
                        }
                    }
                    Xi2[a * n2 + j] = v;
                }
            ro_qtf_slender(d->qs, n2, d->qs_w, d->qs_k, beta_deg * 0.017453292519943295, Xi2, qtf);
            ro_design dq = *d;                                                /* calcHydroForce_2ndOrd with fowt.qtf, heads_2nd = [beta] */
            double head = beta_deg * 0.017453292519943295;
            dq.n_qtf_w = n2; dq.n_qtf_head = 1; dq.qtf_w = d->qs_w; dq.qtf_heads = &head; dq.qtf = qtf;
            double *S = malloc(sizeof(double) * nw), *zz = malloc(sizeof(double) * nw), fm[6];
            ro_sea_state(d->w, nw, d->dw, spec, Hs, Tp, gamma, S, zz);
            ro_hydro_force_2nd(&dq, head, S, fm, F2);                         /* F_lin += Fhydro_2nd  (:1129-1130) */
            free(rao); free(Xi2); free(qtf); free(S); free(zz);
            flagQTF = 1;
            continue;
        }
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

/* ======================================================================================================
 * Generalised degrees of freedom (flexible members, FOWT.nDOF > 6): FOWT.calcHydroExcitation / calcHydroLinearization
 * with the transformation matrix fowt.T (raft_fowt.py:1854-1857, :1886-1888, :1913-1929; raft_member.py:1960-1992,
 * 2040-2052, 2119-2126).  Every strip node carries the 6 x nDOF block of T of its structural node (Tn) and its offset
 * from that node (rr; zero for the nodes of a flexible member): node motion = Tn Xi, node load -> Tn^T [f ; rr x f].
 * GROUNDWORK for the next row of SURVEY.md 8(f): the CUDA path is rigid 6-DOF only; this pins the checker first.
 * ====================================================================================================== */
typedef struct {
    const ro_design *d;          /* node tables (node_r, node_mem, Imat, areas, coefficients, member frames, grid)  */
    int nDOF;
    const double *Tn;            /* [Ns][6][nDOF]                                                                  */
    const double *rr;            /* [Ns][3]                                                                        */
} ro_general;

/* zeta [nw], F_iner [nDOF][nw], u [Ns][3][nw] */
int ro_general_excitation(const ro_general *g, int spec, double Hs, double Tp, double gamma, double beta_deg,
                          double *zeta, cplx *F_iner, cplx *u_out)
{
    const ro_design *d = g->d;
    int nw = d->nw, Ns = d->n_nodes, n = g->nDOF;
    double *S = malloc(sizeof(double) * nw);
    int rc = ro_sea_state(d->w, nw, d->dw, spec, Hs, Tp, gamma, S, zeta);
    free(S);
    if (rc) return rc;
    double beta = beta_deg * (M_PI / 180.0);
    cplx *ud = malloc(sizeof(cplx) * 3 * nw), *pD = malloc(sizeof(cplx) * nw);
    for (size_t i = 0; i < (size_t)n * nw; i++) F_iner[i] = 0;
    for (int il = 0; il < Ns; il++) {
        const double *r = d->node_r + 3 * il, *q = d->mem_q + 3 * d->node_mem[il];
        const double *Tn = g->Tn + (size_t)il * 6 * n, *rr = g->rr + 3 * il;
        cplx *u = u_out + (size_t)il * 3 * nw;
        ro_wave_kin(zeta, beta, d->w, d->k, d->depth, r, nw, 1025.0, 9.81, u, ud, pD);
        const double *Im = d->node_Imat + 9 * il;
        for (int i = 0; i < nw; i++) {
            cplx f[3], f6[6];
            for (int a = 0; a < 3; a++) {
                if (d->node_Imat_w) {
                    const cplx *Iw = d->node_Imat_w + (size_t)il * 9 * nw;
                    f[a] = Iw[(3 * a) * nw + i] * ud[i] + Iw[(3 * a + 1) * nw + i] * ud[nw + i]
                         + Iw[(3 * a + 2) * nw + i] * ud[2 * nw + i] + pD[i] * d->node_a_i[il] * q[a];
                } else
                f[a] = Im[3 * a] * ud[i] + Im[3 * a + 1] * ud[nw + i] + Im[3 * a + 2] * ud[2 * nw + i]
                     + pD[i] * d->node_a_i[il] * q[a];
            }
            translate_force(f, rr, f6);
            for (int c = 0; c < n; c++) {                              /* T^T (fowt:1888) */
                cplx s = 0;
                for (int b = 0; b < 6; b++) s += Tn[b * n + c] * f6[b];
                F_iner[(size_t)c * nw + i] += s;
            }
        }
    }
    free(ud); free(pD);
    return 0;
}

/* Xi [nDOF][nw] in; Bmat [Ns][3][3], B_drag [nDOF][nDOF], F_drag [nDOF][nw] out */
void ro_general_linearization(const ro_general *g, const cplx *u_all, const cplx *Xi, double *Bmat_all, double *B_drag, cplx *F_drag)
{
    const ro_design *d = g->d;
    int nw = d->nw, Ns = d->n_nodes, n = g->nDOF;
    cplx *Xin = malloc(sizeof(cplx) * 6 * nw), *vrel = malloc(sizeof(cplx) * 3 * nw);
    for (size_t i = 0; i < (size_t)n * n; i++) B_drag[i] = 0;
    for (size_t i = 0; i < (size_t)n * nw; i++) F_drag[i] = 0;
    for (int il = 0; il < Ns; il++) {
        int m = d->node_mem[il];
        const double *q = d->mem_q + 3 * m, *p1 = d->mem_p1 + 3 * m, *p2 = d->mem_p2 + 3 * m;
        const double *Tn = g->Tn + (size_t)il * 6 * n, *rr = g->rr + 3 * il;
        const cplx *u = u_all + (size_t)il * 3 * nw;
        int circ = d->mem_circ[m];
        for (int i = 0; i < nw; i++)                                   /* Xi_nodes = node.T @ Xi (fowt:1921) */
            for (int a = 0; a < 6; a++) {
                cplx s = 0;
                for (int b = 0; b < n; b++) s += Tn[a * n + b] * Xi[(size_t)b * nw + i];
                Xin[a * nw + i] = s;
            }
        double sq = 0, sp = 0, sp1 = 0, sp2 = 0;
        for (int i = 0; i < nw; i++) {
            const cplx th0 = Xin[3 * nw + i], th1 = Xin[4 * nw + i], th2 = Xin[5 * nw + i];
            cplx dr[3], v[3];
            dr[0] = Xin[0 * nw + i] + (-th2 * rr[1] + th1 * rr[2]);
            dr[1] = Xin[1 * nw + i] + ( th2 * rr[0] - th0 * rr[2]);
            dr[2] = Xin[2 * nw + i] + (-th1 * rr[0] + th0 * rr[1]);
            for (int a = 0; a < 3; a++) { v[a] = I * d->w[i] * dr[a]; vrel[a * nw + i] = u[a * nw + i] - v[a]; }
            cplx aq = 0, a1 = 0, a2 = 0;
            for (int a = 0; a < 3; a++) { aq += vrel[a * nw + i] * q[a]; a1 += vrel[a * nw + i] * p1[a]; a2 += vrel[a * nw + i] * p2[a]; }
            for (int a = 0; a < 3; a++) {
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
        /* B_drag += Tn^T B6 Tn  (fowt:1927) */
        double *TB = malloc(sizeof(double) * 6 * n);
        for (int a = 0; a < 6; a++) for (int cc = 0; cc < n; cc++) {
            double s = 0; for (int l = 0; l < 6; l++) s += B6[a][l] * Tn[l * n + cc]; TB[a * n + cc] = s;
        }
        for (int r_ = 0; r_ < n; r_++) for (int cc = 0; cc < n; cc++) {
            double s = 0; for (int l = 0; l < 6; l++) s += Tn[l * n + r_] * TB[l * n + cc]; B_drag[(size_t)r_ * n + cc] += s;
        }
        free(TB);
        for (int i = 0; i < nw; i++) {
            cplx f[3], f6[6];
            for (int a = 0; a < 3; a++) f[a] = Bmat[a][0] * u[i] + Bmat[a][1] * u[nw + i] + Bmat[a][2] * u[2 * nw + i];
            translate_force(f, rr, f6);
            for (int cc = 0; cc < n; cc++) {
                cplx s = 0;
                for (int b = 0; b < 6; b++) s += Tn[b * n + cc] * f6[b];
                F_drag[(size_t)cc * nw + i] += s;
            }
        }
    }
    free(Xin); free(vrel);
}

/* Model.solveDynamics for one FOWT with nDOF reduced degrees of freedom and one single-train case
 * (raft_model.py:994-1156, 1189-1216): M, B, C [nDOF][nDOF] constant matrices (M_struc + A_hydro_morison + ..., B_struc + ...,
 * C_struc + C_hydro + C_moor + C_elast).  Xi_out [nDOF][nw]; status = {passes, converged, nan}. */
int ro_general_solve_dynamics(const ro_general *g, const double *M, const double *B, const double *Cm, int spec, double Hs, double Tp,
                              double gamma, double beta_deg, int nIter, double tol, double XiStart, cplx *Xi_out, int *status)
{
    const ro_design *d = g->d;
    int nw = d->nw, Ns = d->n_nodes > 0 ? d->n_nodes : 1, n = g->nDOF;
    double *zeta = malloc(sizeof(double) * nw), *Bmat = malloc(sizeof(double) * 9 * Ns), *B_drag = malloc(sizeof(double) * n * n);
    cplx *F_iner = malloc(sizeof(cplx) * n * nw), *u = malloc(sizeof(cplx) * (size_t)Ns * 3 * nw);
    cplx *XiLast = malloc(sizeof(cplx) * n * nw), *Xi = malloc(sizeof(cplx) * n * nw), *F_drag = malloc(sizeof(cplx) * n * nw);
    cplx *Z = malloc(sizeof(cplx) * (size_t)n * n * nw);
    status[0] = status[1] = status[2] = 0;
    int rc = ro_general_excitation(g, spec, Hs, Tp, gamma, beta_deg, zeta, F_iner, u);
    if (rc) goto done;
    for (int i = 0; i < n * nw; i++) XiLast[i] = XiStart;
    int passes = 0, conv = 0;
    for (int iiter = 0; iiter < nIter + 1; iiter++) {
        ro_general_linearization(g, u, XiLast, Bmat, B_drag, F_drag);
        passes++;
        int nan = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(|:nan)
#endif
        for (int ii = 0; ii < nw; ii++) {
            cplx *A = malloc(sizeof(cplx) * n * n), *b = malloc(sizeof(cplx) * n);
            double wv = d->w[ii];
            for (int a = 0; a < n; a++) {
                for (int c = 0; c < n; c++)
                    A[a * n + c] = -wv * wv * M[a * n + c] + I * wv * (B[a * n + c] + B_drag[a * n + c]) + Cm[a * n + c];
                b[a] = F_iner[a * nw + ii] + F_drag[a * nw + ii];
            }
            memcpy(Z + (size_t)ii * n * n, A, sizeof(cplx) * n * n);
            ro_zgesv(n, A, b);
            for (int a = 0; a < n; a++) { Xi[a * nw + ii] = b[a]; if (isnan(creal(b[a])) || isnan(cimag(b[a]))) nan = 1; }
            free(A); free(b);
        }
        if (nan) { status[2] = 1; break; }
        int all = 1;
        for (int i = 0; i < n * nw; i++) {
            double tc = cabs(Xi[i] - XiLast[i]) / (cabs(Xi[i]) + tol);
            if (!(tc < tol)) { all = 0; break; }
        }
        if (all) { conv = 1; break; }
        for (int i = 0; i < n * nw; i++) XiLast[i] = 0.2 * XiLast[i] + 0.8 * Xi[i];
    }
    status[0] = passes; status[1] = conv;
    /* system response through the explicit inverse of the last Z (raft_model.py:1189-1216) */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int ii = 0; ii < nw; ii++) {
        cplx *A = malloc(sizeof(cplx) * n * n), *Ai = malloc(sizeof(cplx) * n * n);
        memcpy(A, Z + (size_t)ii * n * n, sizeof(cplx) * n * n);
        if (ro_zinv(n, A, Ai)) { for (int a = 0; a < n; a++) Xi_out[a * nw + ii] = NAN; }
        else for (int a = 0; a < n; a++) {
            cplx s = 0;
            for (int c = 0; c < n; c++) s += Ai[a * n + c] * (F_iner[c * nw + ii] + F_drag[c * nw + ii]);
            Xi_out[a * nw + ii] = s;
        }
        free(A); free(Ai);
    }
done:
    free(zeta); free(Bmat); free(B_drag); free(F_iner); free(u); free(XiLast); free(Xi); free(F_drag); free(Z);
    return rc;
}
