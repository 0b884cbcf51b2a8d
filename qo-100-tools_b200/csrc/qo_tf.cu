/*
 * qo_tf.cu -- instantiations, launcher and plan-time self-check of the transfer-function kernel (qo_tf.cuh).
 * Its own translation unit (compiles in parallel with qo_cuda.cu / qo_ladder.cu).
 */
#include <cuda_runtime.h>
#include <complex>
#include <vector>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "qo_tf.cuh"
#include "qo_tf_launch.h"

/* launch shapes (tools/tf_sweep.py, profiles/r01h_*): the |S21| modes run 8 points per thread (coefficient loads and loop
 * overhead amortised over 8 Horner sets), the coupler mode 4 (six chains + the coupler block's state) */
#define QO_TF_TPB 128
#define QO_TF_MINB 4
#define QO_TF_CPL_TPB 128
#define QO_TF_CPL_MINB 3

typedef void (*tf_fn)(const TfParams);

template <int MODE, int PP, int TPB, int MINB> static tf_fn tf_pick(int K)
{
#define QO_TF_ROW(KK) case KK: return qo_mc_tf_kernel<KK, MODE, PP, TPB, MINB>;
    switch (K) {
        QO_TF_ROW(1) QO_TF_ROW(2) QO_TF_ROW(3) QO_TF_ROW(4) QO_TF_ROW(5) QO_TF_ROW(6) QO_TF_ROW(7) QO_TF_ROW(8)
        QO_TF_ROW(9) QO_TF_ROW(10) QO_TF_ROW(11) QO_TF_ROW(12) QO_TF_ROW(13) QO_TF_ROW(14) QO_TF_ROW(15)
    default: return nullptr;
    }
#undef QO_TF_ROW
}

extern "C" int qo_tf_launch(int K, int mode, int pp, int variant, int sm_count, const TfParams *P, cudaStream_t st)
{
    tf_fn fn = nullptr;
    int tpb = QO_TF_TPB, minb = QO_TF_MINB;
    (void)variant;
#ifdef QO_TF_EXPERIMENT
    /* development builds: launch shapes of the K = 12 |S21| kernel, chosen with QO100NET_LAD_VARIANT */
    if (mode == QO_TF_S21 && K == 12 && pp == 4) {
        switch (variant) {
        case 1: fn = qo_mc_tf_kernel<12, QO_TF_S21, 4, 128, 3>; tpb = 128; minb = 3; break;
        case 2: fn = qo_mc_tf_kernel<12, QO_TF_S21, 4, 256, 2>; tpb = 256; minb = 2; break;
        case 3: fn = qo_mc_tf_kernel<12, QO_TF_S21, 4, 64, 8>; tpb = 64; minb = 8; break;
        case 5: fn = qo_mc_tf_kernel<12, QO_TF_S21, 4, 64, 6>; tpb = 64; minb = 6; break;
        default: break;
        }
    }
    if (mode == QO_TF_CPL && K == 12) {
        if (pp == 2) switch (variant) {
        case 1: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 2, 128, 3>; tpb = 128; minb = 3; break;
        case 2: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 2, 128, 2>; tpb = 128; minb = 2; break;
        case 3: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 2, 64, 6>; tpb = 64; minb = 6; break;
        default: break;
        }
        if (pp == 1) switch (variant) {
        case 4: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 1, 256, 2>; tpb = 256; minb = 2; break;
        case 5: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 1, 256, 3>; tpb = 256; minb = 3; break;
        case 6: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 1, 128, 5>; tpb = 128; minb = 5; break;
        default: break;
        }
        if (pp == 4) switch (variant) {
        case 7: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 4, 128, 2>; tpb = 128; minb = 2; break;
        case 8: fn = qo_mc_tf_kernel<12, QO_TF_CPL, 4, 64, 4>; tpb = 64; minb = 4; break;
        default: break;
        }
    }
    if (!fn)
#endif
    switch (mode) {
    case QO_TF_S21: if (pp == 4) fn = tf_pick<QO_TF_S21, 4, QO_TF_TPB, QO_TF_MINB>(K); break;
    case QO_TF_S21_NOD: if (pp == 4) fn = tf_pick<QO_TF_S21_NOD, 4, QO_TF_TPB, QO_TF_MINB>(K); break;
    case QO_TF_S21_E: if (pp == 4) fn = tf_pick<QO_TF_S21_E, 4, QO_TF_TPB, QO_TF_MINB>(K); break;
    case QO_TF_CPL: if (pp == 2) { fn = tf_pick<QO_TF_CPL, 2, QO_TF_CPL_TPB, QO_TF_CPL_MINB>(K); tpb = QO_TF_CPL_TPB; minb = QO_TF_CPL_MINB; } break;
    case QO_TF_CPL_E: if (pp == 2) { fn = tf_pick<QO_TF_CPL_E, 2, QO_TF_CPL_TPB, QO_TF_CPL_MINB>(K); tpb = QO_TF_CPL_TPB; minb = QO_TF_CPL_MINB; } break;
    default: break;
    }
    if (!fn) return -1;
    const unsigned long long warps = (unsigned long long)(tpb / 32);
    unsigned long long blocks = (P->nsamples + warps - 1) / warps;
    const unsigned long long resident = (unsigned long long)sm_count * (unsigned long long)minb;
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    fn<<<(unsigned)blocks, tpb, 0, st>>>(*P);
    return (int)cudaGetLastError();
}

/* ---- plan-time self-check (host) ------------------------------------------------------------------------- */
typedef std::complex<double> cplx;

/* the device's expansion, in the same order and with the same normalisation */
static void tf_expand_host(const double (*rec)[QO_TF_REC], int n_el, double rl, double zn, double *p, double *q, double *d, double *ee)
{
    const int NC = 2 * QO_TF_MAXK + 2;
    for (int i = 0; i < NC; i++) p[i] = q[i] = d[i] = 0.0;
    p[0] = rl; q[0] = zn; d[0] = 1.0;
    /* E(y) = prod_e (E0 + E1 y + E2 y^2), all 2 n_el + 1 coefficients (the kernel keeps the first K) */
    const int NE = 2 * QO_TF_MAXEL + 1;
    for (int i = 0; i < NE; i++) ee[i] = 0.0;
    ee[0] = 1.0;
    for (int e = n_el - 1; e >= 0; e--) {
        double ne[2 * QO_TF_MAXEL + 1];
        for (int i = 0; i < NE; i++)
            ne[i] = fma(rec[e][6], ee[i], fma(rec[e][7], i >= 1 ? ee[i - 1] : 0.0, rec[e][8] * (i >= 2 ? ee[i - 2] : 0.0)));
        for (int i = 0; i < NE; i++) ee[i] = ne[i];
    }
    std::vector<double> na(NC), nb(NC), ndv(NC);
    for (int e = n_el - 1; e >= 0; e--) {
        const double *r = rec[e];
        const bool series = r[9] != 0.0;
        double *a = series ? p : q, *b = series ? q : p;
        for (int i = 0; i < NC; i++) {
            const double a1 = i >= 1 ? a[i - 1] : 0.0, a2 = i >= 2 ? a[i - 2] : 0.0, b1 = i >= 1 ? b[i - 1] : 0.0, b2 = i >= 2 ? b[i - 2] : 0.0;
            const double d1 = i >= 1 ? d[i - 1] : 0.0, d2 = i >= 2 ? d[i - 2] : 0.0;
            double v = fma(r[3], a[i], fma(r[4], a1, r[5] * a2));
            na[i] = fma(r[0], b[i], fma(r[1], b1, fma(r[2], b2, v)));
            nb[i] = fma(r[3], b[i], fma(r[4], b1, r[5] * b2));
            ndv[i] = fma(r[3], d[i], fma(r[4], d1, r[5] * d2));
        }
        for (int i = 0; i < NC; i++) { a[i] = na[i]; b[i] = nb[i]; d[i] = ndv[i]; }
    }
}

static cplx tf_horner_host(const double *c, int K, double x)
{
    const double y = -x * x;
    double re = c[2 * (K - 1)], im = c[2 * (K - 1) + 1];
    for (int k = K - 2; k >= 0; k--) { re = fma(re, y, c[2 * k]); im = fma(im, y, c[2 * k + 1]); }
    return cplx(re, im * x);
}

extern "C" int qo_tf_plan_check(const DevProg *hp, int mode_reduce_only, int precision, int generic, const double *f, int nf,
                                const unsigned char *mask, int *K, int *mode, double *wref, int *el0, int *n_el, int *cpl_op,
                                double *err, const char **reason)
{
    static const char *why = "";
    *reason = why;
#define QO_TF_NO(msg) do { *reason = msg; return 0; } while (0)
    const char *force = getenv("QO100NET_KERNEL");
    if (force && (strcmp(force, "interp") == 0 || strcmp(force, "ladder") == 0)) QO_TF_NO("QO100NET_KERNEL override");
    if (generic || !mode_reduce_only || precision != 64 || hp->need_gd || hp->need_s11) QO_TF_NO("not a reduce-only FP64 |S21| job");
    if (hp->nspec < 1 || hp->nspec > QO_LAD_NSPEC || hp->n_var > QO_MAX_VAR) QO_TF_NO("spec / variable count");
    for (int s = 0; s < hp->nspec; s++)
        if (hp->spec_kind[s] != SK_DEN2_MAX && hp->spec_kind[s] != SK_DEN2_MIN) QO_TF_NO("spec kind");
    int e0 = hp->op0;
    *cpl_op = -1;
    if (hp->n_ops > e0 && hp->opcode[e0] == OP_CPL) { *cpl_op = e0; e0++; }
    const int nl = hp->n_ops - e0;
    if (nl < 1 || nl > QO_TF_MAXEL) QO_TF_NO("element count");
    int deg = 0, has_d = 0;
    for (int e = 0; e < nl; e++) {
        const int op = hp->opcode[e0 + e], dg = qo_tf_degree(op);
        if (dg < 0) QO_TF_NO("non-lumped element");
        deg += dg;
        if (!(op == OP_SER_R || op == OP_SER_L || op == OP_SHUNT_C)) has_d = 1;     /* these have D == 1 exactly */
    }
    if (deg > 2 * QO_TF_MAXK - 1) QO_TF_NO("polynomial degree");
    const int Kk = deg / 2 + 1;
    *K = Kk; *el0 = e0; *n_el = nl;
    *mode = *cpl_op >= 0 ? QO_TF_CPL : has_d ? QO_TF_S21 : QO_TF_S21_NOD;

    const double two_pi = 6.283185307179586476925286766559;
    double fmin = f[0], fmax = f[0];
    for (int k = 1; k < nf; k++) { if (f[k] < fmin) fmin = f[k]; if (f[k] > fmax) fmax = f[k]; }
    const double wr = two_pi * sqrt(fmin * fmax);
    *wref = wr;
    const double zn = sqrt(hp->rs * hp->rl), zni = 1.0 / zn;

    /* nominal, and both all-at-one-end corners of the tolerance box.  First pass: can |D|^2 be evaluated as the
     * polynomial E(y) cut after K coefficients (dropped tail below 2e-12 of E everywhere on the grid)?  Second pass:
     * the device algorithm of the chosen mode against the per-element evaluation. */
    double trunc = 2e-12;
    if (getenv("QO100NET_TF_TRUNC")) trunc = atof(getenv("QO100NET_TF_TRUNC"));
    int emode = has_d && !getenv("QO100NET_TF_NO_E");
    double worst = 0.0;
    for (int pass = 0; pass < 2; pass++) {
        for (int corner = -1; corner <= 1; corner++) {
            double rec[QO_TF_MAXEL][QO_TF_REC];
            double nd[QO_TF_MAXEL][6];
            int ser[QO_TF_MAXEL];
            for (int e = 0; e < nl; e++) {
                double p[6];
                for (int k = 0; k < 6; k++) {
                    p[k] = hp->nom[e0 + e][k];
                    if (hp->tvar[e0 + e][k] >= 0) p[k] = qo_stream_apply(p[k], hp->ttol[e0 + e][k], (double)corner, hp->tmode[e0 + e][k]);
                }
                ser[e] = qo_tf_element(hp->opcode[e0 + e], p, wr, nd[e]);
                const double sc = ser[e] ? zni : zn;
                rec[e][0] = nd[e][0] * sc; rec[e][1] = nd[e][1] * sc; rec[e][2] = nd[e][2] * sc;
                rec[e][3] = nd[e][3]; rec[e][4] = nd[e][4]; rec[e][5] = nd[e][5];
                rec[e][6] = nd[e][3] * nd[e][3]; rec[e][7] = fma(2.0 * nd[e][3], nd[e][5], -nd[e][4] * nd[e][4]); rec[e][8] = nd[e][5] * nd[e][5];
                rec[e][9] = ser[e] ? 1.0 : 0.0;
            }
            double pp[2 * QO_TF_MAXK + 2], qq[2 * QO_TF_MAXK + 2], dd[2 * QO_TF_MAXK + 2], ee[2 * QO_TF_MAXEL + 1];
            tf_expand_host(rec, nl, hp->rl, zn, pp, qq, dd, ee);
            for (int i = 2 * Kk; i < 2 * QO_TF_MAXK + 2; i++)
                if (pp[i] != 0.0 || qq[i] != 0.0 || dd[i] != 0.0) QO_TF_NO("degree accounting");
            for (int k = 0; k < nf; k++) {
                const double x = two_pi * f[k] / wr, y = -x * x;
                if (pass == 0) {
                    if (!emode) break;
                    double tail = 0.0, full = 0.0, xp = 1.0;
                    for (int m = 0; m <= 2 * nl; m++) { if (m >= Kk) tail += fabs(ee[m]) * xp; full += ee[m] * (m & 1 ? -xp : xp); xp *= x * x; }
                    if (!(full > 0.0) || tail > trunc * full) emode = 0;
                    continue;
                }
                const cplx P_ = tf_horner_host(pp, Kk, x), Q_ = tf_horner_host(qq, Kk, x) * zni;
                double d2;
                if (emode) { d2 = ee[Kk - 1]; for (int m = Kk - 2; m >= 0; m--) d2 = fma(d2, y, ee[m]); }
                else d2 = std::norm(tf_horner_host(dd, Kk, x));
                if (!(d2 > 1e-70 && d2 < 1e70)) QO_TF_NO("|D|^2 leaves the range of the batched reciprocal");
                if (!mask[k]) continue;
                /* per-element evaluation, column vector from the load end */
                const cplx sj(0.0, x);
                cplx a(hp->rl, 0.0), b(1.0, 0.0);
                double dref = 1.0;
                for (int e = nl - 1; e >= 0; e--) {
                    const cplx N = nd[e][0] + sj * (nd[e][1] + sj * nd[e][2]), D = nd[e][3] + sj * (nd[e][4] + sj * nd[e][5]);
                    const cplx imm = N / D;
                    if (ser[e]) a += imm * b; else b += imm * a;
                    dref *= std::norm(D);
                }
                /* |den|^2 = |numerator polynomial|^2 / |D|^2 against |per-element chain|^2 */
                double rel;
                if (*cpl_op >= 0) {
                    /* P/D and Q/D as complex numbers (the coupler's row vector mixes them), plus the kernel's |D|^2 against the true one */
                    const cplx Dc = tf_horner_host(dd, Kk, x);
                    rel = 2.0 * (std::abs(P_ / Dc - a) + zn * std::abs(Q_ / Dc - b)) / (std::abs(a) + zn * std::abs(b)) + fabs(d2 / dref - 1.0);
                } else {
                    const double got = std::norm(P_ + hp->rs * Q_) / d2, ref = std::norm(a + hp->rs * b);
                    rel = fabs(got - ref) / ref;
                }
                if (!(rel == rel)) QO_TF_NO("self-check produced a NaN");
                if (rel > worst) worst = rel;
            }
        }
    }
    if (emode) *mode = *cpl_op >= 0 ? QO_TF_CPL_E : QO_TF_S21_E;
    *err = worst;
    double tol = 1e-10;
    const char *t = getenv("QO100NET_TF_TOL");
    if (t) tol = atof(t);
    if (worst > tol) QO_TF_NO("polynomial expansion is too ill-conditioned on this grid");
    *reason = "ok";
    return 1;
#undef QO_TF_NO
}
