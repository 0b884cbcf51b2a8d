/*
 * qo_s2p.c -- Touchstone v1 two-port files as network blocks, SURVEY row N3.
 *
 * Reference artefacts: the measured inductors the bias-network schematics pull in with
 *   <SPfile ... "11SQ39N.S2P" ... "polar" "linear" ...>   util/pa-bias-simulation/pa-bias-simulation.sch:39
 *   <SPfile ... "06HP47N.s2p" ... "polar" "linear" ...>   util/preamp-bias-simulation/preamp-bias-simulation.sch:32
 * (Coilcraft 1111SQ-39N, 659 points 10-3300 MHz, "# MHZ S MA R 50", 11SQ39N.S2P:1-5; 0603HP-47N, 189 points
 * 1-6000 MHz log, 06HP47N.s2p:1) and the driver measurement docs/pa-driver/pa_20W_vdd_32V_idq_180mA.s2p:5
 * ("# HZ S RI R 50", 501 points).  What Qucs' SPfile component does with them -- interpolate the data in
 * polar (|S|, unwrapped phase) or rectangular form, linearly in frequency, extrapolating the end segments
 * outside the measured range -- is restated here; the interpolated S is converted to an ABCD block at the file's
 * reference impedance and cascaded like any other element (kernel opcode OP_SBLOCK).
 *
 * Also here: the least-squares fit of the ESR/SRF inductor model Z = (R(f) + jwL) || 1/(jwCp),
 * R(f) = r0 + r1 sqrt(f), to such a block -- the evidence behind qo_net_add_parasitics' model.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "qo_internal.h"

static const double S2P_PI = 3.14159265358979323846;

qo_s2p *qo_s2p_alloc(int n)
{
    qo_s2p *b = (qo_s2p *)calloc(1, sizeof *b);
    if (!b) return NULL;
    b->n = n;
    b->z0 = 50.0;
    b->f = (double *)malloc((size_t)(n ? n : 1) * sizeof(double));
    b->s = (qo_c64 *)malloc((size_t)(n ? n : 1) * 4 * sizeof(qo_c64));
    if (!b->f || !b->s) { qo_s2p_free(b); return NULL; }
    return b;
}

qo_s2p *qo_s2p_clone(const qo_s2p *a)
{
    qo_s2p *b = qo_s2p_alloc(a->n);
    if (!b) return NULL;
    b->z0 = a->z0;
    memcpy(b->f, a->f, (size_t)a->n * sizeof(double));
    memcpy(b->s, a->s, (size_t)a->n * 4 * sizeof(qo_c64));
    return b;
}

void qo_s2p_free(qo_s2p *b)
{
    if (!b) return;
    free(b->f); free(b->s); free(b);
}

int qo_s2p_num_points(const qo_s2p *b) { return b ? b->n : QO_ERR_ARG; }
double qo_s2p_z0(const qo_s2p *b) { return b ? b->z0 : 0.0; }

int qo_s2p_get(const qo_s2p *b, double *f, qo_c64 *s11, qo_c64 *s21, qo_c64 *s12, qo_c64 *s22, int cap)
{
    if (!b) return QO_ERR_ARG;
    int n = b->n < cap ? b->n : cap;
    qo_c64 *dst[4] = { s11, s21, s12, s22 };
    for (int k = 0; k < n; k++) {
        if (f) f[k] = b->f[k];
        for (int j = 0; j < 4; j++) if (dst[j]) dst[j][k] = b->s[4 * (size_t)k + j];
    }
    return b->n;
}

static int s2p_check_grid(const double *f, int n)
{
    for (int k = 0; k < n; k++) {
        if (!(f[k] >= 0.0) || !isfinite(f[k])) return 0;
        if (k && !(f[k] > f[k - 1])) return 0;
    }
    return 1;
}

int qo_s2p_from_arrays(const double *f, int n, const qo_c64 *s11, const qo_c64 *s21, const qo_c64 *s12, const qo_c64 *s22,
                       double z0, qo_s2p **out)
{
    qo_clear_error();
    if (!f || n < 1 || !s11 || !s21 || !s12 || !s22 || !(z0 > 0) || !out) return QO_ERR_ARG;
    if (!s2p_check_grid(f, n)) { qo_set_error("block frequencies must be finite, non-negative and strictly increasing"); return QO_ERR_ARG; }
    qo_s2p *b = qo_s2p_alloc(n);
    if (!b) return QO_ERR_NOMEM;
    b->z0 = z0;
    for (int k = 0; k < n; k++) {
        b->f[k] = f[k];
        b->s[4 * (size_t)k + 0] = s11[k]; b->s[4 * (size_t)k + 1] = s21[k];
        b->s[4 * (size_t)k + 2] = s12[k]; b->s[4 * (size_t)k + 3] = s22[k];
    }
    *out = b;
    return QO_OK;
}

/* Touchstone v1, two ports: option line "# <HZ|KHZ|MHZ|GHZ> S <MA|DB|RI> R <z0>" (any order, case-insensitive,
 * defaults GHZ S MA R 50), '!' comments, nine numbers per point: f, S11, S21, S12, S22 as value pairs. */
int qo_s2p_load(const char *path, qo_s2p **out)
{
    qo_clear_error();
    if (!path || !out) return QO_ERR_ARG;
    size_t len;
    char *txt = qo_read_file(path, &len);
    if (!txt) return QO_ERR_IO;
    double fscale = 1e9, z0 = 50.0;
    int fmt = 0;    /* 0 MA, 1 DB, 2 RI */
    int cap = 256, n = 0, nv = 0, line = 0, rc = QO_OK, seen_opt = 0;
    double *vals = (double *)malloc((size_t)cap * 9 * sizeof(double));
    double cur[9];
    if (!vals) { free(txt); return QO_ERR_NOMEM; }
    char *save = NULL;
    for (char *ln = strtok_r(txt, "\n", &save); ln && rc == QO_OK; ln = strtok_r(NULL, "\n", &save)) {
        line++;
        char *bang = strchr(ln, '!');
        if (bang) *bang = '\0';
        while (*ln && isspace((unsigned char)*ln)) ln++;
        if (!*ln) continue;
        if (*ln == '#') {
            if (seen_opt) continue;           /* only the first option line counts */
            seen_opt = 1;
            char *sv2 = NULL;
            int want_r = 0;
            for (char *t = strtok_r(ln + 1, " \t\r", &sv2); t; t = strtok_r(NULL, " \t\r", &sv2)) {
                for (char *c = t; *c; c++) *c = (char)toupper((unsigned char)*c);
                if (want_r) { z0 = atof(t); want_r = 0; if (!(z0 > 0)) { qo_set_error("%s:%d: bad reference impedance", path, line); rc = QO_ERR_PARSE; } }
                else if (!strcmp(t, "HZ")) fscale = 1.0;
                else if (!strcmp(t, "KHZ")) fscale = 1e3;
                else if (!strcmp(t, "MHZ")) fscale = 1e6;
                else if (!strcmp(t, "GHZ")) fscale = 1e9;
                else if (!strcmp(t, "MA")) fmt = 0;
                else if (!strcmp(t, "DB")) fmt = 1;
                else if (!strcmp(t, "RI")) fmt = 2;
                else if (!strcmp(t, "R")) want_r = 1;
                else if (!strcmp(t, "S")) { }
                else if (!strcmp(t, "Y") || !strcmp(t, "Z") || !strcmp(t, "H") || !strcmp(t, "G")) {
                    qo_set_error("%s:%d: only S-parameter files are supported", path, line); rc = QO_ERR_UNSUPPORTED;
                }
            }
            continue;
        }
        /* data: numbers may wrap across lines; a point is complete after nine of them */
        char *p = ln;
        while (*p && rc == QO_OK) {
            char *end;
            double v = strtod(p, &end);
            if (end == p) {
                while (*p && isspace((unsigned char)*p)) p++;
                if (*p) { qo_set_error("%s:%d: unexpected text '%.16s'", path, line, p); rc = QO_ERR_PARSE; }
                break;
            }
            cur[nv++] = v;
            p = end;
            if (nv == 9) {
                if (n == cap) {
                    cap *= 2;
                    double *nvv = (double *)realloc(vals, (size_t)cap * 9 * sizeof(double));
                    if (!nvv) { rc = QO_ERR_NOMEM; break; }
                    vals = nvv;
                }
                memcpy(vals + 9 * (size_t)n, cur, sizeof cur);
                n++; nv = 0;
            }
        }
    }
    if (rc == QO_OK && nv != 0) { qo_set_error("%s: incomplete last data point (%d of 9 values); is this a 2-port file?", path, nv); rc = QO_ERR_PARSE; }
    if (rc == QO_OK && n == 0) { qo_set_error("%s: no data points", path); rc = QO_ERR_PARSE; }
    qo_s2p *b = NULL;
    if (rc == QO_OK) {
        b = qo_s2p_alloc(n);
        if (!b) rc = QO_ERR_NOMEM;
    }
    if (rc == QO_OK) {
        b->z0 = z0;
        for (int k = 0; k < n; k++) {
            const double *v = vals + 9 * (size_t)k;
            b->f[k] = v[0] * fscale;
            for (int j = 0; j < 4; j++) {
                const double a = v[1 + 2 * j], c = v[2 + 2 * j];
                qo_c64 z;
                if (fmt == 2) { z.re = a; z.im = c; }
                else {
                    const double mag = fmt == 1 ? pow(10.0, a / 20.0) : a, ph = c * S2P_PI / 180.0;
                    z.re = mag * cos(ph); z.im = mag * sin(ph);
                }
                b->s[4 * (size_t)k + j] = z;          /* file order S11 S21 S12 S22 == storage order */
            }
        }
        if (!s2p_check_grid(b->f, n)) { qo_set_error("%s: frequencies are not strictly increasing", path); rc = QO_ERR_PARSE; }
    }
    free(vals); free(txt);
    if (rc) { qo_s2p_free(b); return rc; }
    *out = b;
    return QO_OK;
}

/* S at frequency f: linear in f between the bracketing points, in rectangular form (polar == 0) or in |S| and
 * phase with each phase step taken along the shorter arc (polar != 0).  Outside the measured range the first /
 * last segment is EXTRAPOLATED linearly -- that, not holding the end value, is what Qucs' SPfile does: the
 * reference's own dataset util/pa-bias-simulation/pa-bias-simulation.dat (sweep 1 MHz - 10 GHz, inductor data
 * 10 MHz - 3.3 GHz) is reproduced to 1e-10 only with extrapolation (tests/test_nodal.py); holding the end
 * values is off by 0.5 % - 300 % above 3.3 GHz. */
void qo_s2p_eval(const qo_s2p *b, double f, int polar, qo_c64 s[4])
{
    const int n = b->n;
    if (n == 1) { memcpy(s, b->s, 4 * sizeof(qo_c64)); return; }
    int lo = 0, hi = n - 1;
    if (f <= b->f[0]) hi = 1;
    else if (f >= b->f[n - 1]) lo = n - 2;
    else while (hi - lo > 1) { int mid = (lo + hi) / 2; if (b->f[mid] <= f) lo = mid; else hi = mid; }
    const double t = (f - b->f[lo]) / (b->f[hi] - b->f[lo]);       /* < 0 or > 1 when extrapolating */
    for (int j = 0; j < 4; j++) {
        const qo_c64 a = b->s[4 * (size_t)lo + j], c = b->s[4 * (size_t)hi + j];
        if (!polar) {
            s[j].re = a.re + t * (c.re - a.re); s[j].im = a.im + t * (c.im - a.im);
        } else {
            const double ma = hypot(a.re, a.im), mc = hypot(c.re, c.im);
            const double pa = atan2(a.im, a.re);
            double dp = atan2(c.im, c.re) - pa;
            if (dp > S2P_PI) dp -= 2.0 * S2P_PI;
            if (dp < -S2P_PI) dp += 2.0 * S2P_PI;
            const double m = ma + t * (mc - ma), p = pa + t * dp;
            s[j].re = m * cos(p); s[j].im = m * sin(p);
        }
    }
}

int qo_s2p_interp(const qo_s2p *b, const double *f, int nf, int polar, qo_c64 *s11, qo_c64 *s21, qo_c64 *s12, qo_c64 *s22)
{
    qo_clear_error();
    if (!b || !f || nf <= 0) return QO_ERR_ARG;
    for (int k = 0; k < nf; k++) {
        qo_c64 s[4];
        qo_s2p_eval(b, f[k], polar, s);
        if (s11) s11[k] = s[0];
        if (s21) s21[k] = s[1];
        if (s12) s12[k] = s[2];
        if (s22) s22[k] = s[3];
    }
    return QO_OK;
}

static qo_c64 c_mul(qo_c64 a, qo_c64 b) { qo_c64 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static qo_c64 c_div(qo_c64 a, qo_c64 b)
{
    const double d = b.re * b.re + b.im * b.im;
    qo_c64 r = { (a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d };
    return r;
}

/* S (reference z0) -> ABCD; abcd = {A, B, C, D}.  Returns 0 when S21 == 0 (no transmission: no chain matrix). */
int qo_s_to_abcd(const qo_c64 s[4], double z0, qo_c64 abcd[4])
{
    const qo_c64 s11 = s[0], s21 = s[1], s12 = s[2], s22 = s[3];
    if (s21.re == 0.0 && s21.im == 0.0) return 0;
    const qo_c64 p = c_mul(s12, s21), two21 = { 2.0 * s21.re, 2.0 * s21.im };
    const qo_c64 a1 = { 1.0 + s11.re, s11.im }, a2 = { 1.0 - s11.re, -s11.im };
    const qo_c64 d1 = { 1.0 + s22.re, s22.im }, d2 = { 1.0 - s22.re, -s22.im };
    qo_c64 t;
    t = c_mul(a1, d2); t.re += p.re; t.im += p.im; abcd[0] = c_div(t, two21);
    t = c_mul(a1, d1); t.re -= p.re; t.im -= p.im; abcd[1] = c_div(t, two21); abcd[1].re *= z0; abcd[1].im *= z0;
    t = c_mul(a2, d2); t.re -= p.re; t.im -= p.im; abcd[2] = c_div(t, two21); abcd[2].re /= z0; abcd[2].im /= z0;
    t = c_mul(a2, d1); t.re += p.re; t.im += p.im; abcd[3] = c_div(t, two21);
    return 1;
}

/* ---- inductor-model fit ---------------------------------------------------
 * The block is read as a series element: its series impedance is the chain matrix' B.  With Y = 1/B,
 *   Im Y = w Cp - wL / (R^2 + w^2 L^2)  ~  w Cp - (1/L)(1/w)      (wL >> R)
 *   Re Y = R / (R^2 + w^2 L^2)          ~  R / (wL)^2
 * 1. self-resonance: the first inductive -> capacitive sign change of Im Y inside [fmin, fmax], if any;
 * 2. (Cp, 1/L) by linear least squares on Im Y over the points below 0.3 SRF (all points when no resonance is
 *    seen), each weighted by 1/|Y|^2; when a resonance was seen, Cp is re-derived from it: Cp = 1/(w_srf^2 L),
 *    which is far better conditioned than the small w Cp term at low frequency (the Coilcraft files carry
 *    three significant digits);
 * 3. R(f) = r0 + r1 sqrt(f) by linear least squares on Re Y (wL)^2 over the same points with wL > 5 R.
 * rms_rel = rms over the fitted points of |Z_model - Z_meas| / |Z_meas|. */
int qo_s2p_fit_inductor(const qo_s2p *b, double fmin, double fmax, double *L, double *r0, double *r1, double *cp, double *srf, double *rms_rel)
{
    qo_clear_error();
    if (!b || !L || !r0 || !r1 || !cp) return QO_ERR_ARG;
    const int n = b->n;
    double *w = (double *)malloc((size_t)n * 5 * sizeof(double));
    if (!w) return QO_ERR_NOMEM;
    double *yr = w + n, *yi = w + 2 * n, *zr = w + 3 * n, *zi = w + 4 * n;
    int m = 0;
    for (int k = 0; k < n; k++) {
        if (b->f[k] < fmin || b->f[k] > fmax || !(b->f[k] > 0)) continue;
        qo_c64 abcd[4];
        if (!qo_s_to_abcd(b->s + 4 * (size_t)k, b->z0, abcd)) continue;
        const qo_c64 one = { 1.0, 0.0 }, y = c_div(one, abcd[1]);
        w[m] = 2.0 * S2P_PI * b->f[k]; yr[m] = y.re; yi[m] = y.im; zr[m] = abcd[1].re; zi[m] = abcd[1].im;
        m++;
    }
    if (m < 4) { free(w); qo_set_error("fewer than 4 usable points in [%g, %g] Hz", fmin, fmax); return QO_ERR_RANGE; }
    double w_srf = 0.0;
    for (int k = 1; k < m; k++)
        if (yi[k - 1] < 0.0 && yi[k] >= 0.0) { w_srf = w[k - 1] + (w[k] - w[k - 1]) * (-yi[k - 1]) / (yi[k] - yi[k - 1]); break; }
    const double w_hi = w_srf > 0 ? 0.3 * w_srf : w[m - 1];
    int mfit = 0;
    double a11 = 0, a12 = 0, a22 = 0, b1 = 0, b2 = 0, Cp, invL;
    for (int k = 0; k < m && w[k] <= w_hi; k++) mfit++;
    if (mfit < 3) { free(w); qo_set_error("too few points below the self-resonance to fit"); return QO_ERR_RANGE; }
    if (w_srf > 0) {
        /* resonance known: Cp = 1/(w_srf^2 L), so Im Y = (1/L) (w / w_srf^2 - 1/w) has ONE unknown */
        for (int k = 0; k < mfit; k++) {
            const double wt = 1.0 / (yr[k] * yr[k] + yi[k] * yi[k]), x = w[k] / (w_srf * w_srf) - 1.0 / w[k];
            a11 += wt * x * x; b1 += wt * x * yi[k];
        }
        invL = b1 / a11;
        Cp = invL / (w_srf * w_srf);
    } else {
        for (int k = 0; k < mfit; k++) {
            const double wt = 1.0 / (yr[k] * yr[k] + yi[k] * yi[k]), x1 = w[k], x2 = -1.0 / w[k];
            a11 += wt * x1 * x1; a12 += wt * x1 * x2; a22 += wt * x2 * x2; b1 += wt * x1 * yi[k]; b2 += wt * x2 * yi[k];
        }
        const double det2 = a11 * a22 - a12 * a12;
        if (!(fabs(det2) > 0)) { free(w); qo_set_error("degenerate fit"); return QO_ERR_RANGE; }
        Cp = (b1 * a22 - b2 * a12) / det2;
        invL = (a11 * b2 - a12 * b1) / det2;
    }
    if (!(invL > 0)) { free(w); qo_set_error("the block does not look like a series inductor over this band"); return QO_ERR_RANGE; }
    const double Lf = 1.0 / invL;
    double det;
    a11 = a12 = a22 = b1 = b2 = 0;
    for (int k = 0; k < mfit; k++) {
        const double wl = w[k] * Lf, rk = yr[k] * wl * wl, x2 = sqrt(w[k] / (2.0 * S2P_PI));
        if (wl < 5.0 * fabs(rk)) continue;
        a11 += 1.0; a12 += x2; a22 += x2 * x2; b1 += rk; b2 += x2 * rk;
    }
    det = a11 * a22 - a12 * a12;
    double R0 = 0, R1 = 0;
    if (a11 >= 2 && fabs(det) > 0) { R0 = (b1 * a22 - b2 * a12) / det; R1 = (a11 * b2 - a12 * b1) / det; }
    double acc = 0;
    for (int k = 0; k < mfit; k++) {
        const double R = R0 + R1 * sqrt(w[k] / (2.0 * S2P_PI));
        const qo_c64 num = { R, w[k] * Lf }, den = { 1.0 - w[k] * w[k] * Lf * Cp, w[k] * R * Cp };
        const qo_c64 z = c_div(num, den);
        const double dr = z.re - zr[k], di = z.im - zi[k];
        acc += (dr * dr + di * di) / (zr[k] * zr[k] + zi[k] * zi[k]);
    }
    *L = Lf; *r0 = R0; *r1 = R1; *cp = Cp;
    if (srf) *srf = Cp > 0 ? 1.0 / (2.0 * S2P_PI * sqrt(Lf * Cp)) : INFINITY;
    if (rms_rel) *rms_rel = sqrt(acc / mfit);
    free(w);
    return QO_OK;
}
