/*
 * qo100ref.c -- CPU ORACLE (test infrastructure, NOT product code; see
 * qo100ref.h for the rules and for what is / is not pinned).
 *
 * Each function cites the reference artefact whose external-tool model it
 * restates (paths relative to /root/reference) and the SURVEY.md appendix
 * holding the equations that were verified against that artefact.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared qo100ref.c -lm
 */
#define _GNU_SOURCE
#include "qo100ref.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* tiny explicit complex type: every operation is spelled out so that  */
/* the arithmetic sequence is unambiguous (no C99 Annex-G surprises)   */
/* ------------------------------------------------------------------ */
typedef struct { double re, im; } cx;
static inline cx cx_mk(double a, double b) { cx z = { a, b }; return z; }
static inline cx cx_add(cx a, cx b) { return cx_mk(a.re + b.re, a.im + b.im); }
static inline cx cx_sub(cx a, cx b) { return cx_mk(a.re - b.re, a.im - b.im); }
static inline cx cx_mul(cx a, cx b) { return cx_mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline cx cx_scale(cx a, double s) { return cx_mk(a.re * s, a.im * s); }
static inline cx cx_div(cx a, cx b)
{
    double d = b.re * b.re + b.im * b.im;
    return cx_mk((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
static inline cx cx_inv(cx b)
{
    double d = b.re * b.re + b.im * b.im;
    return cx_mk(b.re / d, -b.im / d);
}
static inline double cx_abs2(cx a) { return a.re * a.re + a.im * a.im; }
static inline cx cx_sqrt(cx z)
{
    double m = hypot(z.re, z.im);
    double r = sqrt(0.5 * (m + fabs(z.re)));
    if (r == 0.0) return cx_mk(0, 0);
    if (z.re >= 0) return cx_mk(r, z.im / (2 * r));
    return cx_mk(fabs(z.im) / (2 * r), z.im >= 0 ? r : -r);
}
/* cosh / sinh of a complex argument */
static inline void cx_coshsinh(cx g, cx *ch, cx *sh)
{
    double c = cos(g.im), s = sin(g.im), chr = cosh(g.re), shr = sinh(g.re);
    *ch = cx_mk(chr * c, shr * s);
    *sh = cx_mk(shr * c, chr * s);
}

typedef struct { cx a, b, c, d; } m22;
static inline m22 m_ident(void) { m22 m = { { 1, 0 }, { 0, 0 }, { 0, 0 }, { 1, 0 } }; return m; }
static inline m22 m_mul(m22 x, m22 y)
{
    m22 r;
    r.a = cx_add(cx_mul(x.a, y.a), cx_mul(x.b, y.c));
    r.b = cx_add(cx_mul(x.a, y.b), cx_mul(x.b, y.d));
    r.c = cx_add(cx_mul(x.c, y.a), cx_mul(x.d, y.c));
    r.d = cx_add(cx_mul(x.c, y.b), cx_mul(x.d, y.d));
    return r;
}
static inline m22 m_series(cx z) { m22 m = m_ident(); m.b = z; return m; }
static inline m22 m_shunt(cx y) { m22 m = m_ident(); m.c = y; return m; }

static const double PI = 3.14159265358979323846;

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. SC'11) and the perturbation stream     */
/* SURVEY App. C; known-answer vectors are checked in tests/           */
/* ------------------------------------------------------------------ */
void ref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint64_t ref_bits53(uint64_t seed, uint64_t sample, uint32_t var)
{
    uint32_t ctr[4] = { (uint32_t)sample, (uint32_t)(sample >> 32), var >> 1, 0u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t o[4];
    ref_philox4x32_10(ctr, key, o);
    uint64_t w = (var & 1u) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
    return w >> 11;
}

double ref_uniform01(uint64_t seed, uint64_t sample, uint32_t var)
{
    return (double)ref_bits53(seed, sample, var) * 0x1p-53;
}

/* deterministic natural log built from + - * / fma only, so that the product's
 * host and device twins can reproduce it bit for bit (SURVEY §7 "hard parts") */
double ref_log_det(double x)
{
    uint64_t b;
    memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double m;
    memcpy(&m, &b, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, z, 1.0 / 21.0);
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    double q = fma(s * z, p, s);
    return fma((double)e, 0.6931471805599453, q + q);
}

/* Acklam's rational inverse-normal (|rel err| ~1e-9), Horner with fma */
double ref_norminv(double p)
{
    static const double a[6] = { -3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00 };
    static const double b[5] = { -5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                 6.680131188771972e+01, -1.328068155288572e+01 };
    static const double c[6] = { -7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                 -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00 };
    static const double d[4] = { 7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                 3.754408661907416e+00 };
    const double plow = 0.02425;
    if (p < plow || p > 1.0 - plow) {
        int upper = p > 0.5;
        double pp = upper ? 1.0 - p : p;
        double q = sqrt(-2.0 * ref_log_det(pp));
        double num = c[0];
        num = fma(num, q, c[1]); num = fma(num, q, c[2]); num = fma(num, q, c[3]);
        num = fma(num, q, c[4]); num = fma(num, q, c[5]);
        double den = d[0];
        den = fma(den, q, d[1]); den = fma(den, q, d[2]); den = fma(den, q, d[3]);
        den = fma(den, q, 1.0);
        double x = num / den;
        return upper ? -x : x;
    }
    double q = p - 0.5, r = q * q;
    double num = a[0];
    num = fma(num, r, a[1]); num = fma(num, r, a[2]); num = fma(num, r, a[3]);
    num = fma(num, r, a[4]); num = fma(num, r, a[5]);
    double den = b[0];
    den = fma(den, r, b[1]); den = fma(den, r, b[2]); den = fma(den, r, b[3]);
    den = fma(den, r, b[4]); den = fma(den, r, 1.0);
    return (num * q) / den;
}

double ref_variate(uint64_t seed, uint64_t sample, uint32_t var, int dist)
{
    uint64_t k = ref_bits53(seed, sample, var);
    if (dist == REF_DIST_GAUSS3S) {
        double p = fma((double)k, 0x1p-53, 0x1p-54);
        double z = ref_norminv(p);
        if (z > 3.0) z = 3.0;
        if (z < -3.0) z = -3.0;
        return z / 3.0;
    }
    double u = (double)k * 0x1p-53;
    return fma(2.0, u, -1.0);
}

double ref_perturb_factor(uint64_t seed, uint64_t sample, uint32_t var, int dist, double tol)
{
    return fma(tol, ref_variate(seed, sample, var, dist), 1.0);
}

/* out[n_samples][n_var] = ref_perturb_factor(seed, sample_offset + s, v, dist, tol): the whole stream in one call
 * (the GPU bit-exactness test compares 1e6 factors; SURVEY App. C "test = memcmp of 1e6 factors") */
void ref_perturb_factors(uint64_t seed, uint64_t sample_offset, uint64_t n_samples, int n_var, int dist, double tol, double *out)
{
    for (uint64_t s = 0; s < n_samples; s++)
        for (int v = 0; v < n_var; v++)
            out[s * (uint64_t)n_var + (uint64_t)v] = ref_perturb_factor(seed, sample_offset + s, (uint32_t)v, dist, tol);
}

void ref_perturb(const ref_elem *e, int n, const ref_mc_cfg *cfg, uint64_t sample, ref_elem *out)
{
    memcpy(out, e, (size_t)n * sizeof(ref_elem));
    for (int i = 0; i < cfg->n_tol; i++) {
        const ref_tol *t = &cfg->tol[i];
        if (t->elem < 0 || t->elem >= n || t->param < 0 || t->param >= 6) continue;
        double x = ref_variate(cfg->seed, sample, (uint32_t)t->var, cfg->dist);
        double nom = e[t->elem].p[t->param];
        if (t->mode == REF_TOL_ABS) out[t->elem].p[t->param] = fma(t->tol, x, nom);
        else out[t->elem].p[t->param] = nom * fma(t->tol, x, 1.0);
    }
}

/* ------------------------------------------------------------------ */
/* grids: util/pa-lpf-simulation/pa-lpf-simulation.sch:59 (.SP lin)     */
/* f_k = f0 + k*((f1-f0)/(n-1)), bit-exact vs .dat:6-5005               */
/* ------------------------------------------------------------------ */
void ref_grid_lin(double f0, double f1, int n, double *f)
{
    double step = n > 1 ? (f1 - f0) / (double)(n - 1) : 0.0;
    for (int k = 0; k < n; k++) f[k] = f0 + (double)k * step;
}
void ref_grid_log(double f0, double f1, int n, double *f)
{
    double l0 = log(f0), step = n > 1 ? (log(f1) - log(f0)) / (double)(n - 1) : 0.0;
    for (int k = 0; k < n; k++) f[k] = exp(l0 + (double)k * step);
    if (n > 0) f[0] = f0;
    if (n > 1) f[n - 1] = f1;
}

/* ------------------------------------------------------------------ */
/* synthesis: pcb/generic-filter (README.md:13, "up to 11th order       */
/* Butterworth/Chebyshev"; qo-100-generic-filter.sch:1450-1488,         */
/* 1703-1995: 6 series + 5 shunt branches, series first). SURVEY B.3.   */
/* ------------------------------------------------------------------ */
int ref_cheby_g(int n, double ripple_db, double *g)
{
    if (n < 1 || ripple_db <= 0) return -1;
    double beta = log(1.0 / tanh(ripple_db * log(10.0) / 40.0));
    double gam = sinh(beta / (2.0 * n));
    double ak_prev = 0, bk_prev = 0;
    for (int k = 1; k <= n; k++) {
        double ak = sin((2.0 * k - 1.0) * PI / (2.0 * n));
        double sk = sin(k * PI / n);
        double bk = gam * gam + sk * sk;
        if (k == 1) g[0] = 2.0 * ak / gam;
        else g[k - 1] = 4.0 * ak_prev * ak / (bk_prev * g[k - 2]);
        ak_prev = ak; bk_prev = bk;
    }
    return 0;
}
int ref_butter_g(int n, double *g)
{
    if (n < 1) return -1;
    for (int k = 1; k <= n; k++) g[k - 1] = 2.0 * sin((2.0 * k - 1.0) * PI / (2.0 * n));
    return 0;
}
int ref_ladder_lpf(const double *g, int n, double fc, double z0, int series_first, ref_elem *out)
{
    double wc = 2.0 * PI * fc;
    for (int k = 0; k < n; k++) {
        int series = series_first ? (k % 2 == 0) : (k % 2 == 1);
        memset(&out[k], 0, sizeof(ref_elem));
        if (series) { out[k].kind = REF_SER_L; out[k].p[0] = g[k] * z0 / wc; }
        else { out[k].kind = REF_SHUNT_C; out[k].p[0] = g[k] / (z0 * wc); }
    }
    return n;
}
/* L: R = wc*L/Q, Cp = 1/((2pi*mL*fc)^2 L); C: ESR, Ls = 1/((2pi*mC*fc)^2 C).
 * Parasitic model evidenced by util/pa-bias-simulation/pa-bias-simulation.sch:19-34
 * (C + series R pairs) and the Coilcraft .s2p SRFs. */
void ref_add_parasitics(ref_elem *e, int n, double fc, double q_l, double srf_l_mult, double esr_c, double srf_c_mult)
{
    double wc = 2.0 * PI * fc;
    for (int k = 0; k < n; k++) {
        if (e[k].kind == REF_SER_L || e[k].kind == REF_SHUNT_L) {
            double L = e[k].p[0], ws = 2.0 * PI * srf_l_mult * fc;
            e[k].p[1] = wc * L / q_l;
            e[k].p[2] = 1.0 / (ws * ws * L);
        } else if (e[k].kind == REF_SER_C || e[k].kind == REF_SHUNT_C) {
            double C = e[k].p[0], ws = 2.0 * PI * srf_c_mult * fc;
            e[k].p[1] = esr_c;
            e[k].p[2] = 1.0 / (ws * ws * C);
        }
    }
}

/* ------------------------------------------------------------------ */
/* Qucs 0.0.19 microstrip models -- SURVEY App. A (verified <=3.8e-12   */
/* vs util/pa-lpf-simulation/pa-lpf-simulation.dat)                     */
/* ------------------------------------------------------------------ */
static const double C0 = 299792458.0;
static const double MU0 = 12.566370614e-7;
static const double ZF0 = 376.73031346958504364963;
#define M_E_ 2.7182818284590452354

typedef struct { double er, h, t, tand, rho, D; } subst_t;

static double ms_Zh(double x)
{
    double F = 6.0 + (2.0 * PI - 6.0) * exp(-pow(30.666 / x, 0.7528));
    return ZF0 / (2.0 * PI) * log(F / x + sqrt(1.0 + (2.0 / x) * (2.0 / x)));
}
static double ms_ee(double x, double er)
{
    double x2 = x * x, x4 = x2 * x2;
    double a = 1.0 + log((x4 + (x / 52.0) * (x / 52.0)) / (x4 + 0.432)) / 49.0 + log(1.0 + pow(x / 18.1, 3.0)) / 18.7;
    double b = 0.564 * pow((er - 0.9) / (er + 3.0), 0.053);
    return (er + 1.0) / 2.0 + (er - 1.0) / 2.0 * pow(1.0 + 10.0 / x, -a * b);
}
/* A.1 -- MLIN "Hammerstad" quasi-static (pa-lpf-simulation.sch:21 "Hammerstad") */
void ref_ms_quasi(double W, double h, double t, double er, double *Z, double *E, double *Weff)
{
    double u = W / h, du1 = 0.0, dur = 0.0;
    if (t > 0.0) {
        double tau = t / h;
        double cth = 1.0 / tanh(sqrt(6.517 * u));
        du1 = (tau / PI) * log(1.0 + 4.0 * M_E_ / (tau * cth * cth));
        dur = du1 * (1.0 + 1.0 / cosh(sqrt(er - 1.0))) / 2.0;
    }
    double u1 = u + du1, ur = u + dur;
    double zr = ms_Zh(ur), z1 = ms_Zh(u1), eps = ms_ee(ur, er);
    *Z = zr / sqrt(eps);
    *E = eps * (z1 / zr) * (z1 / zr);
    *Weff = ur * h;
}
/* A.2 -- "Kirschning" dispersion (pa-lpf-simulation.sch:21 "Kirschning") */
void ref_ms_disp(double W, double h, double er, double Z, double E, double f, double *Zf, double *Ef)
{
    double u = W / h, fn = f * h / 1e6;
    double P1 = 0.27488 + (0.6315 + 0.525 / pow(1.0 + 0.0157 * fn, 20.0)) * u - 0.065683 * exp(-8.7513 * u);
    double P2 = 0.33622 * (1.0 - exp(-0.03442 * er));
    double P3 = 0.0363 * exp(-4.6 * u) * (1.0 - exp(-pow(fn / 38.7, 4.97)));
    double P4 = 1.0 + 2.751 * (1.0 - exp(-pow(er / 15.916, 8.0)));
    double Pf = P1 * P2 * pow((P3 * P4 + 0.1844) * fn, 1.5763);
    double ef = er - (er - E) / (1.0 + Pf);
    double R1 = 0.03891 * pow(er, 1.4);
    double R2 = 0.267 * pow(u, 7.0);
    double R3 = 4.766 * exp(-3.228 * pow(u, 0.641));
    double R4 = 0.016 + pow(0.0514 * er, 4.524);
    double R5 = pow(fn / 28.843, 12.0);
    double R6 = 22.20 * pow(u, 1.92);
    double R7 = 1.206 - 0.3144 * exp(-R1) * (1.0 - exp(-R2));
    double R8 = 1.0 + 1.275 * (1.0 - exp(-0.004625 * R3 * pow(er, 1.674) * pow(fn / 18.365, 2.745)));
    double em1_6 = pow(er - 1.0, 6.0);
    double R9 = 5.086 * R4 * R5 / (0.3838 + 0.386 * R4) * exp(-R6) / (1.0 + 1.2992 * R5) * em1_6 / (1.0 + 10.0 * em1_6);
    double R10 = 0.00044 * pow(er, 2.136) + 0.0184;
    double t6 = pow(fn / 19.47, 6.0);
    double R11 = t6 / (1.0 + 0.0962 * t6);
    double R12 = 1.0 / (1.0 + 0.00245 * u * u);
    double R13 = 0.9408 * pow(ef, R8) - 0.9603;
    double R14 = (0.9408 - R9) * pow(E, R8) - 0.9603;
    double R15 = 0.707 * R10 * pow(fn / 12.3, 1.097);
    double R16 = 1.0 + 0.0503 * er * er * R11 * (1.0 - exp(-pow(u / 15.0, 6.0)));
    double R17 = R7 * (1.0 - 1.1241 * R12 / R16 * exp(-0.026 * pow(fn, 1.15656) - R15));
    *Zf = Z * pow(R13 / R14, R17);
    *Ef = ef;
}
/* A.3 + A.4 -- MLIN(W, L, f): pa-lpf-simulation.sch:21,24-28,33-37,42-43,46-48,51-53 */
static m22 ms_mlin(const subst_t *s, double W, double L, double f)
{
    double Z, E, Weff, Zf, Ef;
    ref_ms_quasi(W, s->h, s->t, s->er, &Z, &E, &Weff);
    ref_ms_disp(W, s->h, s->er, Z, E, f, &Zf, &Ef);
    double Rs = sqrt(PI * f * MU0 * s->rho);
    double delta = s->rho / Rs;
    double Ki = exp(-1.2 * pow(Z / ZF0, 0.7));
    double dd = s->D / delta;
    double Kr = 1.0 + (2.0 / PI) * atan(1.4 * dd * dd);
    double ac = Rs / (Z * W) * Ki * Kr;
    double ad = PI * s->er / (s->er - 1.0) * (E - 1.0) / sqrt(E) * s->tand * f / C0;
    cx g = cx_mk((ac + ad) * L, 2.0 * PI * f * sqrt(Ef) / C0 * L);
    cx ch, sh;
    cx_coshsinh(g, &ch, &sh);
    m22 m;
    m.a = ch; m.d = ch;
    m.b = cx_scale(sh, Zf);
    m.c = cx_scale(sh, 1.0 / Zf);
    return m;
}
/* A.5 -- MCORN: pa-lpf-simulation.sch:22-23,29-32,38-41,44-45 */
static m22 ms_mcorn(const subst_t *s, double W, double f)
{
    double er = s->er, h = s->h;
    double CpF = W * ((10.35 * er + 2.5) * W / h + 2.6 * er + 5.64);
    double LnH = 220.0 * h * (1.0 - 1.35 * exp(-0.18 * pow(W / h, 1.39)));
    cx z21 = cx_mk(0.0, -0.5e12 / (PI * f * CpF));
    cx z11 = cx_mk(0.0, 2e-9 * PI * f * LnH + z21.im);
    m22 m;
    m.a = cx_div(z11, z21);
    m.d = m.a;
    m.b = cx_div(cx_sub(cx_mul(z11, z11), cx_mul(z21, z21)), z21);
    m.c = cx_inv(z21);
    return m;
}
/* A.5 -- MOPEN: pa-lpf-simulation.sch:56-57 ; returns Y = j*w*C_end */
static cx ms_mopen(const subst_t *s, double W, double f)
{
    double Z, E, Weff, Zf, Ef;
    double er = s->er, h = s->h;
    ref_ms_quasi(W, h, s->t, er, &Z, &E, &Weff);
    ref_ms_disp(Weff, h, er, Z, E, f, &Zf, &Ef);
    double w = W / h;
    double Q6 = pow(Ef, 0.81), Q7 = pow(w, 0.8544);
    double Q1 = 0.434907 * (Q6 + 0.26) / (Q6 - 0.189) * (Q7 + 0.236) / (Q7 + 0.87);
    double Q2 = pow(w, 0.371) / (2.358 * er + 1.0) + 1.0;
    double Q3 = atan(0.084 * pow(w, 1.9413 / Q2)) * 0.5274 / pow(Ef, 0.9236) + 1.0;
    double Q4 = 0.0377 * (6.0 - 5.0 * exp(0.036 * (1.0 - er))) * atan(0.067 * pow(w, 1.456)) + 1.0;
    double Q5 = 1.0 - 0.218 * exp(-7.5 * w);
    double dl = Q1 * Q3 * Q5 / Q4 * h;
    return cx_mk(0.0, 2.0 * PI * f * dl * sqrt(Ef) / (C0 * Zf));
}
/* A.5 -- MTEE: pa-lpf-simulation.sch:54-55 */
typedef struct { double La, Lb, L2, Ta2, Tb2, Bt; } tee_t;
static void ms_mtee(const subst_t *s, double Wa, double Wb, double W2, double f, tee_t *o)
{
    double er = s->er, h = s->h;
    double Za, Ea, Zb, Eb, Z2, E2, We, Zla, Era, Zlb, Erb, Zl2, Er2;
    ref_ms_quasi(Wa, h, s->t, er, &Za, &Ea, &We); ref_ms_disp(Wa, h, er, Za, Ea, f, &Zla, &Era);
    ref_ms_quasi(Wb, h, s->t, er, &Zb, &Eb, &We); ref_ms_disp(Wb, h, er, Zb, Eb, f, &Zlb, &Erb);
    ref_ms_quasi(W2, h, s->t, er, &Z2, &E2, &We); ref_ms_disp(W2, h, er, Z2, E2, f, &Zl2, &Er2);
    double Da = ZF0 / Zla * h / sqrt(Era), Db = ZF0 / Zlb * h / sqrt(Erb), D2 = ZF0 / Zl2 * h / sqrt(Er2);
    double fpa = 0.4e6 * Zla / h, fpb = 0.4e6 * Zlb / h;
    double lda = C0 / sqrt(Era) / f, ldb = C0 / sqrt(Erb) / f;
    double da = 0.055 * D2 * Zla / Zl2 * (1.0 - 2.0 * Zla / Zl2 * (f / fpa) * (f / fpa));
    double db = 0.055 * D2 * Zlb / Zl2 * (1.0 - 2.0 * Zlb / Zl2 * (f / fpb) * (f / fpb));
    o->La = 0.5 * W2 - da;
    o->Lb = 0.5 * W2 - db;
    double r = sqrt(Zla * Zlb) / Zl2;
    double q = f * f / (fpa * fpb);
    double d2 = sqrt(Da * Db) * (0.5 - r * (0.05 + 0.7 * exp(-1.6 * r) + 0.25 * r * q - 0.17 * log(r)));
    o->L2 = 0.5 * (Wa > Wb ? Wa : Wb) - d2;
    double ta = 1.0 - PI * (f / fpa) * (f / fpa) * ((Zla / Zl2) * (Zla / Zl2) / 12.0 + (0.5 - d2 / Da) * (0.5 - d2 / Da));
    double tb = 1.0 - PI * (f / fpb) * (f / fpb) * ((Zlb / Zl2) * (Zlb / Zl2) / 12.0 + (0.5 - d2 / Db) * (0.5 - d2 / Db));
    if (ta < 1e-18) ta = 1e-18;
    if (tb < 1e-18) tb = 1e-18;
    o->Ta2 = ta; o->Tb2 = tb;
    o->Bt = 5.5 * sqrt(Da * Db / (lda * ldb)) * (er + 2.0) / er / Zl2 / sqrt(ta * tb) * sqrt(da * db) / D2 *
            (1.0 + 0.9 * log(r) + 4.5 * r * q - 4.4 * exp(-1.3 * r) - 20.0 * (Zl2 / ZF0) * (Zl2 / ZF0));
}

/* ------------------------------------------------------------------ */
/* ideal coupled line, through path -- SURVEY B.4;                      */
/* util/directional-couplers/dir_cpl_*.trc:18-20 (Z0e, Z0o, Ang_l@Freq) */
/* ------------------------------------------------------------------ */
static m22 cpl_thru(double z0e, double z0o, double ang_e, double ang_o, double f0, double zt, double f)
{
    double te = ang_e * (PI / 180.0) * f / f0, to = ang_o * (PI / 180.0) * f / f0;
    /* even / odd mode lines seen in a Zt system */
    double ce = cos(te), se = sin(te), co = cos(to), so = sin(to);
    cx dene = cx_mk(2.0 * ce, z0e * se / zt + se * zt / z0e);
    cx deno = cx_mk(2.0 * co, z0o * so / zt + so * zt / z0o);
    cx ge = cx_div(cx_mk(0.0, z0e * se / zt - se * zt / z0e), dene);
    cx go = cx_div(cx_mk(0.0, z0o * so / zt - so * zt / z0o), deno);
    cx te_ = cx_div(cx_mk(2.0, 0.0), dene);
    cx to_ = cx_div(cx_mk(2.0, 0.0), deno);
    cx s11 = cx_scale(cx_add(ge, go), 0.5);
    cx s21 = cx_scale(cx_add(te_, to_), 0.5);
    /* symmetric reciprocal S -> ABCD at reference Zt */
    cx one = cx_mk(1.0, 0.0);
    cx p = cx_add(one, s11), m = cx_sub(one, s11);
    cx s21sq = cx_mul(s21, s21);
    cx two_s21 = cx_scale(s21, 2.0);
    m22 r;
    r.a = cx_div(cx_add(cx_mul(p, m), s21sq), two_s21);
    r.d = r.a;
    r.b = cx_scale(cx_div(cx_sub(cx_mul(p, p), s21sq), two_s21), zt);
    r.c = cx_scale(cx_div(cx_sub(cx_mul(m, m), s21sq), two_s21), 1.0 / zt);
    return r;
}

/* ------------------------------------------------------------------ */
/* element immittances -- SURVEY B.1 (rf-tools branch grammar,          */
/* util/if-bandpass-filter/schematic.svg:11-166 symbol defs)            */
/* ------------------------------------------------------------------ */
static cx z_ind(double L, double R, double Cp, double w)
{
    /* Z = (R + jwL) || 1/(jwCp) = (R + jwL) / (1 - w^2 L Cp + j w R Cp) */
    return cx_div(cx_mk(R, w * L), cx_mk(1.0 - w * w * L * Cp, w * R * Cp));
}
static cx z_cap(double C, double R, double Ls, double w)
{
    /* Z = R + jwLs + 1/(jwC) */
    return cx_mk(R, w * Ls - 1.0 / (w * C));
}

/* evaluate the whole cascade at one frequency: ABCD, left (source) to right (load) */
/* ------------------------------------------------------------------ */
/* Measured two-port blocks (Qucs SPfile): util/pa-bias-simulation/     */
/* pa-bias-simulation.sch:39, util/preamp-bias-simulation/...sch:32     */
/* ("polar" "linear").  Blocks are registered by index before a sweep.  */
/* ------------------------------------------------------------------ */
#define REF_MAX_BLK 8
static struct { int n; double z0; double *f; double *s; } g_blk[REF_MAX_BLK];

void ref_sblock_clear(void)
{
    for (int i = 0; i < REF_MAX_BLK; i++) { free(g_blk[i].f); free(g_blk[i].s); memset(&g_blk[i], 0, sizeof g_blk[i]); }
}

/* s: n x 8 doubles = (re, im) of S11, S21, S12, S22 per point */
int ref_sblock_register(int idx, const double *f, const double *s, int n, double z0)
{
    if (idx < 0 || idx >= REF_MAX_BLK || n < 1 || !(z0 > 0)) return -1;
    free(g_blk[idx].f); free(g_blk[idx].s);
    g_blk[idx].f = (double *)malloc((size_t)n * sizeof(double));
    g_blk[idx].s = (double *)malloc((size_t)n * 8 * sizeof(double));
    if (!g_blk[idx].f || !g_blk[idx].s) return -5;
    memcpy(g_blk[idx].f, f, (size_t)n * sizeof(double));
    memcpy(g_blk[idx].s, s, (size_t)n * 8 * sizeof(double));
    g_blk[idx].n = n; g_blk[idx].z0 = z0;
    return 0;
}

/* linear interpolation in frequency of one S entry, polar (|S| and phase along the shorter arc) or rectangular;
 * outside the measured range the end segment is extrapolated (what reproduces pa-bias-simulation.dat) */
static cx blk_entry(int idx, int j, double f, int polar)
{
    const int n = g_blk[idx].n;
    const double *F = g_blk[idx].f, *S = g_blk[idx].s;
    if (n == 1) return cx_mk(S[2 * j], S[2 * j + 1]);
    int k = 0;
    while (k + 2 < n && F[k + 1] <= f) k++;
    const cx a = cx_mk(S[8 * (size_t)k + 2 * j], S[8 * (size_t)k + 2 * j + 1]);
    const cx b = cx_mk(S[8 * (size_t)(k + 1) + 2 * j], S[8 * (size_t)(k + 1) + 2 * j + 1]);
    const double t = (f - F[k]) / (F[k + 1] - F[k]);
    if (!polar) return cx_mk(a.re + t * (b.re - a.re), a.im + t * (b.im - a.im));
    double pa = atan2(a.im, a.re), dp = atan2(b.im, b.re) - pa;
    if (dp > PI) dp -= 2.0 * PI;
    if (dp < -PI) dp += 2.0 * PI;
    const double ma = sqrt(cx_abs2(a)), mb = sqrt(cx_abs2(b));
    const double m = ma + t * (mb - ma), ph = pa + t * dp;
    return cx_mk(m * cos(ph), m * sin(ph));
}

int ref_sblock_s(int idx, double f, int polar, double s[8])
{
    if (idx < 0 || idx >= REF_MAX_BLK || g_blk[idx].n == 0) return -1;
    for (int j = 0; j < 4; j++) { cx v = blk_entry(idx, j, f, polar); s[2 * j] = v.re; s[2 * j + 1] = v.im; }
    return 0;
}

/* S (reference z0) -> chain matrix; *det = S12 / S21 */
static int blk_abcd(int idx, double f, int polar, m22 *M, cx *det)
{
    if (idx < 0 || idx >= REF_MAX_BLK || g_blk[idx].n == 0) return -1;
    const double z0 = g_blk[idx].z0;
    const cx s11 = blk_entry(idx, 0, f, polar), s21 = blk_entry(idx, 1, f, polar);
    const cx s12 = blk_entry(idx, 2, f, polar), s22 = blk_entry(idx, 3, f, polar);
    if (s21.re == 0.0 && s21.im == 0.0) return -9;
    const cx one = cx_mk(1, 0), p = cx_mul(s12, s21), d2 = cx_scale(s21, 2.0);
    M->a = cx_div(cx_add(cx_mul(cx_add(one, s11), cx_sub(one, s22)), p), d2);
    M->b = cx_scale(cx_div(cx_sub(cx_mul(cx_add(one, s11), cx_add(one, s22)), p), d2), z0);
    M->c = cx_scale(cx_div(cx_sub(cx_mul(cx_sub(one, s11), cx_sub(one, s22)), p), d2), 1.0 / z0);
    M->d = cx_div(cx_add(cx_mul(cx_sub(one, s11), cx_add(one, s22)), p), d2);
    *det = cx_div(s12, s21);
    return 0;
}

static int eval_abcd_det(const ref_elem *e, int n, double f, m22 *out, cx *det_out)
{
    cx det = cx_mk(1, 0);
    double w = 2.0 * PI * f;
    m22 M = m_ident(), Mmain = m_ident();
    subst_t sub = { 0, 0, 0, 0, 0, 0 };
    int have_sub = 0, in_side = 0;
    tee_t tee = { 0 };
    double teeWa = 0, teeWb = 0;
    for (int i = 0; i < n; i++) {
        const double *p = e[i].p;
        switch (e[i].kind) {
        case REF_SER_R: M = m_mul(M, m_series(cx_mk(p[0], 0))); break;
        case REF_SHUNT_R: M = m_mul(M, m_shunt(cx_mk(1.0 / p[0], 0))); break;
        case REF_SER_L: M = m_mul(M, m_series(z_ind(p[0], p[1], p[2], w))); break;
        case REF_SHUNT_L: M = m_mul(M, m_shunt(cx_inv(z_ind(p[0], p[1], p[2], w)))); break;
        case REF_SER_C: M = m_mul(M, m_series(z_cap(p[0], p[1], p[2], w))); break;
        case REF_SHUNT_C: M = m_mul(M, m_shunt(cx_inv(z_cap(p[0], p[1], p[2], w)))); break;
        case REF_SER_LC_SER: M = m_mul(M, m_series(cx_mk(0, w * p[0] - 1.0 / (w * p[1])))); break;
        case REF_SER_LC_PAR: M = m_mul(M, m_series(cx_inv(cx_mk(0, w * p[1] - 1.0 / (w * p[0]))))); break;
        case REF_SHUNT_LC_SER: M = m_mul(M, m_shunt(cx_inv(cx_mk(0, w * p[0] - 1.0 / (w * p[1]))))); break;
        case REF_SHUNT_LC_PAR: M = m_mul(M, m_shunt(cx_mk(0, w * p[1] - 1.0 / (w * p[0])))); break;
        case REF_TLINE: {
            double th = p[1] * (PI / 180.0) * f / p[2];
            m22 t;
            t.a = cx_mk(cos(th), 0); t.d = t.a;
            t.b = cx_mk(0, p[0] * sin(th));
            t.c = cx_mk(0, sin(th) / p[0]);
            M = m_mul(M, t);
            break;
        }
        case REF_CPL_THRU: M = m_mul(M, cpl_thru(p[0], p[1], p[2], p[3], p[4], p[5], f)); break;
        case REF_CPL_MS: {        /* physical coupled microstrip on the current substrate: QucsTranscalc analysis, then B.4 */
            if (!have_sub) return -4;
            double z0e, z0o, ae, ao;
            ref_cpl_analyze(p[0], p[1], sub.h, sub.t, sub.er, p[3], p[4], p[2], &z0e, &z0o, &ae, &ao);
            M = m_mul(M, cpl_thru(z0e, z0o, ae, ao, p[4], p[5], f));
            break;
        }
        case REF_SBLOCK: {
            m22 B;
            cx bd;
            int rc = blk_abcd((int)p[0], f, p[1] != 0.0, &B, &bd);
            if (rc) return rc;
            M = m_mul(M, B);
            det = cx_mul(det, bd);
            break;
        }
        case REF_SUBST:
            sub.er = p[0]; sub.h = p[1]; sub.t = p[2]; sub.tand = p[3]; sub.rho = p[4]; sub.D = p[5];
            have_sub = 1;
            break;
        case REF_MLIN:
            if (!have_sub) return -4;
            M = m_mul(M, ms_mlin(&sub, p[0], p[1], f));
            break;
        case REF_MCORN:
            if (!have_sub) return -4;
            M = m_mul(M, ms_mcorn(&sub, p[0], f));
            break;
        case REF_MTEE:
            if (!have_sub || in_side) return -4;
            ms_mtee(&sub, p[0], p[1], p[2], f, &tee);
            teeWa = p[0]; teeWb = p[1];
            Mmain = M;
            M = ms_mlin(&sub, p[2], tee.L2, f);   /* arm 2, junction -> outward */
            in_side = 1;
            break;
        case REF_MOPEN: {
            if (!have_sub || !in_side) return -4;
            cx yo = ms_mopen(&sub, p[0], f);
            cx yin = cx_div(cx_add(M.c, cx_mul(M.d, yo)), cx_add(M.a, cx_mul(M.b, yo)));
            double sa = sqrt(tee.Ta2), sb = sqrt(tee.Tb2);
            m22 Ta = m_ident(), Tb = m_ident();
            Ta.a = cx_mk(1.0 / sa, 0); Ta.d = cx_mk(sa, 0);
            Tb.a = cx_mk(sb, 0); Tb.d = cx_mk(1.0 / sb, 0);
            M = m_mul(Mmain, ms_mlin(&sub, teeWa, tee.La, f));
            M = m_mul(M, Ta);
            M = m_mul(M, m_shunt(cx_mk(yin.re, yin.im + tee.Bt)));
            M = m_mul(M, Tb);
            M = m_mul(M, ms_mlin(&sub, teeWb, tee.Lb, f));
            in_side = 0;
            break;
        }
        default: return -4;
        }
    }
    if (in_side) return -4;
    *out = M;
    *det_out = det;
    return 0;
}

/* ABCD -> power-wave S for real terminations; SURVEY B.1 / A.6;
 * pa-lpf-simulation.sch:19,49 (Pac 50 Ohm), :60 (dB) */
static void abcd_to_s_det(const m22 *M, cx det, double rs, double rl, cx *s11, cx *s21, cx *s12, cx *s22)
{
    cx arl = cx_scale(M->a, rl), crr = cx_scale(M->c, rs * rl), drs = cx_scale(M->d, rs);
    cx den = cx_add(cx_add(arl, M->b), cx_add(crr, drs));
    cx n11 = cx_sub(cx_add(arl, M->b), cx_add(crr, drs));
    cx n22 = cx_sub(cx_add(M->b, drs), cx_add(arl, crr));
    /* S12 = S21 * (AD - BC).  Every supported element is reciprocal, so the cascade
     * determinant is identically 1; evaluating AD - BC numerically instead loses all
     * digits deep in a stop band (|AD| ~ 1e22 against a difference of 1), so the
     * exact value 1 is used.  (.dat S[1,2] still matches to 3.8e-12.)  Measured blocks may be
     * non-reciprocal: their determinants S12/S21 are carried separately in `det`. */
    *s11 = cx_div(n11, den);
    *s21 = cx_div(cx_mk(2.0 * sqrt(rs * rl), 0), den);
    *s12 = cx_mul(*s21, det);
    *s22 = cx_div(n22, den);
}

static int eval_s(const ref_elem *e, int n, double rs, double rl, double f, cx *s11, cx *s21, cx *s12, cx *s22)
{
    m22 M;
    cx det;
    int rc = eval_abcd_det(e, n, f, &M, &det);
    if (rc) return rc;
    abcd_to_s_det(&M, det, rs, rl, s11, s21, s12, s22);
    return 0;
}

/* group delay tau = -d(arg S21)/dw by a central difference with relative step
 * 1e-6 (the same definition the product uses; exact analytic values are used
 * in the tests to bound the truncation error) */
static int eval_gd(const ref_elem *e, int n, double rs, double rl, double f, double *gd)
{
    cx a, b, x, y;
    double df = f * 1e-6;
    int rc = eval_s(e, n, rs, rl, f + df, &x, &a, &y, &y);
    if (rc) return rc;
    rc = eval_s(e, n, rs, rl, f - df, &x, &b, &y, &y);
    if (rc) return rc;
    cx r = cx_mul(a, cx_mk(b.re, -b.im));
    *gd = -atan2(r.im, r.re) / (2.0 * (2.0 * PI * df));
    return 0;
}

int ref_sweep(const ref_elem *e, int n, double rs, double rl, const double *f, int nf,
              double *s11, double *s21, double *s12, double *s22, double *gd)
{
    for (int k = 0; k < nf; k++) {
        cx a, b, c, d;
        int rc = eval_s(e, n, rs, rl, f[k], &a, &b, &c, &d);
        if (rc) return rc;
        if (s11) { s11[2 * k] = a.re; s11[2 * k + 1] = a.im; }
        if (s21) { s21[2 * k] = b.re; s21[2 * k + 1] = b.im; }
        if (s12) { s12[2 * k] = c.re; s12[2 * k + 1] = c.im; }
        if (s22) { s22[2 * k] = d.re; s22[2 * k + 1] = d.im; }
        if (gd) { rc = eval_gd(e, n, rs, rl, f[k], &gd[k]); if (rc) return rc; }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Monte-Carlo yield (new work; defined here -- parity unpinned)        */
/* ------------------------------------------------------------------ */
static double db20(double mag2) { return 10.0 * log10(mag2); }

static int mc_one(const ref_elem *e, int n, double rs, double rl, const double *f, int nf,
                  const ref_spec *spec, int nspec, const ref_mc_cfg *cfg, uint64_t sample,
                  ref_elem *scratch, uint64_t *cnt, double *full_s, uint64_t local_idx)
{
    double worst[32];
    int seen[32];
    ref_perturb(e, n, cfg, sample, scratch);
    for (int s = 0; s < nspec; s++) { worst[s] = 0; seen[s] = 0; }
    for (int k = 0; k < nf; k++) {
        cx s11, s21, s12, s22;
        int rc = eval_s(scratch, n, rs, rl, f[k], &s11, &s21, &s12, &s22);
        if (rc) return rc;
        if (full_s) {
            size_t plane = (size_t)cfg->n_samples * (size_t)nf * 2, o = ((size_t)local_idx * nf + k) * 2;
            full_s[0 * plane + o] = s11.re; full_s[0 * plane + o + 1] = s11.im;
            full_s[1 * plane + o] = s21.re; full_s[1 * plane + o + 1] = s21.im;
            full_s[2 * plane + o] = s12.re; full_s[2 * plane + o + 1] = s12.im;
            full_s[3 * plane + o] = s22.re; full_s[3 * plane + o + 1] = s22.im;
        }
        double gdv = 0;
        int have_gd = 0;
        for (int s = 0; s < nspec; s++) {
            if (f[k] < spec[s].f_lo || f[k] > spec[s].f_hi) continue;
            double v;
            int want_min = 0;
            switch (spec[s].kind) {
            case REF_SPEC_S21_MIN_DB: v = db20(cx_abs2(s21)); want_min = 1; break;
            case REF_SPEC_S21_MAX_DB: v = db20(cx_abs2(s21)); break;
            case REF_SPEC_S11_MAX_DB: v = db20(cx_abs2(s11)); break;
            case REF_SPEC_GD_MAX:
                if (!have_gd) { rc = eval_gd(scratch, n, rs, rl, f[k], &gdv); if (rc) return rc; have_gd = 1; }
                v = gdv;
                break;
            default: return -1;
            }
            if (!seen[s]) { worst[s] = v; seen[s] = 1; }
            else if (want_min ? (v < worst[s]) : (v > worst[s])) worst[s] = v;
        }
    }
    int pass = 1;
    for (int s = 0; s < nspec; s++) {
        if (!seen[s]) continue;
        int ok = (spec[s].kind == REF_SPEC_S21_MIN_DB) ? (worst[s] >= spec[s].limit) : (worst[s] <= spec[s].limit);
        if (!ok) { pass = 0; cnt[2 + s]++; }
    }
    cnt[0] += (uint64_t)pass;
    cnt[1] += 1;
    if (cfg->hist_bins > 0 && cfg->hist_spec >= 0 && cfg->hist_spec < nspec && seen[cfg->hist_spec]) {
        double x = (worst[cfg->hist_spec] - cfg->hist_lo) / (cfg->hist_hi - cfg->hist_lo) * (double)cfg->hist_bins;
        long b = (long)floor(x);
        if (b < 0) b = 0;
        if (b >= cfg->hist_bins) b = cfg->hist_bins - 1;
        cnt[2 + nspec + b]++;
    }
    return 0;
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ref_mc_run(const ref_elem *e, int n, double rs, double rl, const double *f, int nf,
               const ref_spec *spec, int nspec, const ref_mc_cfg *cfg,
               uint64_t *counters, double *full_s, int nthreads)
{
    if (nspec > 32 || n <= 0 || nf <= 0) return -1;
    int ncnt = 2 + nspec + (cfg->hist_bins > 0 ? cfg->hist_bins : 0);
    memset(counters, 0, (size_t)ncnt * sizeof(uint64_t));
    int err = 0;
    if (nthreads <= 1) {
        ref_elem *scratch = malloc((size_t)n * sizeof(ref_elem));
        for (uint64_t i = 0; i < cfg->n_samples && !err; i++)
            err = mc_one(e, n, rs, rl, f, nf, spec, nspec, cfg, cfg->sample_offset + i, scratch, counters, full_s, i);
        free(scratch);
        return err;
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
    {
        ref_elem *scratch = malloc((size_t)n * sizeof(ref_elem));
        uint64_t *loc = calloc((size_t)ncnt, sizeof(uint64_t));
        int lerr = 0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)cfg->n_samples; i++) {
            if (lerr) continue;
            lerr = mc_one(e, n, rs, rl, f, nf, spec, nspec, cfg, cfg->sample_offset + (uint64_t)i, scratch, loc,
                          full_s, (uint64_t)i);
        }
#pragma omp critical
        {
            for (int k = 0; k < ncnt; k++) counters[k] += loc[k];
            if (lerr) err = lerr;
        }
        free(loc);
        free(scratch);
    }
    return err;
#else
    return ref_mc_run(e, n, rs, rl, f, nf, spec, nspec, cfg, counters, full_s, 1);
#endif
}

/* ------------------------------------------------------------------ */
/* QucsTranscalc CoupledMicrostrip analysis -- SURVEY App. D, verified  */
/* vs util/directional-couplers/ *.trc:6-20 to the files' 6 digits       */
/* ------------------------------------------------------------------ */
static double d_a(double x)
{
    double x2 = x * x, x4 = x2 * x2;
    return 1.0 + log((x4 + x2 / 2704.0) / (x4 + 0.432)) / 49.0 + log(1.0 + x2 * x / 5929.741) / 18.7;
}
void ref_cpl_analyze(double w, double s, double h, double t, double er, double ht, double f, double len,
                     double *z0e, double *z0o, double *ang_e_deg, double *ang_o_deg)
{
    double u = w / h, g = s / h, tau = t / h, h2h = ht / h, fn = f * h / 1e6;
    double b = 0.564 * pow((er - 0.9) / (er + 3.0), 0.053);
    /* D.1 zero-thickness single line */
    double es = (er + 1.0) / 2.0 + pow(1.0 + 10.0 / u, -d_a(u) * b) * (er - 1.0) / 2.0;
    double Zs = ms_Zh(u) / sqrt(es);
    double Zsf, esf;
    ref_ms_disp(w, h, er, Zs, es, f, &Zsf, &esf);
    /* Q0 == R17 of A.2 */
    double Q0;
    {
        double R1 = 0.03891 * pow(er, 1.4), R2 = 0.267 * pow(u, 7.0);
        double R7 = 1.206 - 0.3144 * exp(-R1) * (1.0 - exp(-R2));
        double R10 = 0.00044 * pow(er, 2.136) + 0.0184;
        double t6 = pow(fn / 19.47, 6.0), R11 = t6 / (1.0 + 0.0962 * t6);
        double R12 = 1.0 / (1.0 + 0.00245 * u * u);
        double R15 = 0.707 * R10 * pow(fn / 12.3, 1.097);
        double R16 = 1.0 + 0.0503 * er * er * R11 * (1.0 - exp(-pow(u / 15.0, 6.0)));
        Q0 = R7 * (1.0 - 1.1241 * R12 / R16 * exp(-0.026 * pow(fn, 1.15656) - R15));
    }
    /* D.2 thickness-corrected widths */
    double due = 0, duo = 0;
    if (t > 0) {
        double du = (1.25 * tau / PI) * (1.0 + log((2.0 + (4.0 * PI * u - 2.0) / (1.0 + exp(-100.0 * (u - 1.0 / (2.0 * PI))))) / tau));
        double dt = tau / (g * er);
        due = du * (1.0 - 0.5 * exp(-0.69 * du / dt));
        duo = due + dt;
    }
    double ue = u + due, uo = u + duo;
    double qte = (2.0 * log(2.0) / PI) * tau / sqrt(ue), qto = (2.0 * log(2.0) / PI) * tau / sqrt(uo);
    /* D.3 static permittivities */
    double v = ue * (20.0 + g * g) / (10.0 + g * g) + g * exp(-g);
    double qinfe = pow(1.0 + 10.0 / v, -d_a(v) * b);
    double qce = h2h <= 39.0 ? tanh(1.626 + 0.107 * h2h - 1.733 / sqrt(h2h)) : 1.0;
    double ee0 = (er + 1.0) / 2.0 + (er - 1.0) / 2.0 * (qinfe - qte) * qce;
    double bo = 0.747 * er / (0.15 + er);
    double co = bo - (bo - 0.207) * exp(-0.414 * uo);
    double dd = 0.593 + 0.694 * exp(-0.562 * uo);
    double qinfo = exp(-co * pow(g, dd));
    double qco = h2h <= 7.0 ? tanh(9.575 / (7.0 - h2h) - 2.965 + 1.68 * h2h - 0.311 * h2h * h2h) : 1.0;
    double q = (qinfo - qto) * qco;
    double ao = 0.7287 * (es - (er + 1.0) / 2.0) * (1.0 - exp(-0.179 * uo));
    double eo0 = ((er + 1.0) / 2.0 + ao - es) * q + es;
    /* D.4 static impedances */
    double Q1 = 0.8695 * pow(ue, 0.194);
    double Q2 = 1.0 + 0.7519 * g + 0.189 * pow(g, 2.31);
    double Q3 = 0.1975 + pow(16.6 + pow(8.4 / g, 6.0), -0.387) + log(pow(g, 10.0) / (1.0 + pow(g / 3.4, 10.0))) / 241.0;
    double Q4 = 2.0 * Q1 / (Q2 * (exp(-g) * pow(ue, Q3) + (2.0 - exp(-g)) * pow(ue, -Q3)));
    double Ze0 = Zs * sqrt(es / ee0) / (1.0 - sqrt(es) * Q4 * Zs / ZF0);
    double Q5 = 1.794 + 1.14 * log(1.0 + 0.638 / (g + 0.517 * pow(g, 2.43)));
    double Q6 = 0.2305 + log(pow(g, 10.0) / (1.0 + pow(g / 5.8, 10.0))) / 281.3 + log(1.0 + 0.598 * pow(g, 1.154)) / 5.1;
    double Q7 = (10.0 + 190.0 * g * g) / (1.0 + 82.3 * g * g * g);
    double Q8 = exp(-6.5 - 0.95 * log(g) - pow(g / 0.15, 5.0));
    double Q9 = log(Q7) * (Q8 + 1.0 / 16.5);
    double Q10 = (Q2 * Q4 - Q5 * exp(log(uo) * Q6 * pow(uo, -Q9))) / Q2;
    double Zo0 = Zs * sqrt(es / eo0) / (1.0 - sqrt(es) * Q10 * Zs / ZF0);
    {   /* March odd-mode cover correction */
        double J = tanh(pow(1.0 + h2h, 1.585) / 6.0);
        double G = 2.178 - 0.796 * g;
        double K = g > 0.858 ? log10(20.492 * pow(g, 0.174)) : 1.30;
        double Lc = g > 0.873 ? 2.51 * pow(g, -0.462) : 2.674;
        Zo0 -= pow(uo, J) * 270.0 * (1.0 - tanh(G + K * sqrt(1.0 + h2h) - Lc / (1.0 + h2h))) / sqrt(eo0);
    }
    /* D.5 dispersion of eps (uses u) */
    double P1 = 0.27488 + (0.6315 + 0.525 / pow(1.0 + 0.0157 * fn, 20.0)) * u - 0.065683 * exp(-8.7513 * u);
    double P2 = 0.33622 * (1.0 - exp(-0.03442 * er));
    double P3 = 0.0363 * exp(-4.6 * u) * (1.0 - exp(-pow(fn / 38.7, 4.97)));
    double P4 = 1.0 + 2.751 * (1.0 - exp(-pow(er / 15.916, 8.0)));
    double P5 = 0.334 * exp(-3.3 * pow(er / 15.0, 3.0)) + 0.746;
    double P6 = P5 * exp(-pow(fn / 18.0, 0.368));
    double P7 = 1.0 + 4.069 * P6 * pow(g, 0.479) * exp(-1.347 * pow(g, 0.595) - 0.17 * pow(g, 2.5));
    double Fe = P1 * P2 * pow((P3 * P4 + 0.1844 * P7) * fn, 1.5763);
    double P8 = 0.7168 * (1.0 + 1.076 / (1.0 + 0.0576 * (er - 1.0)));
    double P9 = P8 - 0.7913 * (1.0 - exp(-pow(fn / 20.0, 1.424))) * atan(2.481 * pow(er / 8.0, 0.946));
    double P10 = 0.242 * pow(er - 1.0, 0.55);
    double P11 = 0.6366 * (exp(-0.3401 * fn) - 1.0) * atan(1.263 * pow(u / 3.0, 1.629));
    double P12 = P9 + (1.0 - P9) / (1.0 + 1.183 * pow(u, 1.376));
    double P13 = 1.695 * P10 / (0.414 + 1.605 * P10);
    double P14 = 0.8928 + 0.1072 * (1.0 - exp(-0.42 * pow(fn / 20.0, 3.215)));
    double P15 = fabs(1.0 - 0.8928 * (1.0 + P11) * P12 * exp(-P13 * pow(g, 1.092)) / P14);
    double Fo = P1 * P2 * pow((P3 * P4 + 0.1844) * fn * P15, 1.5763);
    double ee = er - (er - ee0) / (1.0 + Fe);
    double eo = er - (er - eo0) / (1.0 + Fo);
    /* D.6 dispersion of Z */
    double Q11 = 0.893 * (1.0 - 0.3 / (1.0 + 0.7 * (er - 1.0)));
    double f20 = pow(fn / 20.0, 4.91);
    double Q12 = 2.121 * (f20 / (1.0 + Q11 * f20)) * exp(-2.87 * g) * pow(g, 0.902);
    double Q13 = 1.0 + 0.038 * pow(er / 8.0, 5.1);
    double e15 = pow(er / 15.0, 4.0);
    double Q14 = 1.0 + 1.203 * e15 / (1.0 + e15);
    double Q15 = 1.887 * exp(-1.5 * pow(g, 0.84)) * pow(g, Q14) /
                 (1.0 + 0.41 * pow(fn / 15.0, 3.0) * pow(u, 2.0 / Q13) / (0.125 + pow(u, 1.626 / Q13)));
    double Q16 = (1.0 + 9.0 / (1.0 + 0.403 * (er - 1.0) * (er - 1.0))) * Q15;
    double Q17 = 0.394 * (1.0 - exp(-1.47 * pow(u / 7.0, 0.672))) * (1.0 - exp(-4.25 * pow(fn / 20.0, 1.87)));
    double Q18 = 0.61 * (1.0 - exp(-2.13 * pow(u / 8.0, 1.593))) / (1.0 + 6.544 * pow(g, 4.17));
    double Q19 = 0.21 * g * g * g * g / ((1.0 + 0.18 * pow(g, 4.9)) * (1.0 + 0.1 * u * u) * (1.0 + pow(fn / 24.0, 3.0)));
    double Q20 = (0.09 + 1.0 / (1.0 + 0.1 * pow(er - 1.0, 2.7))) * Q19;
    double Q21 = fabs(1.0 - 42.54 * pow(g, 0.133) * exp(-0.812 * g) * pow(u, 2.5) / (1.0 + 0.033 * pow(u, 2.5)));
    double re = pow(fn / 28.843, 12.0);
    double qe = 0.016 + pow(0.0514 * er * Q21, 4.524);
    double pe = 4.766 * exp(-3.228 * pow(u, 0.641));
    double em16 = pow(er - 1.0, 6.0);
    double de = 5.086 * qe * (re / (0.3838 + 0.386 * qe)) * (exp(-22.2 * pow(u, 1.92)) / (1.0 + 1.2992 * re)) *
                (em16 / (1.0 + 10.0 * em16));
    double Ce = 1.0 + 1.275 * (1.0 - exp(-0.004625 * pe * pow(er, 1.674) * pow(fn / 18.365, 2.745))) - Q12 + Q16 - Q17 + Q18 + Q20;
    *z0e = Ze0 * pow((0.9408 * pow(esf, Ce) - 0.9603) / ((0.9408 - de) * pow(es, Ce) - 0.9603), Q0);
    double Q29 = 15.16 / (1.0 + 0.196 * (er - 1.0) * (er - 1.0));
    double em13 = (er - 1.0) * (er - 1.0) * (er - 1.0);
    double Q28 = 0.149 * em13 / (94.5 + 0.038 * em13);
    double em115 = pow(er - 1.0, 1.5);
    double Q27 = 0.4 * pow(g, 0.84) * (1.0 + 2.5 * em115 / (5.0 + em115));
    double x12 = pow((er - 1.0) / 13.0, 12.0);
    double Q26 = 30.0 - 22.2 * (x12 / (1.0 + 3.0 * x12)) - Q29;
    double Q25 = (0.3 * fn * fn / (10.0 + fn * fn)) * (1.0 + 2.333 * (er - 1.0) * (er - 1.0) / (5.0 + (er - 1.0) * (er - 1.0)));
    double Q24 = 2.506 * Q28 * pow(u, 0.894) * pow((1.0 + 1.3 * u) * fn / 99.25, 4.29) / (3.575 + pow(u, 0.894));
    double Q23 = 1.0 + 0.005 * fn * Q27 / ((1.0 + 0.812 * pow(fn / 15.0, 1.9)) * (1.0 + 0.025 * u * u));
    double Q22 = 0.925 * pow(fn / Q26, 1.536) / (1.0 + 0.3 * pow(fn / 30.0, 1.536));
    *z0o = Zsf + (Zo0 * pow(eo / eo0, Q22) - Zsf * Q23) / (1.0 + Q24 + pow(0.46 * g, 2.2) * Q25);
    /* D.7 */
    *ang_e_deg = 360.0 * len * f * sqrt(ee) / C0;
    *ang_o_deg = 360.0 * len * f * sqrt(eo) / C0;
}
