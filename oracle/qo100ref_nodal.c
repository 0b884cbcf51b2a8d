/*
 * qo100ref_nodal.c -- CPU ORACLE, N-port nodal analysis.  TEST INFRASTRUCTURE ONLY
 * (see qo100ref.h for the rules).
 *
 * What it restates: qucsator's S-parameter analysis of a general linear network, as exercised by
 * util/pa-bias-simulation/pa-bias-simulation.sch:19-72 (R, C, GND, Pac ports, an ideal VCVS buffer :40,59 and a
 * measured inductor pulled in with SPfile :39) whose result is the reference dataset
 * util/pa-bias-simulation/pa-bias-simulation.dat:1-85035 (5 ports, 5000 points 1 MHz - 10 GHz) -- PINNED:
 * this file reproduces that dataset to <= 2e-10 relative on every through entry (tests/test_nodal.py).
 *
 * Method (modified nodal analysis): unknowns = node voltages 1..n (0 = ground) plus one branch current per
 * VCVS.  A = Y_nodal + port terminations 1/Z_k.  With port j driven by a 1 V source behind Z_j (Norton: 1/Z_j
 * into its node) and all other ports terminated:   S[k][j] = 2 sqrt(Z_j/Z_k) V_k - delta_kj.
 * One LU factorisation (partial pivoting) per frequency, one substitution per port.
 */
#define _GNU_SOURCE
#include "qo100ref.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cx;
static inline cx cmk(double a, double b) { cx z = { a, b }; return z; }
static inline cx cadd(cx a, cx b) { return cmk(a.re + b.re, a.im + b.im); }
static inline cx csub(cx a, cx b) { return cmk(a.re - b.re, a.im - b.im); }
static inline cx cmul(cx a, cx b) { return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline cx cdiv(cx a, cx b)
{
    double d = b.re * b.re + b.im * b.im;
    return cmk((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
static inline double cabs2(cx a) { return a.re * a.re + a.im * a.im; }

#define NMAX 32
static const double PI_ = 3.14159265358979323846;

int ref_sblock_s(int idx, double f, int polar, double s[8]);   /* qo100ref.c: interpolated S11,S21,S12,S22 (re,im) */

static void stamp_y(cx A[NMAX][NMAX], int a, int b, cx y)
{
    if (a) A[a - 1][a - 1] = cadd(A[a - 1][a - 1], y);
    if (b) A[b - 1][b - 1] = cadd(A[b - 1][b - 1], y);
    if (a && b) { A[a - 1][b - 1] = csub(A[a - 1][b - 1], y); A[b - 1][a - 1] = csub(A[b - 1][a - 1], y); }
}

/* S at one frequency; s_out[np*np] row-major S[k][j].  Returns 0, -4 unsupported, -9 singular */
static int nodal_point(const ref_branch *br, int nb, int n_nodes, const int *port_node, const double *port_z0, int np,
                       double f, cx *s_out)
{
    cx A[NMAX][NMAX];
    const double w = 2.0 * PI_ * f;
    int n = n_nodes;
    for (int i = 0; i < nb; i++) if (br[i].kind == REF_NB_VCVS) n++;
    if (n > NMAX) return -4;
    memset(A, 0, sizeof A);
    int extra = n_nodes;
    for (int i = 0; i < nb; i++) {
        const double *p = br[i].p;
        const int *nd = br[i].node;
        switch (br[i].kind) {
        case REF_NB_R: stamp_y(A, nd[0], nd[1], cmk(1.0 / p[0], 0)); break;
        case REF_NB_L: {            /* (R + jwL) || 1/(jwCp) */
            cx z = cdiv(cmk(p[1], w * p[0]), cmk(1.0 - w * w * p[0] * p[2], w * p[1] * p[2]));
            stamp_y(A, nd[0], nd[1], cdiv(cmk(1, 0), z));
            break;
        }
        case REF_NB_C: {            /* R + jwLs + 1/(jwC) */
            cx z = cmk(p[1], w * p[2] - 1.0 / (w * p[0]));
            stamp_y(A, nd[0], nd[1], cdiv(cmk(1, 0), z));
            break;
        }
        case REF_NB_VCVS: {         /* nodes in+, out+, out-, in-; V(o+) - V(o-) = G e^{-jwT} (V(i+) - V(i-)) */
            const int ip = nd[0], op = nd[1], om = nd[2], im = nd[3], k = extra++;
            const cx g = cmk(p[0] * cos(w * p[1]), -p[0] * sin(w * p[1]));
            if (op) { A[op - 1][k] = cadd(A[op - 1][k], cmk(1, 0)); A[k][op - 1] = cadd(A[k][op - 1], cmk(1, 0)); }
            if (om) { A[om - 1][k] = csub(A[om - 1][k], cmk(1, 0)); A[k][om - 1] = csub(A[k][om - 1], cmk(1, 0)); }
            if (ip) A[k][ip - 1] = csub(A[k][ip - 1], g);
            if (im) A[k][im - 1] = cadd(A[k][im - 1], g);
            break;
        }
        case REF_NB_SBLOCK: {       /* terminals t1, t2, reference r: Y = (I - S)(I + S)^-1 / z0 */
            double sv[8];
            if (ref_sblock_s((int)p[0], f, p[1] != 0.0, sv)) return -4;
            const double z0 = p[2];
            const cx s11 = cmk(sv[0], sv[1]), s21 = cmk(sv[2], sv[3]), s12 = cmk(sv[4], sv[5]), s22 = cmk(sv[6], sv[7]);
            /* (I + S)^-1 = 1/det [1+s22, -s12; -s21, 1+s11] */
            const cx a = cadd(cmk(1, 0), s11), d = cadd(cmk(1, 0), s22);
            const cx det = csub(cmul(a, d), cmul(s12, s21));
            const cx m11 = csub(cmk(1, 0), s11), m22 = csub(cmk(1, 0), s22);
            /* (I - S) (I + S)^-1 */
            cx y[2][2];
            y[0][0] = cdiv(cadd(cmul(m11, d), cmul(s12, s21)), det);
            y[0][1] = cdiv(csub(cmul(cmk(-s12.re, -s12.im), a), cmul(m11, s12)), det);
            y[1][0] = cdiv(csub(cmul(cmk(-s21.re, -s21.im), d), cmul(m22, s21)), det);
            y[1][1] = cdiv(cadd(cmul(s21, s12), cmul(m22, a)), det);
            const int t[3] = { nd[0], nd[1], nd[2] };
            /* indefinite admittance: third terminal row/column = minus the sums */
            cx Y3[3][3];
            for (int r = 0; r < 2; r++) for (int c = 0; c < 2; c++) Y3[r][c] = cmk(y[r][c].re / z0, y[r][c].im / z0);
            for (int r = 0; r < 2; r++) Y3[r][2] = cmk(-(Y3[r][0].re + Y3[r][1].re), -(Y3[r][0].im + Y3[r][1].im));
            for (int c = 0; c < 3; c++) Y3[2][c] = cmk(-(Y3[0][c].re + Y3[1][c].re), -(Y3[0][c].im + Y3[1][c].im));
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
                if (t[r] && t[c]) A[t[r] - 1][t[c] - 1] = cadd(A[t[r] - 1][t[c] - 1], Y3[r][c]);
            break;
        }
        default: return -4;
        }
    }
    for (int k = 0; k < np; k++) stamp_y(A, port_node[k], 0, cmk(1.0 / port_z0[k], 0));
    /* LU with partial pivoting, in place; perm[i] = original row now at i */
    int perm[NMAX];
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int c = 0; c < n; c++) {
        int piv = c;
        double best = cabs2(A[c][c]);
        for (int r = c + 1; r < n; r++) { double v = cabs2(A[r][c]); if (v > best) { best = v; piv = r; } }
        if (!(best > 0)) return -9;
        if (piv != c) {
            for (int k = 0; k < n; k++) { cx t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
            int t = perm[c]; perm[c] = perm[piv]; perm[piv] = t;
        }
        const cx inv = cdiv(cmk(1, 0), A[c][c]);
        for (int r = c + 1; r < n; r++) {
            const cx l = cmul(A[r][c], inv);
            A[r][c] = l;
            for (int k = c + 1; k < n; k++) A[r][k] = csub(A[r][k], cmul(l, A[c][k]));
        }
    }
    for (int j = 0; j < np; j++) {
        cx x[NMAX];
        for (int i = 0; i < n; i++) x[i] = cmk(perm[i] == port_node[j] - 1 ? 1.0 / port_z0[j] : 0.0, 0);
        for (int i = 0; i < n; i++) for (int k = 0; k < i; k++) x[i] = csub(x[i], cmul(A[i][k], x[k]));
        for (int i = n - 1; i >= 0; i--) {
            for (int k = i + 1; k < n; k++) x[i] = csub(x[i], cmul(A[i][k], x[k]));
            x[i] = cdiv(x[i], A[i][i]);
        }
        for (int k = 0; k < np; k++) {
            const double sc = 2.0 * sqrt(port_z0[j] / port_z0[k]);
            cx v = x[port_node[k] - 1];
            s_out[k * np + j] = cmk(sc * v.re - (k == j ? 1.0 : 0.0), sc * v.im);
        }
    }
    return 0;
}

int ref_nodal_sweep(const ref_branch *br, int nb, int n_nodes, const int *port_node, const double *port_z0, int np,
                    const double *f, int nf, double *s_out)
{
    for (int k = 0; k < np; k++) if (port_node[k] < 1 || port_node[k] > n_nodes) return -1;
    for (int k = 0; k < nf; k++) {
        int rc = nodal_point(br, nb, n_nodes, port_node, port_z0, np, f[k], (cx *)(s_out + (size_t)k * np * np * 2));
        if (rc) return rc;
    }
    return 0;
}

/* Monte Carlo over branch parameters (tolerance entries address (branch, param)); specs on |S[row][col]| in dB.
 * counters as in ref_mc_run; full_s (nullable): [n_samples][nf][np][np] (re,im). */
int ref_nodal_mc_run(const ref_branch *br, int nb, int n_nodes, const int *port_node, const double *port_z0, int np,
                     const double *f, int nf, const ref_nspec *spec, int nspec, const ref_mc_cfg *cfg,
                     uint64_t *counters, double *full_s, int nthreads)
{
    if (nspec > 32 || nb <= 0 || nf <= 0 || np < 1 || np > 8) return -1;
    const int ncnt = 2 + nspec + (cfg->hist_bins > 0 ? cfg->hist_bins : 0);
    memset(counters, 0, (size_t)ncnt * sizeof(uint64_t));
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        ref_branch *loc = malloc((size_t)nb * sizeof(ref_branch));
        uint64_t *cnt = calloc((size_t)ncnt, sizeof(uint64_t));
        cx *sp = malloc((size_t)np * np * sizeof(cx));
        int lerr = 0;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)cfg->n_samples; i++) {
            if (lerr) continue;
            const uint64_t sample = cfg->sample_offset + (uint64_t)i;
            memcpy(loc, br, (size_t)nb * sizeof(ref_branch));
            for (int t = 0; t < cfg->n_tol; t++) {
                const ref_tol *tl = &cfg->tol[t];
                const double x = ref_variate(cfg->seed, sample, (uint32_t)tl->var, cfg->dist);
                double *v = &loc[tl->elem].p[tl->param];
                *v = tl->mode == REF_TOL_ABS ? fma(tl->tol, x, *v) : *v * fma(tl->tol, x, 1.0);
            }
            double worst[32];
            int seen[32];
            for (int s = 0; s < nspec; s++) { worst[s] = 0; seen[s] = 0; }
            for (int k = 0; k < nf && !lerr; k++) {
                lerr = nodal_point(loc, nb, n_nodes, port_node, port_z0, np, f[k], sp);
                if (lerr) break;
                if (full_s) memcpy(full_s + (((size_t)i * nf + k) * np * np) * 2, sp, (size_t)np * np * sizeof(cx));
                for (int s = 0; s < nspec; s++) {
                    if (f[k] < spec[s].f_lo || f[k] > spec[s].f_hi) continue;
                    const double v = 10.0 * log10(cabs2(sp[spec[s].row * np + spec[s].col]));
                    const int want_min = spec[s].kind == REF_SPEC_S21_MIN_DB;
                    if (!seen[s]) { worst[s] = v; seen[s] = 1; }
                    else if (want_min ? (v < worst[s]) : (v > worst[s])) worst[s] = v;
                }
            }
            if (lerr) continue;
            int pass = 1;
            for (int s = 0; s < nspec; s++) {
                if (!seen[s]) continue;
                const int ok = spec[s].kind == REF_SPEC_S21_MIN_DB ? worst[s] >= spec[s].limit : worst[s] <= spec[s].limit;
                if (!ok) { pass = 0; cnt[2 + s]++; }
            }
            cnt[0] += (uint64_t)pass;
            cnt[1] += 1;
            if (cfg->hist_bins > 0 && cfg->hist_spec >= 0 && cfg->hist_spec < nspec && seen[cfg->hist_spec]) {
                const double xb = (worst[cfg->hist_spec] - cfg->hist_lo) / (cfg->hist_hi - cfg->hist_lo) * (double)cfg->hist_bins;
                long b = (long)floor(xb);
                if (b < 0) b = 0;
                if (b >= cfg->hist_bins) b = cfg->hist_bins - 1;
                cnt[2 + nspec + b]++;
            }
        }
#pragma omp critical
        {
            for (int k = 0; k < ncnt; k++) counters[k] += cnt[k];
            if (lerr) err = lerr;
        }
        free(loc); free(cnt); free(sp);
    }
    return err;
}
