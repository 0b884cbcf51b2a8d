/*
 * qo_net.c -- element lists, ladder synthesis, grids, error strings and the host
 * twins of the perturbation stream.  Host C only; no CUDA in this file.
 */
#define _GNU_SOURCE
#include "qo_internal.h"
#include "qo_stream.h"
#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[512];

void qo_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
void qo_clear_error(void) { g_err[0] = 0; }
const char *qo_last_error(void) { return g_err; }
const char *qo_version(void) { return "qo100net 0.1 (sm_100a)"; }

const char *qo_strerror(int st)
{
    switch (st) {
    case QO_OK: return "ok";
    case QO_ERR_ARG: return "invalid argument";
    case QO_ERR_IO: return "i/o error";
    case QO_ERR_PARSE: return "parse error";
    case QO_ERR_UNSUPPORTED: return "unsupported element or topology";
    case QO_ERR_NOMEM: return "out of memory";
    case QO_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case QO_ERR_CUDA: return "CUDA error";
    case QO_ERR_NCCL: return "NCCL error";
    case QO_ERR_RANGE: return "value out of range";
    default: return "unknown status";
    }
}

char *qo_read_file(const char *path, size_t *len)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) { qo_set_error("cannot open %s", path); return NULL; }
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (n < 0) { fclose(fp); return NULL; }
    char *buf = (char *)malloc((size_t)n + 1);
    if (!buf) { fclose(fp); return NULL; }
    size_t got = fread(buf, 1, (size_t)n, fp);
    fclose(fp);
    buf[got] = 0;
    if (len) *len = got;
    return buf;
}

/* number + optional SI-prefixed unit: "4.700 pF", "1.300 uH", "100.00 Ω", "0.6 mm",
 * "34.79 um", "10 GHz", ".75e-3", "20 cm", "0 mil".  *unit (nullable) receives the
 * unit text after the prefix has been applied (e.g. "F", "H", "m", "Hz", "Ohm"). */
int qo_parse_value(const char *s, double *out, const char **unit)
{
    while (*s && isspace((unsigned char)*s)) s++;
    char *end;
    double v = strtod(s, &end);
    if (end == s) return QO_ERR_PARSE;
    while (*end && isspace((unsigned char)*end)) end++;
    const char *u = end;
    double scale = 1.0;
    /* whole-word units first (so that "mil", "m", "mm", "cm" are told apart) */
    if (!strncmp(u, "mil", 3)) { scale = 25.4e-6; u += 3; }
    else if (!strncmp(u, "cm", 2)) { scale = 1e-2; u += 1; }
    else if (!strncmp(u, "dBm", 3) || !strncmp(u, "dB", 2) || !strncmp(u, "Deg", 3) || !strncmp(u, "NA", 2)) { /* no prefix */ }
    else if ((unsigned char)u[0] == 0xC2 && (unsigned char)u[1] == 0xB5) { scale = 1e-6; u += 2; } /* µ */
    else if ((unsigned char)u[0] == 0xCE && (unsigned char)u[1] == 0xBC) { scale = 1e-6; u += 2; } /* μ */
    else if (*u && u[1] && !isspace((unsigned char)u[1])) {
        /* a prefix is only a prefix when a unit follows it */
        switch (*u) {
        case 'f': scale = 1e-15; u++; break;
        case 'p': scale = 1e-12; u++; break;
        case 'n': scale = 1e-9; u++; break;
        case 'u': scale = 1e-6; u++; break;
        case 'm': scale = 1e-3; u++; break;
        case 'k': scale = 1e3; u++; break;
        case 'M': scale = 1e6; u++; break;
        case 'G': scale = 1e9; u++; break;
        case 'T': scale = 1e12; u++; break;
        default: break;
        }
    }
    else if (*u && (!u[1] || isspace((unsigned char)u[1]))) {
        /* a bare engineering suffix, as Qucs accepts it ("26.5p", pa-bias-simulation.sch:42); a lone "m" stays
         * the unit metre (QucsTranscalc lengths) */
        switch (*u) {
        case 'f': scale = 1e-15; u++; break;
        case 'p': scale = 1e-12; u++; break;
        case 'n': scale = 1e-9; u++; break;
        case 'u': scale = 1e-6; u++; break;
        case 'k': scale = 1e3; u++; break;
        case 'M': scale = 1e6; u++; break;
        case 'G': scale = 1e9; u++; break;
        case 'T': scale = 1e12; u++; break;
        default: break;
        }
    }
    *out = v * scale;
    if (scale != 1.0 && scale != 25.4e-6) {
        /* decimal prefixes: re-read "33.00" + "e-9" so that the value is the correctly rounded
         * decimal (33e-9), not 33 * 1e-9 which is one ulp off */
        size_t nlen = (size_t)(end - s);
        while (nlen && isspace((unsigned char)s[nlen - 1])) nlen--;
        if (nlen < 48 && !memchr(s, 'e', nlen) && !memchr(s, 'E', nlen)) {
            char tmp[64];
            memcpy(tmp, s, nlen);
            snprintf(tmp + nlen, sizeof tmp - nlen, "e%d", (int)lround(log10(scale)));
            *out = strtod(tmp, NULL);
        }
    }
    if (unit) *unit = u;
    return QO_OK;
}

qo_net *qo_net_alloc(void)
{
    qo_net *n = (qo_net *)calloc(1, sizeof(qo_net));
    if (n) { n->rs = 50.0; n->rl = 50.0; }
    return n;
}

void qo_net_free(qo_net *net)
{
    if (!net) return;
    for (int i = 0; i < net->nblk; i++) qo_s2p_free(net->blk[i]);
    free(net);
}

/* QO_SBLOCK elements enter a network only through qo_net_from_sblock / qo_net_concat (they carry an index
 * into the net's own block list), never through a raw element list */
static int kind_ok(int k) { return (k >= QO_SER_R && k <= QO_MOPEN) || k == QO_CPL_MS; }

int qo_net_from_sblock(const qo_s2p *blk, int polar, double rs, double rl, qo_net **out)
{
    qo_clear_error();
    if (!blk || !out || !(rs > 0) || !(rl > 0)) { qo_set_error("bad arguments"); return QO_ERR_ARG; }
    qo_net *net = qo_net_alloc();
    if (!net) return QO_ERR_NOMEM;
    net->blk[0] = qo_s2p_clone(blk);
    if (!net->blk[0]) { free(net); return QO_ERR_NOMEM; }
    net->nblk = 1;
    net->n = 1; net->rs = rs; net->rl = rl;
    net->e[0].kind = QO_SBLOCK;
    net->e[0].p[0] = 0.0;
    net->e[0].p[1] = polar ? 1.0 : 0.0;
    snprintf(net->title, sizeof net->title, "S-parameter block, %d points, z0 %g", blk->n, blk->z0);
    *out = net;
    return QO_OK;
}

/* per-kind parameter ranges: a non-positive or non-finite L, C, Z0, f0 ... would only surface as inf / NaN on the device
 * (qo_nodal_add_branch applies the same rules to its branches) */
static const char *elem_param_error(const qo_elem *el)
{
    const double *p = el->p;
    for (int k = 0; k < QO_NPARAM; k++) if (!isfinite(p[k])) return "non-finite parameter";
    switch (el->kind) {
    case QO_SER_R: return p[0] >= 0 ? NULL : "R must be >= 0";
    case QO_SHUNT_R: return p[0] > 0 ? NULL : "R must be > 0";
    case QO_SER_L: case QO_SHUNT_L: return p[0] > 0 && p[1] >= 0 && p[2] >= 0 ? NULL : "L must be > 0, ESR and Cp >= 0";
    case QO_SER_C: case QO_SHUNT_C: return p[0] > 0 && p[1] >= 0 && p[2] >= 0 ? NULL : "C must be > 0, ESR and ESL >= 0";
    case QO_SER_LC_SER: case QO_SER_LC_PAR: case QO_SHUNT_LC_SER: case QO_SHUNT_LC_PAR: return p[0] > 0 && p[1] > 0 ? NULL : "L and C must be > 0";
    case QO_TLINE: return p[0] > 0 && p[2] > 0 ? NULL : "Z0 and f0 must be > 0";
    case QO_CPL_THRU: return p[0] > 0 && p[1] > 0 && p[4] > 0 && p[5] > 0 ? NULL : "Z0e, Z0o, f0 and Zt must be > 0";
    case QO_SUBST: return p[0] >= 1 && p[1] > 0 && p[2] >= 0 && p[3] >= 0 && p[4] >= 0 && p[5] >= 0 ? NULL : "SUBST needs er >= 1, h > 0, t, tand, rho, D >= 0";
    case QO_MLIN: return p[0] > 0 ? NULL : "W must be > 0";
    case QO_MCORN: case QO_MOPEN: return p[0] > 0 ? NULL : "W must be > 0";
    case QO_MTEE: return p[0] > 0 && p[1] > 0 && p[2] > 0 ? NULL : "Wa, Wb, W2 must be > 0";
    case QO_SBLOCK: return p[0] >= 0 && p[0] == floor(p[0]) ? NULL : "block index must be a non-negative integer";
    default: return NULL;
    }
}

int qo_net_from_elements(const qo_elem *e, int n, double rs, double rl, qo_net **out)
{
    qo_clear_error();
    if (!e || !out || n <= 0 || !(rs > 0) || !(rl > 0) || !isfinite(rs) || !isfinite(rl)) { qo_set_error("bad arguments"); return QO_ERR_ARG; }
    if (n > QO_MAX_ELEMS) { qo_set_error("too many elements (%d > %d)", n, QO_MAX_ELEMS); return QO_ERR_RANGE; }
    int side = 0, have_sub = 0, n_cplms = 0;
    for (int i = 0; i < n; i++) {
        if (!kind_ok(e[i].kind)) { qo_set_error("element %d: unknown kind %d", i, e[i].kind); return QO_ERR_UNSUPPORTED; }
        { const char *why = elem_param_error(&e[i]); if (why) { qo_set_error("element %d (kind %d): %s", i, e[i].kind, why); return QO_ERR_ARG; } }
        if (e[i].kind == QO_SUBST) have_sub = 1;
        if (e[i].kind >= QO_MLIN && !have_sub) { qo_set_error("element %d: microstrip element before any SUBST", i); return QO_ERR_ARG; }
        if (e[i].kind == QO_CPL_MS) {
            if (++n_cplms > 1) { qo_set_error("element %d: at most one physical coupled-line element per network", i); return QO_ERR_UNSUPPORTED; }
            if (!(e[i].p[0] > 0 && e[i].p[1] > 0 && e[i].p[2] > 0 && e[i].p[3] > 0 && e[i].p[4] > 0 && e[i].p[5] > 0)) { qo_set_error("element %d: QO_CPL_MS needs W, S, L, H_t, f0, Zt > 0", i); return QO_ERR_ARG; }
        }
        if (e[i].kind == QO_MTEE) { if (side) { qo_set_error("element %d: nested MTEE", i); return QO_ERR_UNSUPPORTED; } side = 1; }
        if (e[i].kind == QO_MOPEN) { if (!side) { qo_set_error("element %d: MOPEN outside a tee side arm", i); return QO_ERR_UNSUPPORTED; } side = 0; }
    }
    if (side) { qo_set_error("MTEE side arm not closed by MOPEN"); return QO_ERR_UNSUPPORTED; }
    qo_net *net = qo_net_alloc();
    if (!net) return QO_ERR_NOMEM;
    net->n = n; net->rs = rs; net->rl = rl;
    memcpy(net->e, e, (size_t)n * sizeof(qo_elem));
    *out = net;
    return QO_OK;
}

int qo_net_num_elements(const qo_net *net) { return net ? net->n : QO_ERR_ARG; }
int qo_net_get_elements(const qo_net *net, qo_elem *out, int cap)
{
    if (!net || !out || cap < net->n) return QO_ERR_ARG;
    memcpy(out, net->e, (size_t)net->n * sizeof(qo_elem));
    return net->n;
}
int qo_net_terminations(const qo_net *net, double *rs, double *rl)
{
    if (!net) return QO_ERR_ARG;
    if (rs) *rs = net->rs;
    if (rl) *rl = net->rl;
    return QO_OK;
}
const char *qo_net_title(const qo_net *net) { return net ? net->title : ""; }

int qo_net_concat(const qo_net *a, const qo_net *b, qo_net **out)
{
    qo_clear_error();
    if (!a || !b || !out) return QO_ERR_ARG;
    if (a->n + b->n > QO_MAX_ELEMS) { qo_set_error("too many elements"); return QO_ERR_RANGE; }
    if (a->nblk + b->nblk > QO_MAX_BLK) { qo_set_error("too many S-parameter blocks (%d > %d)", a->nblk + b->nblk, QO_MAX_BLK); return QO_ERR_RANGE; }
    qo_net *net = qo_net_alloc();
    if (!net) return QO_ERR_NOMEM;
    memcpy(net->e, a->e, (size_t)a->n * sizeof(qo_elem));
    memcpy(net->e + a->n, b->e, (size_t)b->n * sizeof(qo_elem));
    for (int i = 0; i < a->nblk + b->nblk; i++) {
        net->blk[i] = qo_s2p_clone(i < a->nblk ? a->blk[i] : b->blk[i - a->nblk]);
        if (!net->blk[i]) { net->nblk = i; qo_net_free(net); return QO_ERR_NOMEM; }
    }
    net->nblk = a->nblk + b->nblk;
    for (int i = 0; i < b->n; i++)           /* b's blocks follow a's in the merged list */
        if (net->e[a->n + i].kind == QO_SBLOCK) net->e[a->n + i].p[0] += (double)a->nblk;
    net->n = a->n + b->n; net->rs = a->rs; net->rl = b->rl;
    snprintf(net->title, sizeof net->title, "%.200s | %.200s", a->title, b->title);
    *out = net;
    return QO_OK;
}

/* ---- ladder synthesis (pcb/generic-filter/README.md:13) ------------------
 * Lowpass prototype g-values: Chebyshev by the closed form of Matthaei/Young/
 * Jones (equal-ripple, odd order => g_{n+1}=1), Butterworth 2 sin((2k-1)pi/2n). */
static int proto_g(int order, double ripple_db, int cheby, double *g)
{
    const double pi = 3.14159265358979323846;
    if (order < 1 || order > 31) return QO_ERR_RANGE;
    if (!cheby) {
        for (int k = 1; k <= order; k++) g[k - 1] = 2.0 * sin((2.0 * k - 1.0) * pi / (2.0 * order));
        return QO_OK;
    }
    if (!(ripple_db > 0)) return QO_ERR_ARG;
    double beta = log(1.0 / tanh(ripple_db * log(10.0) / 40.0));
    double gamma = sinh(beta / (2.0 * order));
    double a_prev = sin(pi / (2.0 * order)), b_prev = 0.0;
    g[0] = 2.0 * a_prev / gamma;
    b_prev = gamma * gamma + sin(pi / order) * sin(pi / order);
    for (int k = 2; k <= order; k++) {
        double a = sin((2.0 * k - 1.0) * pi / (2.0 * order));
        g[k - 1] = 4.0 * a_prev * a / (b_prev * g[k - 2]);
        double sk = sin(k * pi / order);
        a_prev = a;
        b_prev = gamma * gamma + sk * sk;
    }
    return QO_OK;
}

static int ladder(int order, const double *g, double fc, double z0, int series_first, const char *name, qo_net **out)
{
    const double pi = 3.14159265358979323846;
    if (!(fc > 0) || !(z0 > 0) || order > QO_MAX_ELEMS) return QO_ERR_ARG;
    qo_net *net = qo_net_alloc();
    if (!net) return QO_ERR_NOMEM;
    double wc = 2.0 * pi * fc;
    for (int k = 0; k < order; k++) {
        int series = series_first ? (k % 2 == 0) : (k % 2 == 1);
        if (series) { net->e[k].kind = QO_SER_L; net->e[k].p[0] = g[k] * z0 / wc; }
        else { net->e[k].kind = QO_SHUNT_C; net->e[k].p[0] = g[k] / (z0 * wc); }
    }
    net->n = order; net->rs = z0; net->rl = z0;
    snprintf(net->title, sizeof net->title, "%s order %d fc %.9g Hz z0 %.6g", name, order, fc, z0);
    *out = net;
    return QO_OK;
}

int qo_net_cheby_lpf(int order, double ripple_db, double fc, double z0, int series_first, qo_net **out)
{
    qo_clear_error();
    double g[32];
    if (!out) return QO_ERR_ARG;
    if (order % 2 == 0) { qo_set_error("even-order Chebyshev needs unequal terminations; not supported"); return QO_ERR_UNSUPPORTED; }
    int rc = proto_g(order, ripple_db, 1, g);
    if (rc) return rc;
    return ladder(order, g, fc, z0, series_first, "chebyshev-lpf", out);
}

int qo_net_butter_lpf(int order, double fc, double z0, int series_first, qo_net **out)
{
    qo_clear_error();
    double g[32];
    if (!out) return QO_ERR_ARG;
    int rc = proto_g(order, 0.0, 0, g);
    if (rc) return rc;
    return ladder(order, g, fc, z0, series_first, "butterworth-lpf", out);
}

int qo_net_add_parasitics(qo_net *net, double fc, double q_l, double srf_l_mult, double esr_c, double srf_c_mult)
{
    const double pi = 3.14159265358979323846;
    if (!net || !(fc > 0) || !(q_l > 0) || !(srf_l_mult > 0) || !(srf_c_mult > 0) || esr_c < 0) return QO_ERR_ARG;
    double wc = 2.0 * pi * fc, wl = 2.0 * pi * srf_l_mult * fc, wcap = 2.0 * pi * srf_c_mult * fc;
    for (int k = 0; k < net->n; k++) {
        qo_elem *e = &net->e[k];
        if (e->kind == QO_SER_L || e->kind == QO_SHUNT_L) {
            e->p[1] = wc * e->p[0] / q_l;
            e->p[2] = 1.0 / (wl * wl * e->p[0]);
        } else if (e->kind == QO_SER_C || e->kind == QO_SHUNT_C) {
            e->p[1] = esr_c;
            e->p[2] = 1.0 / (wcap * wcap * e->p[0]);
        }
    }
    return QO_OK;
}

/* ---- grids (pa-lpf-simulation.sch:59) ------------------------------------ */
int qo_grid_lin(double f0, double f1, int n, double *f)
{
    if (!f || n < 1) return QO_ERR_ARG;
    volatile double step = n > 1 ? (f1 - f0) / (double)(n - 1) : 0.0;
    for (int k = 0; k < n; k++) {
        volatile double prod = (double)k * step; /* keep mul and add separately rounded */
        f[k] = f0 + prod;
    }
    return QO_OK;
}
int qo_grid_log(double f0, double f1, int n, double *f)
{
    if (!f || n < 1 || !(f0 > 0) || !(f1 > 0)) return QO_ERR_ARG;
    double l0 = log(f0), step = n > 1 ? (log(f1) - l0) / (double)(n - 1) : 0.0;
    for (int k = 0; k < n; k++) f[k] = exp(l0 + (double)k * step);
    f[0] = f0;
    if (n > 1) f[n - 1] = f1;
    return QO_OK;
}

/* ---- host twins of the device perturbation stream ------------------------ */
void qo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    qo_philox_rounds(c, key[0], key[1]);
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
double qo_variate(uint64_t seed, uint64_t sample, uint32_t var, int dist)
{
    return qo_stream_variate(seed, sample, var, dist);
}
double qo_perturb_factor(uint64_t seed, uint64_t sample, uint32_t var, int dist, double tol)
{
    return QO_FMA(tol, qo_stream_variate(seed, sample, var, dist), 1.0);
}
