/*
 * qo_dat.c -- Qucs dataset (.dat) reader / writer, SURVEY row N2.
 *
 * Layout as written by Qucs 0.0.19 for util/pa-lpf-simulation/pa-lpf-simulation.dat:1-35018:
 *   <Qucs Dataset 0.0.19>
 *   <indep NAME N> ... </indep>          one value per line, "  %+.20e"
 *   <dep NAME INDEP> ... </dep>          real "  %+.20e" or complex "  %+.20e{+|-}j%.20e"
 * With it a GPU sweep can be opened in Qucs next to the reference's own dataset, and the
 * reference dataset can be read without the Python fixture path.  21 significant digits
 * round-trip every double exactly, so read -> write reproduces the reference file byte for byte
 * (tests/test_host.py::test_qucs_dataset_round_trip).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "qo_internal.h"

#define DAT_NAME 64

typedef struct {
    char name[DAT_NAME], indep[DAT_NAME];   /* indep[0] == 0 for an independent variable */
    int n, is_complex;
    double *re, *im;
} dat_var;

struct qo_dat {
    int nvar, cap;
    dat_var *v;
};

int qo_dat_create(qo_dat **out)
{
    qo_clear_error();
    if (!out) return QO_ERR_ARG;
    qo_dat *d = (qo_dat *)calloc(1, sizeof *d);
    if (!d) return QO_ERR_NOMEM;
    *out = d;
    return QO_OK;
}

void qo_dat_free(qo_dat *d)
{
    if (!d) return;
    for (int i = 0; i < d->nvar; i++) { free(d->v[i].re); free(d->v[i].im); }
    free(d->v);
    free(d);
}

static dat_var *dat_find(const qo_dat *d, const char *name)
{
    for (int i = 0; i < d->nvar; i++)
        if (strcmp(d->v[i].name, name) == 0) return &d->v[i];
    return NULL;
}

static int dat_push(qo_dat *d, const char *name, const char *indep, const double *re, const double *im, int n, dat_var **slot)
{
    if (!d || !name || !*name || n < 0 || (n > 0 && !re && !slot)) return QO_ERR_ARG;
    if (strlen(name) >= DAT_NAME || (indep && strlen(indep) >= DAT_NAME) || strpbrk(name, " \t\n<>")) {
        qo_set_error("dataset variable name '%s' is empty, too long or contains blanks", name);
        return QO_ERR_ARG;
    }
    if (dat_find(d, name)) { qo_set_error("dataset already has a variable '%s'", name); return QO_ERR_ARG; }
    if (d->nvar == d->cap) {
        int nc = d->cap ? 2 * d->cap : 8;
        dat_var *nv = (dat_var *)realloc(d->v, (size_t)nc * sizeof *nv);
        if (!nv) return QO_ERR_NOMEM;
        d->v = nv; d->cap = nc;
    }
    dat_var *v = &d->v[d->nvar];
    memset(v, 0, sizeof *v);
    strcpy(v->name, name);
    if (indep) strcpy(v->indep, indep);
    v->n = n;
    v->is_complex = im != NULL;
    v->re = (double *)malloc((size_t)(n ? n : 1) * sizeof(double));
    v->im = (double *)calloc((size_t)(n ? n : 1), sizeof(double));
    if (!v->re || !v->im) { free(v->re); free(v->im); return QO_ERR_NOMEM; }
    if (re) memcpy(v->re, re, (size_t)n * sizeof(double));
    if (im) memcpy(v->im, im, (size_t)n * sizeof(double));
    d->nvar++;
    if (slot) *slot = v;
    return QO_OK;
}

int qo_dat_add_indep(qo_dat *d, const char *name, const double *v, int n)
{
    qo_clear_error();
    return dat_push(d, name, NULL, v, NULL, n, NULL);
}

int qo_dat_add_dep(qo_dat *d, const char *name, const char *indep, const double *re, const double *im, int n)
{
    qo_clear_error();
    if (!d || !indep) return QO_ERR_ARG;
    const dat_var *iv = dat_find(d, indep);
    if (!iv || iv->indep[0]) { qo_set_error("'%s' is not an independent variable of this dataset", indep); return QO_ERR_ARG; }
    if (iv->n != n) { qo_set_error("'%s' has %d points, '%s' has %d", name ? name : "?", n, indep, iv->n); return QO_ERR_ARG; }
    return dat_push(d, name, indep, re, im, n, NULL);
}

int qo_dat_count(const qo_dat *d) { return d ? d->nvar : QO_ERR_ARG; }

int qo_dat_info(const qo_dat *d, int i, const char **name, const char **indep, int *n, int *is_complex)
{
    if (!d || i < 0 || i >= d->nvar) return QO_ERR_ARG;
    if (name) *name = d->v[i].name;
    if (indep) *indep = d->v[i].indep;
    if (n) *n = d->v[i].n;
    if (is_complex) *is_complex = d->v[i].is_complex;
    return QO_OK;
}

int qo_dat_get(const qo_dat *d, const char *name, double *re, double *im, int cap)
{
    qo_clear_error();
    if (!d || !name) return QO_ERR_ARG;
    const dat_var *v = dat_find(d, name);
    if (!v) { qo_set_error("dataset has no variable '%s'", name); return QO_ERR_ARG; }
    int n = v->n < cap ? v->n : cap;
    if (re && n > 0) memcpy(re, v->re, (size_t)n * sizeof(double));
    if (im && n > 0) memcpy(im, v->im, (size_t)n * sizeof(double));
    return v->n;
}

int qo_dat_write(const qo_dat *d, const char *path)
{
    qo_clear_error();
    if (!d || !path) return QO_ERR_ARG;
    FILE *fp = fopen(path, "w");
    if (!fp) { qo_set_error("cannot create %s", path); return QO_ERR_IO; }
    fputs("<Qucs Dataset 0.0.19>\n", fp);
    for (int i = 0; i < d->nvar; i++) {
        const dat_var *v = &d->v[i];
        if (v->indep[0]) fprintf(fp, "<dep %s %s>\n", v->name, v->indep);
        else fprintf(fp, "<indep %s %d>\n", v->name, v->n);
        for (int k = 0; k < v->n; k++) {
            if (v->is_complex) {
                /* Qucs prints the imaginary part's sign in front of the 'j' */
                const double im = v->im[k];
                fprintf(fp, "  %+.20e%cj%.20e\n", v->re[k], signbit(im) ? '-' : '+', fabs(im));
            } else {
                fprintf(fp, "  %+.20e\n", v->re[k]);
            }
        }
        fputs(v->indep[0] ? "</dep>\n" : "</indep>\n", fp);
    }
    if (fclose(fp) != 0) { qo_set_error("write error on %s", path); return QO_ERR_IO; }
    return QO_OK;
}

/* one value: "+1.0e+07" or "-5.4e-05-j2.4e-03"; returns 1 real, 2 complex, 0 error */
static int dat_value(const char *s, double *re, double *im)
{
    char *end;
    *re = strtod(s, &end);
    if (end == s) return 0;
    *im = 0.0;
    while (*end == ' ' || *end == '\t') end++;
    if (*end == '\0' || *end == '\r' || *end == '\n') return 1;
    if ((*end == '+' || *end == '-') && end[1] == 'j') {
        const double sign = *end == '-' ? -1.0 : 1.0;
        char *e2;
        const double v = strtod(end + 2, &e2);
        if (e2 == end + 2) return 0;
        *im = sign * v;
        return 2;
    }
    return 0;
}

int qo_dat_read(const char *path, qo_dat **out)
{
    qo_clear_error();
    if (!path || !out) return QO_ERR_ARG;
    size_t len;
    char *txt = qo_read_file(path, &len);
    if (!txt) return QO_ERR_IO;
    qo_dat *d = NULL;
    int rc = qo_dat_create(&d);
    if (rc) { free(txt); return rc; }
    int line = 0, got = 0;
    dat_var *cur = NULL;
    char *save = NULL;
    for (char *ln = strtok_r(txt, "\n", &save); ln; ln = strtok_r(NULL, "\n", &save)) {
        line++;
        while (*ln == ' ' || *ln == '\t') ln++;
        if (!*ln || *ln == '\r') continue;
        if (line == 1) {
            if (strncmp(ln, "<Qucs Dataset", 13) != 0) { qo_set_error("%s:1: not a Qucs dataset", path); rc = QO_ERR_PARSE; break; }
            continue;
        }
        if (ln[0] == '<') {
            char name[DAT_NAME], third[DAT_NAME];
            if (strncmp(ln, "</", 2) == 0) {
                if (!cur) { qo_set_error("%s:%d: closing tag without a variable", path, line); rc = QO_ERR_PARSE; break; }
                if (got != cur->n) { qo_set_error("%s:%d: '%s' has %d values, header says %d", path, line, cur->name, got, cur->n); rc = QO_ERR_PARSE; break; }
                cur = NULL;
            } else if (sscanf(ln, "<indep %63s %63[^>]>", name, third) == 2) {
                int n = atoi(third);
                if (n < 0) { qo_set_error("%s:%d: bad point count", path, line); rc = QO_ERR_PARSE; break; }
                rc = dat_push(d, name, NULL, NULL, NULL, n, &cur);
                if (rc) break;
                got = 0;
            } else if (sscanf(ln, "<dep %63s %63[^>]>", name, third) == 2) {
                /* a dependency list may name several independents ("<dep X a b>"): the point count is their product */
                int n = 1;
                char deps[DAT_NAME];
                strcpy(deps, third);
                char *sv2 = NULL;
                for (char *t = strtok_r(deps, " ", &sv2); t; t = strtok_r(NULL, " ", &sv2)) {
                    const dat_var *iv = dat_find(d, t);
                    if (!iv) { qo_set_error("%s:%d: '%s' depends on unknown '%s'", path, line, name, t); rc = QO_ERR_PARSE; break; }
                    n *= iv->n;
                }
                if (rc) break;
                rc = dat_push(d, name, third, NULL, NULL, n, &cur);
                if (rc) break;
                got = 0;
            } else { qo_set_error("%s:%d: unrecognised tag", path, line); rc = QO_ERR_PARSE; break; }
            continue;
        }
        if (!cur) { qo_set_error("%s:%d: value outside a variable", path, line); rc = QO_ERR_PARSE; break; }
        if (got >= cur->n) { qo_set_error("%s:%d: '%s' has more values than its header says (%d)", path, line, cur->name, cur->n); rc = QO_ERR_PARSE; break; }
        double re, im;
        int kind = dat_value(ln, &re, &im);
        if (!kind) { qo_set_error("%s:%d: cannot parse value", path, line); rc = QO_ERR_PARSE; break; }
        if (kind == 2) cur->is_complex = 1;
        cur->re[got] = re; cur->im[got] = im;
        got++;
    }
    if (rc == QO_OK && cur) { qo_set_error("%s: variable '%s' is not closed", path, cur->name); rc = QO_ERR_PARSE; }
    free(txt);
    if (rc) { qo_dat_free(d); return rc; }
    *out = d;
    return QO_OK;
}

int qo_dat_from_sweep(const double *f, int nf, const qo_c64 *s11, const qo_c64 *s12, const qo_c64 *s21, const qo_c64 *s22, qo_dat **out)
{
    qo_clear_error();
    if (!f || nf <= 0 || !s11 || !s12 || !s21 || !s22 || !out) return QO_ERR_ARG;
    qo_dat *d = NULL;
    int rc = qo_dat_create(&d);
    if (rc) return rc;
    double *re = (double *)malloc((size_t)nf * sizeof(double)), *im = (double *)malloc((size_t)nf * sizeof(double));
    if (!re || !im) { free(re); free(im); qo_dat_free(d); return QO_ERR_NOMEM; }
    rc = qo_dat_add_indep(d, "frequency", f, nf);
    /* the two Eqn traces of pa-lpf-simulation.sch:60 first, then the S-matrix, as in the reference dataset */
    const qo_c64 *dbsrc[2] = { s11, s21 };
    const char *dbname[2] = { "S11_dB", "S21_dB" };
    for (int t = 0; t < 2 && rc == QO_OK; t++) {
        for (int k = 0; k < nf; k++) re[k] = 20.0 * log10(hypot(dbsrc[t][k].re, dbsrc[t][k].im));
        rc = qo_dat_add_dep(d, dbname[t], "frequency", re, NULL, nf);
    }
    const qo_c64 *ssrc[4] = { s11, s12, s21, s22 };
    const char *sname[4] = { "S[1,1]", "S[1,2]", "S[2,1]", "S[2,2]" };
    for (int t = 0; t < 4 && rc == QO_OK; t++) {
        for (int k = 0; k < nf; k++) { re[k] = ssrc[t][k].re; im[k] = ssrc[t][k].im; }
        rc = qo_dat_add_dep(d, sname[t], "frequency", re, im, nf);
    }
    free(re); free(im);
    if (rc) { qo_dat_free(d); return rc; }
    *out = d;
    return QO_OK;
}
