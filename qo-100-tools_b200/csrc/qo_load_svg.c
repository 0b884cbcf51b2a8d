/*
 * qo_load_svg.c -- rf-tools.com LC-filter SVG export -> element list.
 *
 * Input artefacts (reference tree): util/if-bandpass-filter/schematic.svg:174-217,
 * util/gpsdo-ouput-filters/10M/schematic.svg:174-231, docs/gpsdo-filters/{15M,40M,
 * 60M}.svg:174-241, docs/upconverter/upconverter-lol-filter.svg:174-249.
 *
 * Format: after </defs>, one <use xlink:href="#SYMBOL"> per branch in electrical
 * order (source -> load), then <text> elements in (label, value) pairs in the same
 * order: "RS","50.00 Ω","C1","4.700 pF",...  Two-part branches (lc_*) carry two
 * pairs; which one is L and which is C is told by the label's first letter.
 */
#define _GNU_SOURCE
#include "qo_internal.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char label[16]; double value; } lv_t;

static int text_content(const char *p, const char *end, char *out, size_t cap, const char **next)
{
    const char *t = strstr(p, "<text");
    if (!t || t >= end) return 0;
    const char *gt = strchr(t, '>');
    if (!gt) return 0;
    const char *close = strstr(gt, "</text>");
    if (!close) return 0;
    size_t n = (size_t)(close - gt - 1);
    if (n >= cap) n = cap - 1;
    memcpy(out, gt + 1, n);
    out[n] = 0;
    *next = close + 7;
    return 1;
}

static int is_label(const char *s)
{
    if (!strcmp(s, "RS") || !strcmp(s, "RL")) return 1;
    if ((s[0] == 'L' || s[0] == 'C' || s[0] == 'R') && isdigit((unsigned char)s[1])) {
        for (const char *q = s + 1; *q; q++) if (!isdigit((unsigned char)*q)) return 0;
        return 1;
    }
    return 0;
}

int qo_net_load_rftools_svg(const char *path, qo_net **out)
{
    qo_clear_error();
    if (!path || !out) return QO_ERR_ARG;
    size_t len;
    char *buf = qo_read_file(path, &len);
    if (!buf) return QO_ERR_IO;
    int rc = QO_OK;
    qo_net *net = NULL;
    const char *body = strstr(buf, "</defs>");
    if (!body) { qo_set_error("%s: no </defs> (not an rf-tools export?)", path); rc = QO_ERR_PARSE; goto done; }
    const char *end = buf + len;

    /* 1. label/value pairs, in document order */
    lv_t lv[2 * QO_MAX_ELEMS];
    int nlv = 0;
    char txt[256], title[QO_TITLE_MAX] = "";
    const char *p = body, *next;
    char pending[16] = "";
    while (text_content(p, end, txt, sizeof txt, &next)) {
        p = next;
        if (pending[0]) {
            double v;
            if (qo_parse_value(txt, &v, NULL) != QO_OK) { qo_set_error("%s: bad value '%s' for %s", path, txt, pending); rc = QO_ERR_PARSE; goto done; }
            if (nlv >= 2 * QO_MAX_ELEMS) { rc = QO_ERR_RANGE; goto done; }
            snprintf(lv[nlv].label, sizeof lv[nlv].label, "%s", pending);
            lv[nlv].value = v;
            nlv++;
            pending[0] = 0;
        } else if (is_label(txt)) {
            snprintf(pending, sizeof pending, "%s", txt);
        } else if (txt[0] && !strstr(txt, "rf-tools.com")) {
            size_t have = strlen(title);
            snprintf(title + have, sizeof title - have, "%s%s", have ? "; " : "", txt);
        }
    }

    /* 2. branches */
    net = qo_net_alloc();
    if (!net) { rc = QO_ERR_NOMEM; goto done; }
    snprintf(net->title, sizeof net->title, "%s", title);
    int ilv = 0;
    p = body;
    while ((p = strstr(p, "xlink:href=\"#")) != NULL && p < end) {
        p += 13;
        char sym[64];
        size_t k = 0;
        while (p[k] && p[k] != '"' && k < sizeof sym - 1) { sym[k] = p[k]; k++; }
        sym[k] = 0;
        /* the *_only and *_half variants draw the same electrical branch */
        char *suffix;
        if ((suffix = strstr(sym, "_only")) != NULL) *suffix = 0;
        if ((suffix = strstr(sym, "_half")) != NULL) *suffix = 0;
        int kind = 0, npairs = 1;
        if (!strncmp(sym, "s_branch", 8) || !strncmp(sym, "bare_branch", 11) || !strncmp(sym, "open_branch", 11)) continue;
        if (!strcmp(sym, "r_branch_series")) kind = -1;       /* RS */
        else if (!strcmp(sym, "r_branch_shunt")) kind = -2;   /* RL */
        else if (!strcmp(sym, "l_branch_series")) kind = QO_SER_L;
        else if (!strcmp(sym, "c_branch_series")) kind = QO_SER_C;
        else if (!strcmp(sym, "l_branch_shunt")) kind = QO_SHUNT_L;
        else if (!strcmp(sym, "c_branch_shunt")) kind = QO_SHUNT_C;
        else if (!strcmp(sym, "lc_branch_series_series")) { kind = QO_SER_LC_SER; npairs = 2; }
        else if (!strcmp(sym, "lc_branch_series_parallel")) { kind = QO_SER_LC_PAR; npairs = 2; }
        else if (!strcmp(sym, "lc_branch_shunt_series")) { kind = QO_SHUNT_LC_SER; npairs = 2; }
        else if (!strcmp(sym, "lc_branch_shunt_parallel")) { kind = QO_SHUNT_LC_PAR; npairs = 2; }
        else { qo_set_error("%s: unknown branch symbol '%s'", path, sym); rc = QO_ERR_UNSUPPORTED; goto done; }
        if (ilv + npairs > nlv) { qo_set_error("%s: branch '%s' has no value text", path, sym); rc = QO_ERR_PARSE; goto done; }
        if (kind == -1) { net->rs = lv[ilv++].value; continue; }
        if (kind == -2) { net->rl = lv[ilv++].value; continue; }
        if (net->n >= QO_MAX_ELEMS) { rc = QO_ERR_RANGE; goto done; }
        qo_elem *e = &net->e[net->n++];
        e->kind = kind;
        if (npairs == 1) {
            char want = (kind == QO_SER_L || kind == QO_SHUNT_L) ? 'L' : 'C';
            if (lv[ilv].label[0] != want) { qo_set_error("%s: label %s does not fit branch %s", path, lv[ilv].label, sym); rc = QO_ERR_PARSE; goto done; }
            e->p[0] = lv[ilv++].value;
        } else {
            int gotl = 0, gotc = 0;
            for (int j = 0; j < 2; j++, ilv++) {
                if (lv[ilv].label[0] == 'L') { e->p[0] = lv[ilv].value; gotl = 1; }
                else if (lv[ilv].label[0] == 'C') { e->p[1] = lv[ilv].value; gotc = 1; }
            }
            if (!gotl || !gotc) { qo_set_error("%s: branch %s needs one L and one C", path, sym); rc = QO_ERR_PARSE; goto done; }
        }
    }
    if (net->n == 0) { qo_set_error("%s: no branches found", path); rc = QO_ERR_PARSE; goto done; }
    if (ilv != nlv) { qo_set_error("%s: %d value labels left over", path, nlv - ilv); rc = QO_ERR_PARSE; goto done; }
    if (!(net->rs > 0) || !(net->rl > 0)) { qo_set_error("%s: bad terminations", path); rc = QO_ERR_PARSE; goto done; }
    *out = net;
    net = NULL;
done:
    free(buf);
    qo_net_free(net);
    return rc;
}
