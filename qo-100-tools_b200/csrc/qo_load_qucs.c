/*
 * qo_load_qucs.c -- Qucs 0.0.19 schematic (.sch) and QucsTranscalc (.trc) readers.
 *
 * .sch: util/pa-lpf-simulation/pa-lpf-simulation.sch:19-61 (components),
 *       :63-103 (wires).  A component line is
 *   <TYPE NAME active x y tx ty mirror rot "prop" show "prop" show ...>
 * and a wire line is <x1 y1 x2 y2 "" 0 0 0 "">.  The cascade is recovered from
 * geometry: pin positions follow from (x, y, rot), wires short every node lying
 * on them, and the 2-port chain is walked from port 1 to port 2; an MTEE's third
 * pin opens a side arm that must end in an MOPEN.
 * .trc: util/directional-couplers/ *.trc:5-21 (key value unit lines).
 */
#define _GNU_SOURCE
#include "qo_internal.h"
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXC 128
#define MAXP 8
#define MAXW 256
#define MAXPIN 4

typedef struct {
    char type[16], name[32];
    int x, y, mirror, rot, nprop;
    char prop[MAXP][128];
    int npin, px[MAXPIN], py[MAXPIN], node[MAXPIN];
    int used;
} comp_t;

typedef struct { char name[32]; double value; } const_t;

typedef struct {
    comp_t c[MAXC];
    int nc;
    int w[MAXW][4];
    int nw;
    const_t k[32];
    int nk;
} sch_t;

/* rotate an offset by rot*90 deg, Qucs screen convention (y down): (dx,dy)->(dy,-dx) */
static void rot_off(int rot, int mirror, int *dx, int *dy)
{
    if (mirror) *dy = -*dy;
    for (int i = 0; i < (rot & 3); i++) { int t = *dx; *dx = *dy; *dy = -t; }
}

static void set_pins(comp_t *c)
{
    static const int two[2][2] = { { -30, 0 }, { 30, 0 } };
    static const int corn[2][2] = { { -30, 0 }, { 0, 30 } };
    static const int tee[3][2] = { { -30, 0 }, { 30, 0 }, { 0, 30 } };
    static const int one[1][2] = { { -30, 0 } };
    static const int gnd[1][2] = { { 0, 0 } };
    static const int spf[3][2] = { { -30, 0 }, { 30, 0 }, { 0, 30 } };                       /* port 1, port 2, reference */
    static const int vcvs[4][2] = { { -30, -30 }, { 30, -30 }, { 30, 30 }, { -30, 30 } };    /* in+, out+, out-, in- */
    const int (*tab)[2] = NULL;
    c->npin = 0;
    if (!strcmp(c->type, "MLIN") || !strcmp(c->type, "Pac") || !strcmp(c->type, "R") || !strcmp(c->type, "L") ||
        !strcmp(c->type, "C") || !strcmp(c->type, "TLIN")) { tab = two; c->npin = 2; }
    else if (!strcmp(c->type, "MCORN")) { tab = corn; c->npin = 2; }
    else if (!strcmp(c->type, "MTEE")) { tab = tee; c->npin = 3; }
    else if (!strcmp(c->type, "MOPEN")) { tab = one; c->npin = 1; }
    else if (!strcmp(c->type, "GND")) { tab = gnd; c->npin = 1; }
    else if (!strcmp(c->type, "SPfile")) { tab = spf; c->npin = 3; }
    else if (!strcmp(c->type, "VCVS")) { tab = vcvs; c->npin = 4; }
    for (int i = 0; i < c->npin; i++) {
        int dx = tab[i][0], dy = tab[i][1];
        rot_off(c->rot, c->mirror, &dx, &dy);
        c->px[i] = c->x + dx;
        c->py[i] = c->y + dy;
    }
}

static int parse_component(const char *line, comp_t *c)
{
    memset(c, 0, sizeof *c);
    int active, tx, ty, off = 0;
    if (sscanf(line, " <%15s %31s %d %d %d %d %d %d %d%n", c->type, c->name, &active, &c->x, &c->y, &tx, &ty, &c->mirror,
               &c->rot, &off) < 9)
        return 0;
    const char *p = line + off;
    while (c->nprop < MAXP && (p = strchr(p, '"')) != NULL) {
        const char *q = strchr(p + 1, '"');
        if (!q) break;
        size_t n = (size_t)(q - p - 1);
        if (n > 127) n = 127;
        memcpy(c->prop[c->nprop], p + 1, n);
        c->prop[c->nprop][n] = 0;
        c->nprop++;
        p = q + 1;
    }
    set_pins(c);
    return 1;
}

static int parse_sch(const char *path, sch_t *s)
{
    size_t len;
    char *buf = qo_read_file(path, &len);
    if (!buf) return QO_ERR_IO;
    memset(s, 0, sizeof *s);
    if (strncmp(buf, "<Qucs Schematic", 15)) { free(buf); qo_set_error("%s: not a Qucs schematic", path); return QO_ERR_PARSE; }
    int section = 0, lineno = 0;
    char *save = NULL;
    for (char *line = strtok_r(buf, "\n", &save); line; line = strtok_r(NULL, "\n", &save)) {
        lineno++;
        if (strstr(line, "<Components>")) { section = 1; continue; }
        if (strstr(line, "</Components>")) { section = 0; continue; }
        if (strstr(line, "<Wires>")) { section = 2; continue; }
        if (strstr(line, "</Wires>")) { section = 0; continue; }
        if (section == 1) {
            if (s->nc >= MAXC) { free(buf); return QO_ERR_RANGE; }
            if (!parse_component(line, &s->c[s->nc])) { free(buf); qo_set_error("%s:%d: bad component line", path, lineno); return QO_ERR_PARSE; }
            comp_t *c = &s->c[s->nc++];
            if (!strcmp(c->type, "Eqn")) {
                for (int i = 0; i < c->nprop; i++) {
                    char *eq = strchr(c->prop[i], '=');
                    if (!eq) continue;
                    char *endp;
                    double v = strtod(eq + 1, &endp);
                    if (endp == eq + 1 || *endp) continue;   /* only plain numeric constants */
                    if (s->nk < 32) {
                        size_t n = (size_t)(eq - c->prop[i]);
                        if (n > 31) n = 31;
                        memcpy(s->k[s->nk].name, c->prop[i], n);
                        s->k[s->nk].name[n] = 0;
                        s->k[s->nk].value = v;
                        s->nk++;
                    }
                }
            }
        } else if (section == 2) {
            if (s->nw >= MAXW) { free(buf); return QO_ERR_RANGE; }
            if (sscanf(line, " <%d %d %d %d", &s->w[s->nw][0], &s->w[s->nw][1], &s->w[s->nw][2], &s->w[s->nw][3]) == 4) s->nw++;
        }
    }
    free(buf);
    return QO_OK;
}

static int prop_value(const sch_t *s, const char *txt, double *out)
{
    if (qo_parse_value(txt, out, NULL) == QO_OK) return QO_OK;
    for (int i = 0; i < s->nk; i++)
        if (!strcmp(s->k[i].name, txt)) { *out = s->k[i].value; return QO_OK; }
    qo_set_error("cannot evaluate property '%s'", txt);
    return QO_ERR_PARSE;
}

/* ---- node extraction: union-find over pin and wire-end coordinates -------- */
typedef struct { int x, y, parent; } pt_t;
static int uf_find(pt_t *p, int i) { while (p[i].parent != i) { p[i].parent = p[p[i].parent].parent; i = p[i].parent; } return i; }
static void uf_union(pt_t *p, int a, int b) { a = uf_find(p, a); b = uf_find(p, b); if (a != b) p[b].parent = a; }
static int pt_get(pt_t *p, int *n, int x, int y)
{
    for (int i = 0; i < *n; i++) if (p[i].x == x && p[i].y == y) return i;
    p[*n].x = x; p[*n].y = y; p[*n].parent = *n;
    return (*n)++;
}
static int on_segment(const int w[4], int x, int y)
{
    int x1 = w[0] < w[2] ? w[0] : w[2], x2 = w[0] < w[2] ? w[2] : w[0];
    int y1 = w[1] < w[3] ? w[1] : w[3], y2 = w[1] < w[3] ? w[3] : w[1];
    if (w[0] == w[2]) return x == w[0] && y >= y1 && y <= y2;
    if (w[1] == w[3]) return y == w[1] && x >= x1 && x <= x2;
    return (x == w[0] && y == w[1]) || (x == w[2] && y == w[3]);
}

/* every pin gets the representative index of its electrical node (pins and wire ends that coincide or lie on
 * a common wire are one node) */
static void extract_nodes(sch_t *s)
{
    static __thread pt_t pts[MAXC * MAXPIN + 2 * MAXW];
    int np = 0;
    for (int i = 0; i < s->nc; i++)
        for (int k = 0; k < s->c[i].npin; k++) s->c[i].node[k] = pt_get(pts, &np, s->c[i].px[k], s->c[i].py[k]);
    for (int i = 0; i < s->nw; i++) {
        int a = pt_get(pts, &np, s->w[i][0], s->w[i][1]);
        int b = pt_get(pts, &np, s->w[i][2], s->w[i][3]);
        uf_union(pts, a, b);
    }
    for (int i = 0; i < s->nw; i++) {
        int a = pt_get(pts, &np, s->w[i][0], s->w[i][1]);
        for (int j = 0; j < np; j++)
            if (on_segment(s->w[i], pts[j].x, pts[j].y)) uf_union(pts, a, j);
    }
    for (int i = 0; i < s->nc; i++)
        for (int k = 0; k < s->c[i].npin; k++) s->c[i].node[k] = uf_find(pts, s->c[i].node[k]);
}

static comp_t *find_at(sch_t *s, int node, int *pin)
{
    for (int i = 0; i < s->nc; i++) {
        comp_t *c = &s->c[i];
        if (c->used || !strcmp(c->type, "Pac") || !strcmp(c->type, "GND") || c->npin == 0) continue;
        for (int k = 0; k < c->npin; k++)
            if (c->node[k] == node) { *pin = k; return c; }
    }
    return NULL;
}

static int emit(qo_net *net, int kind, double p0, double p1, double p2)
{
    if (net->n >= QO_MAX_ELEMS) return QO_ERR_RANGE;
    qo_elem *e = &net->e[net->n++];
    memset(e, 0, sizeof *e);
    e->kind = kind; e->p[0] = p0; e->p[1] = p1; e->p[2] = p2;
    return QO_OK;
}

/* walk a chain of 2-pin elements from `node`; stops at `goal` (main path) or at an MOPEN (side arm) */
static int walk(sch_t *s, qo_net *net, int node, int goal, int side)
{
    int guard = 0;
    while (node != goal && guard++ < 4 * MAXC) {
        int pin;
        comp_t *c = find_at(s, node, &pin);
        if (!c) { qo_set_error("open circuit: nothing continues the chain at node %d", node); return QO_ERR_UNSUPPORTED; }
        c->used = 1;
        double a, b, d;
        int rc;
        if (!strcmp(c->type, "MLIN")) {
            if ((rc = prop_value(s, c->prop[1], &a)) || (rc = prop_value(s, c->prop[2], &b))) return rc;
            if ((rc = emit(net, QO_MLIN, a, b, 0))) return rc;
            node = c->node[1 - pin];
        } else if (!strcmp(c->type, "MCORN")) {
            if ((rc = prop_value(s, c->prop[1], &a))) return rc;
            if ((rc = emit(net, QO_MCORN, a, 0, 0))) return rc;
            node = c->node[1 - pin];
        } else if (!strcmp(c->type, "MOPEN")) {
            if (!side) { qo_set_error("%s: open end on the main path", c->name); return QO_ERR_UNSUPPORTED; }
            if ((rc = prop_value(s, c->prop[1], &a))) return rc;
            return emit(net, QO_MOPEN, a, 0, 0);
        } else if (!strcmp(c->type, "MTEE")) {
            if (side || pin == 2) { qo_set_error("%s: tee entered through its side arm / nested tee", c->name); return QO_ERR_UNSUPPORTED; }
            if ((rc = prop_value(s, c->prop[1 + pin], &a)) || (rc = prop_value(s, c->prop[1 + (1 - pin)], &b)) ||
                (rc = prop_value(s, c->prop[3], &d)))
                return rc;
            if ((rc = emit(net, QO_MTEE, a, b, d))) return rc;
            if ((rc = walk(s, net, c->node[2], -1, 1))) return rc;
            node = c->node[1 - pin];
        } else {
            qo_set_error("%s: component type %s is not a supported 2-port cascade element", c->name, c->type);
            return QO_ERR_UNSUPPORTED;
        }
    }
    if (side) { qo_set_error("side arm does not end in an MOPEN"); return QO_ERR_UNSUPPORTED; }
    return node == goal ? QO_OK : QO_ERR_UNSUPPORTED;
}

int qo_net_load_qucs_sch(const char *path, qo_net **out)
{
    qo_clear_error();
    if (!path || !out) return QO_ERR_ARG;
    sch_t *s = (sch_t *)malloc(sizeof(sch_t));
    if (!s) return QO_ERR_NOMEM;
    qo_net *net = NULL;
    int rc = parse_sch(path, s);
    if (rc) goto done;

    extract_nodes(s);

    /* ground net(s) and the two ports */
    int gnd[16], ngnd = 0;
    for (int i = 0; i < s->nc; i++)
        if (!strcmp(s->c[i].type, "GND") && ngnd < 16) gnd[ngnd++] = s->c[i].node[0];
    int port_node[3] = { -1, -1, -1 };
    double port_z[3] = { 0, 50, 50 };
    for (int i = 0; i < s->nc; i++) {
        comp_t *c = &s->c[i];
        if (strcmp(c->type, "Pac")) continue;
        int num = atoi(c->prop[0]);
        if (num < 1 || num > 2) { qo_set_error("%s: only 2-port schematics are supported (port %d)", path, num); rc = QO_ERR_UNSUPPORTED; goto done; }
        if ((rc = prop_value(s, c->prop[1], &port_z[num]))) goto done;
        for (int k = 0; k < 2; k++) {
            int isg = 0;
            for (int g = 0; g < ngnd; g++) if (c->node[k] == gnd[g]) isg = 1;
            if (!isg) port_node[num] = c->node[k];
        }
    }
    if (port_node[1] < 0 || port_node[2] < 0) { qo_set_error("%s: need Pac ports 1 and 2, each with one grounded pin", path); rc = QO_ERR_UNSUPPORTED; goto done; }

    net = qo_net_alloc();
    if (!net) { rc = QO_ERR_NOMEM; goto done; }
    net->rs = port_z[1]; net->rl = port_z[2];
    /* substrate first: every microstrip component must reference the same SUBST */
    const comp_t *sub = NULL;
    for (int i = 0; i < s->nc; i++) if (!strcmp(s->c[i].type, "SUBST")) { if (sub) { qo_set_error("%s: more than one SUBST", path); rc = QO_ERR_UNSUPPORTED; goto done; } sub = &s->c[i]; }
    if (sub) {
        qo_elem *e = &net->e[net->n++];
        e->kind = QO_SUBST;
        for (int k = 0; k < 6; k++) if ((rc = prop_value(s, sub->prop[k], &e->p[k]))) goto done;
        for (int i = 0; i < s->nc; i++)
            if (s->c[i].type[0] == 'M' && s->c[i].npin && strcmp(s->c[i].prop[0], sub->name)) { qo_set_error("%s: %s references unknown substrate %s", path, s->c[i].name, s->c[i].prop[0]); rc = QO_ERR_PARSE; goto done; }
    }
    rc = walk(s, net, port_node[1], port_node[2], 0);
    if (rc) goto done;
    for (int i = 0; i < s->nc; i++) {
        comp_t *c = &s->c[i];
        if (!c->used && c->npin && strcmp(c->type, "Pac") && strcmp(c->type, "GND")) { qo_set_error("%s: component %s is not on the port-1 -> port-2 cascade", path, c->name); rc = QO_ERR_UNSUPPORTED; goto done; }
    }
    snprintf(net->title, sizeof net->title, "qucs:%s", path);
    *out = net;
    net = NULL;
done:
    free(s);
    qo_net_free(net);
    return rc;
}

int qo_qucs_sch_sweep(const char *path, int *type, double *f0, double *f1, int *n)
{
    qo_clear_error();
    if (!path) return QO_ERR_ARG;
    sch_t *s = (sch_t *)malloc(sizeof(sch_t));
    if (!s) return QO_ERR_NOMEM;
    int rc = parse_sch(path, s);
    if (!rc) {
        rc = QO_ERR_PARSE;
        qo_set_error("%s: no .SP simulation", path);
        for (int i = 0; i < s->nc; i++) {
            comp_t *c = &s->c[i];
            if (strcmp(c->type, ".SP")) continue;
            double a, b, cnt;
            if (prop_value(s, c->prop[1], &a) || prop_value(s, c->prop[2], &b) || prop_value(s, c->prop[3], &cnt)) break;
            if (type) *type = !strcmp(c->prop[0], "log");
            if (f0) *f0 = a;
            if (f1) *f1 = b;
            if (n) *n = (int)cnt;
            rc = QO_OK;
            qo_clear_error();
            break;
        }
    }
    free(s);
    return rc;
}

/* ---- QucsTranscalc .trc --------------------------------------------------- */
int qo_cpl_load_trc(const char *path, double *z0e, double *z0o, double *ang_deg, double *f0_hz, double phys[8])
{
    qo_clear_error();
    if (!path) return QO_ERR_ARG;
    size_t len;
    char *buf = qo_read_file(path, &len);
    if (!buf) return QO_ERR_IO;
    int rc = QO_OK, got = 0;
    if (!strstr(buf, "<CoupledMicrostrip>")) { qo_set_error("%s: no <CoupledMicrostrip> block", path); rc = QO_ERR_PARSE; goto done; }
    static const char *keys[] = { "Er", "H", "H_t", "T", "W", "S", "L", "Tand", "Freq", "Z0e", "Z0o", "Ang_l" };
    double v[12] = { 0 };
    char *save = NULL;
    for (char *line = strtok_r(strstr(buf, "<CoupledMicrostrip>"), "\n", &save); line; line = strtok_r(NULL, "\n", &save)) {
        char key[32];
        int off = 0;
        if (sscanf(line, " %31s %n", key, &off) < 1 || key[0] == '<' || key[0] == '#') continue;
        for (int k = 0; k < 12; k++) {
            if (strcmp(key, keys[k])) continue;
            if (qo_parse_value(line + off, &v[k], NULL)) { qo_set_error("%s: bad value for %s", path, key); rc = QO_ERR_PARSE; goto done; }
            got |= 1 << k;
        }
    }
    if ((got & 0xF00) != 0xF00) { qo_set_error("%s: Freq/Z0e/Z0o/Ang_l missing", path); rc = QO_ERR_PARSE; goto done; }
    if (z0e) *z0e = v[9];
    if (z0o) *z0o = v[10];
    if (ang_deg) *ang_deg = v[11];
    if (f0_hz) *f0_hz = v[8];
    if (phys) for (int k = 0; k < 8; k++) phys[k] = v[k];
done:
    free(buf);
    return rc;
}


/* ---- general netlist for the N-port nodal solver ---------------------------------------
 * util/pa-bias-simulation/pa-bias-simulation.sch:19-72 and util/preamp-bias-simulation/preamp-bias-simulation.sch:
 * R, C, L, GND, Pac (numbered ports, one pin grounded), VCVS (:40,59) and SPfile (:39, the Touchstone file is
 * looked up by its base name next to the schematic -- the stored path is the author's home directory). */
int qo_nodal_load_qucs_sch(const char *path, qo_nodal **out)
{
    qo_clear_error();
    if (!path || !out) return QO_ERR_ARG;
    sch_t *s = (sch_t *)malloc(sizeof(sch_t));
    if (!s) return QO_ERR_NOMEM;
    qo_nodal *nd = NULL;
    int rc = parse_sch(path, s);
    if (rc) goto done;
    extract_nodes(s);
    /* number the nodes: every GND pin's node is 0, the others 1.. in order of first appearance */
    int rep[MAXC * MAXPIN], num[MAXC * MAXPIN], nrep = 0, nnode = 0;
    for (int i = 0; i < s->nc; i++)
        if (!strcmp(s->c[i].type, "GND")) { rep[nrep] = s->c[i].node[0]; num[nrep++] = 0; }
    for (int i = 0; i < s->nc; i++)
        for (int k = 0; k < s->c[i].npin; k++) {
            int found = -1;
            for (int j = 0; j < nrep; j++) if (rep[j] == s->c[i].node[k]) found = j;
            if (found < 0) { rep[nrep] = s->c[i].node[k]; num[nrep] = ++nnode; found = nrep++; }
            s->c[i].node[k] = num[found];
        }
    if ((rc = qo_nodal_create(nnode, &nd))) goto done;
    /* ports, in the order of their Pac numbers */
    for (int want = 1; want <= 16; want++) {
        int hit = 0;
        for (int i = 0; i < s->nc; i++) {
            comp_t *c = &s->c[i];
            if (strcmp(c->type, "Pac") || atoi(c->prop[0]) != want) continue;
            double z0;
            if ((rc = prop_value(s, c->prop[1], &z0))) goto done;
            if ((c->node[0] == 0) == (c->node[1] == 0)) { qo_set_error("%s: port %s needs exactly one grounded pin", path, c->name); rc = QO_ERR_UNSUPPORTED; goto done; }
            if ((rc = qo_nodal_add_port(nd, c->node[0] ? c->node[0] : c->node[1], z0)) < 0) goto done;
            rc = QO_OK;
            hit = 1;
        }
        if (!hit) break;
    }
    for (int i = 0; i < s->nc; i++) {
        comp_t *c = &s->c[i];
        qo_branch b;
        memset(&b, 0, sizeof b);
        if (!strcmp(c->type, "R") || !strcmp(c->type, "C") || !strcmp(c->type, "L")) {
            b.kind = c->type[0] == 'R' ? QO_NB_R : c->type[0] == 'C' ? QO_NB_C : QO_NB_L;
            b.node[0] = c->node[0]; b.node[1] = c->node[1];
            if ((rc = prop_value(s, c->prop[0], &b.p[0]))) goto done;
        } else if (!strcmp(c->type, "VCVS")) {
            b.kind = QO_NB_VCVS;
            for (int k = 0; k < 4; k++) b.node[k] = c->node[k];
            if ((rc = prop_value(s, c->prop[0], &b.p[0])) || (rc = prop_value(s, c->prop[1], &b.p[1]))) goto done;
        } else if (!strcmp(c->type, "SPfile")) {
            if (strcmp(c->prop[2], "linear")) { qo_set_error("%s: %s uses '%s' interpolation; only 'linear' is restated", path, c->name, c->prop[2]); rc = QO_ERR_UNSUPPORTED; goto done; }
            if (atoi(c->prop[4]) != 2) { qo_set_error("%s: %s is not a 2-port file", path, c->name); rc = QO_ERR_UNSUPPORTED; goto done; }
            char file[1024];
            const char *base = strrchr(c->prop[0], '/'), *dirend = strrchr(path, '/');
            base = base ? base + 1 : c->prop[0];
            snprintf(file, sizeof file, "%.*s%s", dirend ? (int)(dirend - path + 1) : 0, path, base);
            qo_s2p *blk = NULL;
            if ((rc = qo_s2p_load(file, &blk))) goto done;
            int idx = -1;
            rc = qo_nodal_add_sblock(nd, blk, &idx);
            const double z0 = qo_s2p_z0(blk);
            qo_s2p_free(blk);
            if (rc) goto done;
            b.kind = QO_NB_SBLOCK;
            for (int k = 0; k < 3; k++) b.node[k] = c->node[k];
            b.p[0] = idx; b.p[1] = !strcmp(c->prop[1], "polar"); b.p[2] = z0;
        } else if (!strcmp(c->type, "GND") || !strcmp(c->type, "Pac") || c->npin == 0) {
            continue;                                   /* .SP, Eqn and other annotations carry no pins */
        } else { qo_set_error("%s: component type %s (%s) is not supported by the nodal loader", path, c->type, c->name); rc = QO_ERR_UNSUPPORTED; goto done; }
        if ((rc = qo_nodal_add_branch(nd, &b))) goto done;
    }
    if (qo_nodal_num_ports(nd) < 1) { qo_set_error("%s: no Pac ports", path); rc = QO_ERR_UNSUPPORTED; goto done; }
    *out = nd;
    nd = NULL;
done:
    qo_nodal_free(nd);
    free(s);
    return rc;
}
