/*
 * qo_nodal.c -- netlist container for the N-port nodal solver: nodes 1..n (0 = ground),
 * two-terminal R / L / C branches with the same ESR/SRF parasitic forms as the cascade elements, ideal
 * voltage-controlled voltage sources, measured two-port blocks (Touchstone) with a reference terminal, and
 * numbered ports with real reference impedances.  Replaces what the Qucs netlister hands to qucsator for
 * util/pa-bias-simulation/pa-bias-simulation.sch:19-72.  The solver itself is the kernel in qo_nodal.cu.
 */
#include <stdlib.h>
#include <string.h>
#include "qo_internal.h"

int qo_nodal_create(int n_nodes, qo_nodal **out)
{
    qo_clear_error();
    if (!out || n_nodes < 1 || n_nodes > QO_NODAL_MAX_UNK) { qo_set_error("a nodal netlist needs 1..%d nodes", QO_NODAL_MAX_UNK); return QO_ERR_ARG; }
    qo_nodal *n = (qo_nodal *)calloc(1, sizeof *n);
    if (!n) return QO_ERR_NOMEM;
    n->n_nodes = n_nodes;
    *out = n;
    return QO_OK;
}

void qo_nodal_free(qo_nodal *n)
{
    if (!n) return;
    for (int i = 0; i < n->nblk; i++) qo_s2p_free(n->blk[i]);
    free(n);
}

static int node_ok(const qo_nodal *n, int v) { return v >= 0 && v <= n->n_nodes; }

int qo_nodal_add_branch(qo_nodal *n, const qo_branch *b)
{
    qo_clear_error();
    if (!n || !b) return QO_ERR_ARG;
    if (n->nb >= QO_NODAL_MAX_BR) { qo_set_error("too many branches (limit %d)", QO_NODAL_MAX_BR); return QO_ERR_RANGE; }
    int nn = 2;
    switch (b->kind) {
    case QO_NB_R: if (!(b->p[0] > 0)) { qo_set_error("resistor needs R > 0"); return QO_ERR_ARG; } break;
    case QO_NB_L: case QO_NB_C: if (!(b->p[0] > 0) || b->p[1] < 0 || b->p[2] < 0) { qo_set_error("L / C branch needs a positive value and non-negative parasitics"); return QO_ERR_ARG; } break;
    case QO_NB_VCVS: nn = 4; break;
    case QO_NB_SBLOCK:
        nn = 3;
        if (b->p[0] < 0 || b->p[0] >= n->nblk || !(b->p[2] > 0)) { qo_set_error("S-block branch refers to block %g, the netlist holds %d", b->p[0], n->nblk); return QO_ERR_ARG; }
        break;
    default: qo_set_error("unknown branch kind %d", b->kind); return QO_ERR_UNSUPPORTED;
    }
    for (int k = 0; k < nn; k++)
        if (!node_ok(n, b->node[k])) { qo_set_error("branch node %d out of range (0..%d)", b->node[k], n->n_nodes); return QO_ERR_ARG; }
    if (nn == 2 && b->node[0] == b->node[1]) { qo_set_error("branch shorted onto one node"); return QO_ERR_ARG; }
    int unknowns = n->n_nodes;
    for (int i = 0; i < n->nb; i++) if (n->br[i].kind == QO_NB_VCVS) unknowns++;
    if (b->kind == QO_NB_VCVS && unknowns + 1 > QO_NODAL_MAX_UNK) { qo_set_error("more than %d unknowns", QO_NODAL_MAX_UNK); return QO_ERR_RANGE; }
    n->br[n->nb++] = *b;
    return QO_OK;
}

int qo_nodal_add_port(qo_nodal *n, int node, double z0)
{
    qo_clear_error();
    if (!n) return QO_ERR_ARG;
    if (n->np >= QO_NODAL_MAX_PORTS) { qo_set_error("at most %d ports", QO_NODAL_MAX_PORTS); return QO_ERR_RANGE; }
    if (node < 1 || node > n->n_nodes || !(z0 > 0)) { qo_set_error("port needs a non-ground node and z0 > 0"); return QO_ERR_ARG; }
    n->port_node[n->np] = node; n->port_z0[n->np] = z0;
    return ++n->np;
}

int qo_nodal_add_sblock(qo_nodal *n, const qo_s2p *blk, int *index)
{
    qo_clear_error();
    if (!n || !blk) return QO_ERR_ARG;
    if (n->nblk >= QO_MAX_BLK) { qo_set_error("at most %d S-parameter blocks", QO_MAX_BLK); return QO_ERR_RANGE; }
    n->blk[n->nblk] = qo_s2p_clone(blk);
    if (!n->blk[n->nblk]) return QO_ERR_NOMEM;
    if (index) *index = n->nblk;
    n->nblk++;
    return QO_OK;
}

int qo_nodal_num_nodes(const qo_nodal *n) { return n ? n->n_nodes : QO_ERR_ARG; }
int qo_nodal_num_ports(const qo_nodal *n) { return n ? n->np : QO_ERR_ARG; }
int qo_nodal_num_branches(const qo_nodal *n) { return n ? n->nb : QO_ERR_ARG; }

int qo_nodal_get_branches(const qo_nodal *n, qo_branch *out, int cap)
{
    if (!n) return QO_ERR_ARG;
    const int m = n->nb < cap ? n->nb : cap;
    if (out && m > 0) memcpy(out, n->br, (size_t)m * sizeof(qo_branch));
    return n->nb;
}

int qo_nodal_get_ports(const qo_nodal *n, int *node, double *z0, int cap)
{
    if (!n) return QO_ERR_ARG;
    for (int k = 0; k < n->np && k < cap; k++) { if (node) node[k] = n->port_node[k]; if (z0) z0[k] = n->port_z0[k]; }
    return n->np;
}
