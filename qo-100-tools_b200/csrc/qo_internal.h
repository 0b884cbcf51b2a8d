/* qo_internal.h -- host-side internals shared by the C loader files and the CUDA TU */
#ifndef QO_INTERNAL_H
#define QO_INTERNAL_H
#include "qo100net.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QO_MAX_ELEMS 96
#define QO_TITLE_MAX 512
#define QO_MAX_BLK 8

/* a Touchstone two-port: n points, s[4*k + {0,1,2,3}] = S11, S21, S12, S22 at f[k] (file order) */
struct qo_s2p {
    int n;
    double z0;
    double *f;
    qo_c64 *s;
};

struct qo_net {
    int n;
    double rs, rl;
    qo_elem e[QO_MAX_ELEMS];
    char title[QO_TITLE_MAX];
    int nblk;                          /* S-parameter blocks referenced by QO_SBLOCK elements (owned copies) */
    struct qo_s2p *blk[QO_MAX_BLK];
};

#define QO_NODAL_MAX_UNK 32       /* node voltages + VCVS currents */
#define QO_NODAL_MAX_BR 96
#define QO_NODAL_MAX_PORTS 8
struct qo_nodal {
    int n_nodes, nb, np, nblk;
    qo_branch br[QO_NODAL_MAX_BR];
    int port_node[QO_NODAL_MAX_PORTS];
    double port_z0[QO_NODAL_MAX_PORTS];
    struct qo_s2p *blk[QO_MAX_BLK];
};

/* thread-local error detail */
void qo_set_error(const char *fmt, ...);
void qo_clear_error(void);

/* helpers used by the loaders */
char *qo_read_file(const char *path, size_t *len);                 /* malloc'd, NUL-terminated */
int qo_parse_value(const char *s, double *out, const char **unit); /* "4.700 pF" / "0.6 mm" / ".75e-3" */
qo_net *qo_net_alloc(void);

/* qo_s2p.c */
qo_s2p *qo_s2p_alloc(int n);
qo_s2p *qo_s2p_clone(const qo_s2p *a);
void qo_s2p_eval(const qo_s2p *b, double f, int polar, qo_c64 s[4]);      /* SPfile "linear" interpolation */
int qo_s_to_abcd(const qo_c64 s[4], double z0, qo_c64 abcd[4]);           /* 0 when S21 == 0 */

#ifdef __cplusplus
}
#endif
#endif
