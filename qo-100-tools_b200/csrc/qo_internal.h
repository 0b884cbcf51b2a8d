/* qo_internal.h -- host-side internals shared by the C loader files and the CUDA TU */
#ifndef QO_INTERNAL_H
#define QO_INTERNAL_H
#include "qo100net.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QO_MAX_ELEMS 96
#define QO_TITLE_MAX 512

struct qo_net {
    int n;
    double rs, rl;
    qo_elem e[QO_MAX_ELEMS];
    char title[QO_TITLE_MAX];
};

/* thread-local error detail */
void qo_set_error(const char *fmt, ...);
void qo_clear_error(void);

/* helpers used by the loaders */
char *qo_read_file(const char *path, size_t *len);                 /* malloc'd, NUL-terminated */
int qo_parse_value(const char *s, double *out, const char **unit); /* "4.700 pF" / "0.6 mm" / ".75e-3" */
qo_net *qo_net_alloc(void);

#ifdef __cplusplus
}
#endif
#endif
