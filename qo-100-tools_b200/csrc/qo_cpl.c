/* qo_cpl.c -- placeholder translation unit for the coupled-microstrip analysis (row N1) */
#include "qo_internal.h"
int qo_cpl_analyze(double w, double s, double h, double t, double er, double ht, double f, double len,
                   double *z0e, double *z0o, double *ang_e_deg, double *ang_o_deg)
{
    (void)w; (void)s; (void)h; (void)t; (void)er; (void)ht; (void)f; (void)len;
    (void)z0e; (void)z0o; (void)ang_e_deg; (void)ang_o_deg;
    qo_set_error("qo_cpl_analyze: not implemented yet");
    return QO_ERR_UNSUPPORTED;
}
