/*
 * qo_cpl.c -- host entry point of the QucsTranscalc "CoupledMicrostrip" analysis; the model itself
 * lives in qo_cpl_core.h, shared with the device pre-pass of the physical coupled-line element (QO_CPL_MS).
 */
#include <math.h>
#include "qo_internal.h"
#include "qo_cpl_core.h"

int qo_cpl_analyze(double w, double s, double h, double t, double er, double ht, double f, double len,
                   double *z0e, double *z0o, double *ang_e_deg, double *ang_o_deg)
{
    qo_clear_error();
    if (!z0e || !z0o || !ang_e_deg || !ang_o_deg) return QO_ERR_ARG;
    if (!(w > 0 && s > 0 && h > 0 && t >= 0 && er > 1.0 && ht > 0 && f > 0 && len > 0)) {
        qo_set_error("qo_cpl_analyze: need w, s, h, ht, f, len > 0, t >= 0, er > 1");
        return QO_ERR_ARG;
    }
    double o[4];
    qo_cpl_core(w, s, h, t, er, ht, f, len, o);
    *z0e = o[0]; *z0o = o[1]; *ang_e_deg = o[2]; *ang_o_deg = o[3];
    if (!isfinite(*z0e) || !isfinite(*z0o)) { qo_set_error("qo_cpl_analyze: geometry outside the model's range"); return QO_ERR_RANGE; }
    return QO_OK;
}
