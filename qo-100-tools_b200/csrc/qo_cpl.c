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

/* QucsTranscalc "synthesize" for CoupledMicrostrip (electrical -> physical): the inverse of qo_cpl_analyze.  Z0e and Z0o fix
 * the strip width and gap (damped Newton on (ln W, ln S) with a finite-difference Jacobian of the analysis -- the
 * analysis is smooth and monotone in both), the electrical length Ang_l = sqrt(theta_e theta_o) then fixes L, which the
 * mode angles are proportional to.  Reference: the W / S / L lines of util/directional-couplers/dir_cpl_*.trc:15-17 are
 * what the tool produced from the Z0e / Z0o / Ang_l lines :18-20. */
int qo_cpl_synthesize(double z0e, double z0o, double ang_deg, double h, double t, double er, double ht, double f,
                      double *w, double *s, double *len)
{
    qo_clear_error();
    if (!w || !s || !len) return QO_ERR_ARG;
    if (!(z0e > z0o && z0o > 0 && ang_deg > 0 && h > 0 && t >= 0 && er > 1.0 && ht > 0 && f > 0)) {
        qo_set_error("qo_cpl_synthesize: need Z0e > Z0o > 0, ang, h, ht, f > 0, t >= 0, er > 1");
        return QO_ERR_ARG;
    }
    /* start: a 50-Ohm-ish line of width ~2h (er ~ 3-4) scaled by the target impedance, gap of one substrate height */
    double lw = log(h * 2.0 * 50.0 / sqrt(z0e * z0o)), ls = log(h);
    double o[4], r0, r1;
    const double L0 = 1e-3;
    int it;
    for (it = 0; it < 100; it++) {
        qo_cpl_core(exp(lw), exp(ls), h, t, er, ht, f, L0, o);
        if (!isfinite(o[0]) || !isfinite(o[1])) break;
        r0 = o[0] / z0e - 1.0; r1 = o[1] / z0o - 1.0;
        if (fabs(r0) < 1e-13 && fabs(r1) < 1e-13) break;
        const double dl = 1e-6;
        double a[4], b[4];
        qo_cpl_core(exp(lw + dl), exp(ls), h, t, er, ht, f, L0, a);
        qo_cpl_core(exp(lw), exp(ls + dl), h, t, er, ht, f, L0, b);
        const double j00 = (a[0] - o[0]) / (z0e * dl), j01 = (b[0] - o[0]) / (z0e * dl);
        const double j10 = (a[1] - o[1]) / (z0o * dl), j11 = (b[1] - o[1]) / (z0o * dl);
        const double det = j00 * j11 - j01 * j10;
        if (!(fabs(det) > 1e-300)) break;
        double dw = -(r0 * j11 - r1 * j01) / det, ds = -(j00 * r1 - j10 * r0) / det;
        /* damping: at most a factor e per step in either dimension */
        const double m = fmax(fabs(dw), fabs(ds));
        if (m > 1.0) { dw /= m; ds /= m; }
        lw += dw; ls += ds;
    }
    qo_cpl_core(exp(lw), exp(ls), h, t, er, ht, f, L0, o);
    if (!isfinite(o[0]) || !isfinite(o[1]) || fabs(o[0] / z0e - 1.0) > 1e-9 || fabs(o[1] / z0o - 1.0) > 1e-9) {
        qo_set_error("qo_cpl_synthesize: no coupled-microstrip geometry gives Z0e = %g, Z0o = %g on this substrate", z0e, z0o);
        return QO_ERR_RANGE;
    }
    /* the Kirschning-Jansen fits are stated for 0.1 <= W/h, S/h <= 10; a solution far outside is an extrapolation artefact */
    if (exp(lw) / h < 0.05 || exp(lw) / h > 20.0 || exp(ls) / h < 0.05 || exp(ls) / h > 20.0) {
        qo_set_error("qo_cpl_synthesize: W/h = %g, S/h = %g lies outside the coupled-microstrip model's range", exp(lw) / h, exp(ls) / h);
        return QO_ERR_RANGE;
    }
    *w = exp(lw); *s = exp(ls);
    *len = L0 * ang_deg / sqrt(o[2] * o[3]);
    return QO_OK;
}
