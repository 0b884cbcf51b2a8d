/*
 * qo_cpl_core.h -- QucsTranscalc "CoupledMicrostrip" analysis (physical -> electrical), SURVEY row N1, as one
 * routine shared by the host entry point qo_cpl_analyze (qo_cpl.c) and the device pre-pass that turns per-sample
 * geometry / substrate draws into per-sample (Z0e, Z0o, theta_e, theta_o) for the coupled-line block
 * (qo_cuda.cu::qo_cplms_kernel).  Replaces the GUI action that produced util/directional-couplers/dir_cpl_*.trc:18-20
 * (Z0e, Z0o, Ang_l from W, S, L, Er, H, H_t, T, Freq at :6-17).  Models: Hammerstad-Jensen single line,
 * Kirschning-Jansen 1984 even/odd statics and dispersion, Jansen thickness correction, March cover correction
 * (Appendix A.1/A.2 for the single line).
 *
 * Layout: the calculation is split the way the appendix is (single line -> mode widths -> static
 * permittivities -> static impedances -> dispersion), each stage filling one struct.
 */
#ifndef QO_CPL_CORE_H
#define QO_CPL_CORE_H
#include <math.h>
#include "qo_stream.h"      /* QO_HD */

#define CPL_PI 3.14159265358979323846
#define CPL_C0 299792458.0
#define CPL_ZF0 376.73031346958504364963

typedef struct {
    double u, g, tau, cover, fn, er;   /* W/h, S/h, T/h, H_t/h, f*h [GHz mm], substrate eps_r */
    double b_er;                       /* 0.564 ((er-0.9)/(er+3))^0.053 */
} cpl_geom;

typedef struct { double z, eps, zf, epsf, q0; } cpl_single;   /* zero-thickness single line, static + dispersive */

QO_HD double cpl_sq(double x) { return x * x; }

/* Hammerstad-Jensen Z0 of a line in air, u = W/h */
QO_HD double hj_z_air(double x)
{
    const double F = 6.0 + (2.0 * CPL_PI - 6.0) * exp(-pow(30.666 / x, 0.7528));
    return CPL_ZF0 / (2.0 * CPL_PI) * log(F / x + sqrt(1.0 + 4.0 / (x * x)));
}

/* exponent a(x) of the filling factor (1 + 10/x)^(-a b) */
QO_HD double hj_a(double x)
{
    const double x2 = x * x, x4 = x2 * x2;
    return 1.0 + log((x4 + x2 / 2704.0) / (x4 + 0.432)) / 49.0 + log(1.0 + x2 * x / 5929.741) / 18.7;
}

QO_HD double fill_inf(const cpl_geom *G, double x) { return pow(1.0 + 10.0 / x, -hj_a(x) * G->b_er); }

/* Kirschning-Jansen dispersion of the single line at width u: eps(f), Z(f) and the exponent R17 */
QO_HD void single_line(const cpl_geom *G, cpl_single *S)
{
    const double u = G->u, er = G->er, fn = G->fn;
    S->eps = 0.5 * (er + 1.0) + 0.5 * (er - 1.0) * fill_inf(G, u);
    S->z = hj_z_air(u) / sqrt(S->eps);
    /* permittivity */
    const double P1 = 0.27488 + (0.6315 + 0.525 / pow(1.0 + 0.0157 * fn, 20.0)) * u - 0.065683 * exp(-8.7513 * u);
    const double P2 = 0.33622 * (1.0 - exp(-0.03442 * er));
    const double P3 = 0.0363 * exp(-4.6 * u) * (1.0 - exp(-pow(fn / 38.7, 4.97)));
    const double P4 = 1.0 + 2.751 * (1.0 - exp(-pow(er / 15.916, 8.0)));
    const double Pf = P1 * P2 * pow((P3 * P4 + 0.1844) * fn, 1.5763);
    S->epsf = er - (er - S->eps) / (1.0 + Pf);
    /* impedance */
    const double R1 = 0.03891 * pow(er, 1.4), R2 = 0.267 * pow(u, 7.0), R3 = 4.766 * exp(-3.228 * pow(u, 0.641));
    const double R4 = 0.016 + pow(0.0514 * er, 4.524), R5 = pow(fn / 28.843, 12.0), R6 = 22.20 * pow(u, 1.92);
    const double R7 = 1.206 - 0.3144 * exp(-R1) * (1.0 - exp(-R2));
    const double R8 = 1.0 + 1.275 * (1.0 - exp(-0.004625 * R3 * pow(er, 1.674) * pow(fn / 18.365, 2.745)));
    const double em1_6 = pow(er - 1.0, 6.0);
    const double R9 = 5.086 * R4 * R5 / (0.3838 + 0.386 * R4) * exp(-R6) / (1.0 + 1.2992 * R5) * em1_6 / (1.0 + 10.0 * em1_6);
    const double R10 = 0.00044 * pow(er, 2.136) + 0.0184;
    const double f6 = pow(fn / 19.47, 6.0), R11 = f6 / (1.0 + 0.0962 * f6);
    const double R12 = 1.0 / (1.0 + 0.00245 * u * u);
    const double R13 = 0.9408 * pow(S->epsf, R8) - 0.9603, R14 = (0.9408 - R9) * pow(S->eps, R8) - 0.9603;
    const double R15 = 0.707 * R10 * pow(fn / 12.3, 1.097);
    const double R16 = 1.0 + 0.0503 * er * er * R11 * (1.0 - exp(-pow(u / 15.0, 6.0)));
    S->q0 = R7 * (1.0 - 1.1241 * R12 / R16 * exp(-0.026 * pow(fn, 1.15656) - R15));
    S->zf = S->z * pow(R13 / R14, S->q0);
}


/* inputs in SI units; out = { Z0e, Z0o, theta_e [deg], theta_o [deg] } */
QO_HD void qo_cpl_core(double w, double s, double h, double t, double er, double ht, double f, double len, double out[4])
{
    double z0e_v, z0o_v, ae_v, ao_v;
    double *z0e = &z0e_v, *z0o = &z0o_v, *ang_e_deg = &ae_v, *ang_o_deg = &ao_v;
    cpl_geom G = { w / h, s / h, t / h, ht / h, f * h / 1e6, er, 0.564 * pow((er - 0.9) / (er + 3.0), 0.053) };
    cpl_single S;
    single_line(&G, &S);
    const double u = G.u, g = G.g, fn = G.fn, em1 = er - 1.0;

    /* mode widths with strip thickness (Jansen) */
    double ue = u, uo = u;
    if (G.tau > 0.0) {
        const double step = 1.0 + exp(-100.0 * (u - 1.0 / (2.0 * CPL_PI)));
        const double du = 1.25 * G.tau / CPL_PI * (1.0 + log((2.0 + (4.0 * CPL_PI * u - 2.0) / step) / G.tau));
        const double dt = G.tau / (g * er);
        const double due = du * (1.0 - 0.5 * exp(-0.69 * du / dt));
        ue = u + due;
        uo = u + due + dt;
    }
    const double kq = 2.0 * log(2.0) / CPL_PI * G.tau;      /* thickness term of the filling factors */

    /* static permittivities */
    const double half_sum = 0.5 * (er + 1.0), half_dif = 0.5 * em1;
    const double v = ue * (20.0 + g * g) / (10.0 + g * g) + g * exp(-g);
    const double cov_e = G.cover <= 39.0 ? tanh(1.626 + 0.107 * G.cover - 1.733 / sqrt(G.cover)) : 1.0;
    const double eps_e0 = half_sum + half_dif * (fill_inf(&G, v) - kq / sqrt(ue)) * cov_e;
    const double b_o = 0.747 * er / (0.15 + er);
    const double c_o = b_o - (b_o - 0.207) * exp(-0.414 * uo);
    const double d_o = 0.593 + 0.694 * exp(-0.562 * uo);
    const double cov_o = G.cover <= 7.0 ? tanh(9.575 / (7.0 - G.cover) - 2.965 + 1.68 * G.cover - 0.311 * cpl_sq(G.cover)) : 1.0;
    const double q_o = (exp(-c_o * pow(g, d_o)) - kq / sqrt(uo)) * cov_o;
    const double a_o = 0.7287 * (S.eps - half_sum) * (1.0 - exp(-0.179 * uo));
    const double eps_o0 = (half_sum + a_o - S.eps) * q_o + S.eps;

    /* static impedances */
    const double rt_es = sqrt(S.eps);
    const double Q1 = 0.8695 * pow(ue, 0.194);
    const double Q2 = 1.0 + 0.7519 * g + 0.189 * pow(g, 2.31);
    const double g10 = pow(g, 10.0);
    const double Q3 = 0.1975 + pow(16.6 + pow(8.4 / g, 6.0), -0.387) + log(g10 / (1.0 + pow(g / 3.4, 10.0))) / 241.0;
    const double Q4 = 2.0 * Q1 / (Q2 * (exp(-g) * pow(ue, Q3) + (2.0 - exp(-g)) * pow(ue, -Q3)));
    const double ze0 = S.z * sqrt(S.eps / eps_e0) / (1.0 - rt_es * Q4 * S.z / CPL_ZF0);
    const double Q5 = 1.794 + 1.14 * log(1.0 + 0.638 / (g + 0.517 * pow(g, 2.43)));
    const double Q6 = 0.2305 + log(g10 / (1.0 + pow(g / 5.8, 10.0))) / 281.3 + log(1.0 + 0.598 * pow(g, 1.154)) / 5.1;
    const double Q7 = (10.0 + 190.0 * g * g) / (1.0 + 82.3 * g * g * g);
    const double Q8 = exp(-6.5 - 0.95 * log(g) - pow(g / 0.15, 5.0));
    const double Q9 = log(Q7) * (Q8 + 1.0 / 16.5);
    const double Q10 = (Q2 * Q4 - Q5 * exp(log(uo) * Q6 * pow(uo, -Q9))) / Q2;
    double zo0 = S.z * sqrt(S.eps / eps_o0) / (1.0 - rt_es * Q10 * S.z / CPL_ZF0);
    {   /* March's cover correction of the odd mode */
        const double c1 = 1.0 + G.cover;
        const double J = tanh(pow(c1, 1.585) / 6.0);
        const double K = g > 0.858 ? log10(20.492 * pow(g, 0.174)) : 1.30;
        const double Lc = g > 0.873 ? 2.51 * pow(g, -0.462) : 2.674;
        zo0 -= pow(uo, J) * 270.0 * (1.0 - tanh(2.178 - 0.796 * g + K * sqrt(c1) - Lc / c1)) / sqrt(eps_o0);
    }

    /* dispersion of the mode permittivities (evaluated with u, not ue/uo) */
    const double P1 = 0.27488 + (0.6315 + 0.525 / pow(1.0 + 0.0157 * fn, 20.0)) * u - 0.065683 * exp(-8.7513 * u);
    const double P2 = 0.33622 * (1.0 - exp(-0.03442 * er));
    const double P34 = 0.0363 * exp(-4.6 * u) * (1.0 - exp(-pow(fn / 38.7, 4.97))) * (1.0 + 2.751 * (1.0 - exp(-pow(er / 15.916, 8.0))));
    const double P6 = (0.334 * exp(-3.3 * pow(er / 15.0, 3.0)) + 0.746) * exp(-pow(fn / 18.0, 0.368));
    const double P7 = 1.0 + 4.069 * P6 * pow(g, 0.479) * exp(-1.347 * pow(g, 0.595) - 0.17 * pow(g, 2.5));
    const double F_e = P1 * P2 * pow((P34 + 0.1844 * P7) * fn, 1.5763);
    const double P8 = 0.7168 * (1.0 + 1.076 / (1.0 + 0.0576 * em1));
    const double P9 = P8 - 0.7913 * (1.0 - exp(-pow(fn / 20.0, 1.424))) * atan(2.481 * pow(er / 8.0, 0.946));
    const double P10 = 0.242 * pow(em1, 0.55);
    const double P11 = 0.6366 * (exp(-0.3401 * fn) - 1.0) * atan(1.263 * pow(u / 3.0, 1.629));
    const double P12 = P9 + (1.0 - P9) / (1.0 + 1.183 * pow(u, 1.376));
    const double P13 = 1.695 * P10 / (0.414 + 1.605 * P10);
    const double P14 = 0.8928 + 0.1072 * (1.0 - exp(-0.42 * pow(fn / 20.0, 3.215)));
    const double P15 = fabs(1.0 - 0.8928 * (1.0 + P11) * P12 * exp(-P13 * pow(g, 1.092)) / P14);
    const double F_o = P1 * P2 * pow((P34 + 0.1844) * fn * P15, 1.5763);
    const double eps_e = er - (er - eps_e0) / (1.0 + F_e);
    const double eps_o = er - (er - eps_o0) / (1.0 + F_o);

    /* dispersion of the even-mode impedance */
    const double f20 = pow(fn / 20.0, 4.91);
    const double Q11 = 0.893 * (1.0 - 0.3 / (1.0 + 0.7 * em1));
    const double Q12 = 2.121 * (f20 / (1.0 + Q11 * f20)) * exp(-2.87 * g) * pow(g, 0.902);
    const double Q13 = 1.0 + 0.038 * pow(er / 8.0, 5.1);
    const double e15 = pow(er / 15.0, 4.0);
    const double Q14 = 1.0 + 1.203 * e15 / (1.0 + e15);
    const double Q15 = 1.887 * exp(-1.5 * pow(g, 0.84)) * pow(g, Q14) /
                       (1.0 + 0.41 * pow(fn / 15.0, 3.0) * pow(u, 2.0 / Q13) / (0.125 + pow(u, 1.626 / Q13)));
    const double Q16 = (1.0 + 9.0 / (1.0 + 0.403 * em1 * em1)) * Q15;
    const double Q17 = 0.394 * (1.0 - exp(-1.47 * pow(u / 7.0, 0.672))) * (1.0 - exp(-4.25 * pow(fn / 20.0, 1.87)));
    const double Q18 = 0.61 * (1.0 - exp(-2.13 * pow(u / 8.0, 1.593))) / (1.0 + 6.544 * pow(g, 4.17));
    const double Q19 = 0.21 * cpl_sq(g * g) / ((1.0 + 0.18 * pow(g, 4.9)) * (1.0 + 0.1 * u * u) * (1.0 + pow(fn / 24.0, 3.0)));
    const double Q20 = (0.09 + 1.0 / (1.0 + 0.1 * pow(em1, 2.7))) * Q19;
    const double u25 = pow(u, 2.5);
    const double Q21 = fabs(1.0 - 42.54 * pow(g, 0.133) * exp(-0.812 * g) * u25 / (1.0 + 0.033 * u25));
    const double r_e = pow(fn / 28.843, 12.0);
    const double q_e = 0.016 + pow(0.0514 * er * Q21, 4.524);
    const double p_e = 4.766 * exp(-3.228 * pow(u, 0.641));
    const double em1_6 = pow(em1, 6.0);
    const double d_e = 5.086 * q_e * (r_e / (0.3838 + 0.386 * q_e)) * (exp(-22.2 * pow(u, 1.92)) / (1.0 + 1.2992 * r_e)) *
                       (em1_6 / (1.0 + 10.0 * em1_6));
    const double C_e = 1.0 + 1.275 * (1.0 - exp(-0.004625 * p_e * pow(er, 1.674) * pow(fn / 18.365, 2.745))) - Q12 + Q16 - Q17 + Q18 + Q20;
    *z0e = ze0 * pow((0.9408 * pow(S.epsf, C_e) - 0.9603) / ((0.9408 - d_e) * pow(S.eps, C_e) - 0.9603), S.q0);

    /* dispersion of the odd-mode impedance */
    const double em1_2 = em1 * em1, em1_3 = em1_2 * em1, em1_15 = pow(em1, 1.5);
    const double Q29 = 15.16 / (1.0 + 0.196 * em1_2);
    const double Q28 = 0.149 * em1_3 / (94.5 + 0.038 * em1_3);
    const double Q27 = 0.4 * pow(g, 0.84) * (1.0 + 2.5 * em1_15 / (5.0 + em1_15));
    const double x12 = pow(em1 / 13.0, 12.0);
    const double Q26 = 30.0 - 22.2 * (x12 / (1.0 + 3.0 * x12)) - Q29;
    const double Q25 = (0.3 * fn * fn / (10.0 + fn * fn)) * (1.0 + 2.333 * em1_2 / (5.0 + em1_2));
    const double u894 = pow(u, 0.894);
    const double Q24 = 2.506 * Q28 * u894 * pow((1.0 + 1.3 * u) * fn / 99.25, 4.29) / (3.575 + u894);
    const double Q23 = 1.0 + 0.005 * fn * Q27 / ((1.0 + 0.812 * pow(fn / 15.0, 1.9)) * (1.0 + 0.025 * u * u));
    const double Q22 = 0.925 * pow(fn / Q26, 1.536) / (1.0 + 0.3 * pow(fn / 30.0, 1.536));
    *z0o = S.zf + (zo0 * pow(eps_o / eps_o0, Q22) - S.zf * Q23) / (1.0 + Q24 + pow(0.46 * g, 2.2) * Q25);

    /* electrical lengths of the two modes; QucsTranscalc prints Ang_l = sqrt(theta_e * theta_o) */
    const double deg_per_rt_eps = 360.0 * len * f / CPL_C0;
    *ang_e_deg = deg_per_rt_eps * sqrt(eps_e);
    *ang_o_deg = deg_per_rt_eps * sqrt(eps_o);
    out[0] = z0e_v; out[1] = z0o_v; out[2] = ae_v; out[3] = ao_v;
}
#endif
