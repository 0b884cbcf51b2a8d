/*
 * qo_tf_core.h -- element -> rational immittance, shared by the transfer-function kernel (qo_tf.cuh, device) and
 * its plan-time self-check (qo_tf.cu, host).
 *
 * Every lumped branch of the util/ networks (reference: util/if-bandpass-filter/schematic.svg:191-213 series-LC /
 * parallel-LC branches, util/gpsdo-ouput-filters/10M/schematic.svg:197-231, docs/gpsdo-filters/<name>.svg:195-241 traps,
 * pcb/generic-filter/qo-100-generic-filter.sch:1450-1488 ladder positions with the ESR/SRF parasitic model of
 * SURVEY 8d) has an immittance that is a ratio of real polynomials of degree <= 2 in s = jw:
 *     series branch  Z(s) = N(s) / D(s)          shunt branch  Y(s) = N(s) / D(s)
 * With sn = s / wref (wref = geometric centre of the grid) the coefficient of sn^k carries wref^k.
 */
#pragma once
#include "qo_device.cuh"

#if defined(__CUDACC__)
#define QO_TF_HD __host__ __device__ __forceinline__
#else
#define QO_TF_HD static inline
#endif

/* p[] = the element's (perturbed) parameters in the order of include/qo100net.h; nd[0..2] = N, nd[3..5] = D.
 * returns 1 for a series branch, 0 for a shunt branch, -1 when the opcode has no lumped rational form */
QO_TF_HD int qo_tf_element(int opcode, const double *p, double wr, double *nd)
{
    const double wr2 = wr * wr;
    switch (opcode) {
    case OP_SER_R: nd[0] = p[0]; nd[1] = 0; nd[2] = 0; nd[3] = 1; nd[4] = 0; nd[5] = 0; return 1;
    case OP_SHUNT_G: nd[0] = 1; nd[1] = 0; nd[2] = 0; nd[3] = p[0]; nd[4] = 0; nd[5] = 0; return 0;
    /* inductor (L, R, Cp): Z = (R + sL) / (1 + s R Cp + s^2 L Cp) */
    case OP_SER_L: case OP_SER_LOSSY_L:
        nd[0] = p[1]; nd[1] = p[0] * wr; nd[2] = 0; nd[3] = 1; nd[4] = p[1] * p[2] * wr; nd[5] = p[0] * p[2] * wr2; return 1;
    case OP_SHUNT_L: case OP_SHUNT_LOSSY_L:
        nd[3] = p[1]; nd[4] = p[0] * wr; nd[5] = 0; nd[0] = 1; nd[1] = p[1] * p[2] * wr; nd[2] = p[0] * p[2] * wr2; return 0;
    /* capacitor (C, R, Ls): Z = (1 + s R C + s^2 Ls C) / (s C) */
    case OP_SER_C: case OP_SER_LOSSY_C:
        nd[0] = 1; nd[1] = p[1] * p[0] * wr; nd[2] = p[2] * p[0] * wr2; nd[3] = 0; nd[4] = p[0] * wr; nd[5] = 0; return 1;
    case OP_SHUNT_C: case OP_SHUNT_LOSSY_C:
        nd[3] = 1; nd[4] = p[1] * p[0] * wr; nd[5] = p[2] * p[0] * wr2; nd[0] = 0; nd[1] = p[0] * wr; nd[2] = 0; return 0;
    /* ideal LC pairs (L, C) */
    case OP_SER_LCS:   /* Z = sL + 1/(sC) */
        nd[0] = 1; nd[1] = 0; nd[2] = p[0] * p[1] * wr2; nd[3] = 0; nd[4] = p[1] * wr; nd[5] = 0; return 1;
    case OP_SER_LCP:   /* Z = sL / (1 + s^2 L C) */
        nd[0] = 0; nd[1] = p[0] * wr; nd[2] = 0; nd[3] = 1; nd[4] = 0; nd[5] = p[0] * p[1] * wr2; return 1;
    case OP_SHUNT_LCS: /* Y = sC / (1 + s^2 L C) */
        nd[0] = 0; nd[1] = p[1] * wr; nd[2] = 0; nd[3] = 1; nd[4] = 0; nd[5] = p[0] * p[1] * wr2; return 0;
    case OP_SHUNT_LCP: /* Y = sC + 1/(sL) */
        nd[0] = 1; nd[1] = 0; nd[2] = p[0] * p[1] * wr2; nd[3] = 0; nd[4] = p[0] * wr; nd[5] = 0; return 0;
    default: return -1;
    }
}

/* structural degree an element adds to the polynomials (max(deg N, deg D)); 0 = resistor, -1 = not lumped */
static inline int qo_tf_degree(int opcode)
{
    switch (opcode) {
    case OP_SER_R: case OP_SHUNT_G: return 0;
    case OP_SER_L: case OP_SHUNT_L: case OP_SER_C: case OP_SHUNT_C: return 1;
    case OP_SER_LOSSY_L: case OP_SHUNT_LOSSY_L: case OP_SER_LOSSY_C: case OP_SHUNT_LOSSY_C:
    case OP_SER_LCS: case OP_SER_LCP: case OP_SHUNT_LCS: case OP_SHUNT_LCP: return 2;
    default: return -1;
    }
}
