/*
 * qo_nodal_prog.h -- the device-side description of a nodal job (reference network
 * util/pa-bias-simulation/pa-bias-simulation.sch:19-72).  Shared, as text, between the ahead-of-time kernels of
 * qo_nodal.cu and the run-time compiled kernel (qo_nodal_jit.h hands this file to NVRTC), so it includes nothing:
 * the includer provides the fixed-width integer types and QO_NODAL_MAX_BR / QO_NODAL_MAX_PORTS.
 */
#ifndef QO_NODAL_PROG_H
#define QO_NODAL_PROG_H
#define QN_TPB 64
#define QN_MAX_SPEC 8
#define QN_MAX_VAR 64

struct NodalProg {
    int32_t n_nodes, nb, np, n_unk, nspec, hist_spec, hist_bins, n_var, dist, full;
    uint64_t seed;
    double hist_lo, hist_hi;
    int32_t port_node[QO_NODAL_MAX_PORTS];
    double port_z0[QO_NODAL_MAX_PORTS];
    int32_t kind[QO_NODAL_MAX_BR];
    int32_t node[QO_NODAL_MAX_BR][4];
    double nom[QO_NODAL_MAX_BR][4];
    double ttol[QO_NODAL_MAX_BR][4];
    int16_t tvar[QO_NODAL_MAX_BR][4];
    uint8_t tmode[QO_NODAL_MAX_BR][4];
    int32_t spec_min[QN_MAX_SPEC], spec_row[QN_MAX_SPEC], spec_col[QN_MAX_SPEC];   /* spec_min: 1 = "|S| >= limit" */
    double spec_thr[QN_MAX_SPEC];       /* linear |S|^2 threshold */
    double spec_limit_db[QN_MAX_SPEC];
};
#endif
