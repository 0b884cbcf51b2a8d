/*
 * qo100ref.h -- CPU ORACLE for the qo-100-tools network-simulation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the reported CPU baseline.  The product (libqo100net.so) never links,
 * imports or calls it and has no CPU fallback.
 *
 * What it restates: the reference tree (vankxr/qo-100-tools) contains NO
 * evaluation code for this path -- util/ holds inputs/outputs of external GUI
 * tools (Qucs 0.0.19 qucsator, QucsTranscalc 0.0.19/0.0.20, rf-tools.com;
 * none vendored, none pinned by a lockfile, none installed here).  This file
 * therefore restates the *published models those tools apply* and is pinned
 * against the reference's own artefacts:
 *   - util/pa-lpf-simulation/pa-lpf-simulation.dat:5-35017 (5000-point full
 *     complex S-matrix, 21 digits)              -> PINNED (<= 1e-9 relative)
 *   - util/pa-lpf-simulation/pa-lpf-simulation.dpl:25-28 (4 markers)
 *   - util/directional-couplers/ *.trc:18-20 (Z0e/Z0o/Ang_l, 6 digits)
 *   - rf-tools ladders (SVG element lists; PNG plots only) -> PARITY UNPINNED
 *     numerically; pinned only to 40-digit mpmath values of the textbook
 *     ladder equations that reproduce the PNG edge values.
 *   - Monte-Carlo / yield / ESR-SRF parasitics / group delay do not exist
 *     upstream -> PARITY UNPINNED; this oracle *defines* them.
 *
 * Plain C11, double precision, compiled with -O2 -ffp-contract=off.
 */
#ifndef QO100REF_H
#define QO100REF_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element kinds: numeric values match include/qo100net.h (the tests pass the
 * product loader's element lists straight into the oracle) */
enum {
    REF_SER_R = 1, REF_SHUNT_R = 2,
    REF_SER_L = 3, REF_SHUNT_L = 4,        /* p0=L  p1=ESR p2=Cp  */
    REF_SER_C = 5, REF_SHUNT_C = 6,        /* p0=C  p1=ESR p2=ESL */
    REF_SER_LC_SER = 7,                    /* Z = jwL + 1/(jwC)      p0=L p1=C */
    REF_SER_LC_PAR = 8,                    /* Z = 1/(jwC + 1/(jwL))            */
    REF_SHUNT_LC_SER = 9,                  /* Y = 1/(jwL + 1/(jwC))            */
    REF_SHUNT_LC_PAR = 10,                 /* Y = jwC + 1/(jwL)                */
    REF_TLINE = 11,                        /* p0=Z0 p1=ang_deg p2=f0 (lossless) */
    REF_CPL_THRU = 12,                     /* p0=Z0e p1=Z0o p2=ang_e p3=ang_o p4=f0 p5=Zt */
    REF_SUBST = 13,                        /* p = er,h,t,tand,rho,D */
    REF_MLIN = 14,                         /* p0=W p1=L */
    REF_MCORN = 15,                        /* p0=W */
    REF_MTEE = 16,                         /* p0=Wa p1=Wb p2=W2 ; opens the side arm */
    REF_MOPEN = 17,                        /* p0=W ; closes the side arm */
    REF_SBLOCK = 18,                       /* measured two-port: p0 = registered block index, p1 = 1 polar / 0 rectangular */
    REF_CPL_MS = 19                        /* physical coupled microstrip through path on the preceding REF_SUBST:
                                              p0=W p1=S p2=L p3=H_t p4=f0 p5=Zt (util/directional-couplers/ *.trc:6-20) */
};

typedef struct { int32_t kind; int32_t flags; double p[6]; } ref_elem;

enum { REF_SPEC_S21_MIN_DB = 1, REF_SPEC_S21_MAX_DB = 2, REF_SPEC_S11_MAX_DB = 3, REF_SPEC_GD_MAX = 4 };
typedef struct { int32_t kind; int32_t pad; double f_lo, f_hi, limit; } ref_spec;

enum { REF_DIST_UNIFORM = 0, REF_DIST_GAUSS3S = 1 };
enum { REF_TOL_REL = 0, REF_TOL_ABS = 1 };
typedef struct { int32_t elem, param, var, mode; double tol; } ref_tol;

typedef struct {
    uint64_t seed, sample_offset, n_samples;
    int32_t dist, n_tol;
    const ref_tol *tol;
    int32_t hist_bins, hist_spec;
    double hist_lo, hist_hi;
} ref_mc_cfg;

/* Philox4x32-10, SURVEY App. C */
void   ref_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double ref_uniform01(uint64_t seed, uint64_t sample, uint32_t var);
double ref_variate(uint64_t seed, uint64_t sample, uint32_t var, int dist);  /* in [-1,1] */
double ref_perturb_factor(uint64_t seed, uint64_t sample, uint32_t var, int dist, double tol);
void ref_perturb_factors(uint64_t seed, uint64_t sample_offset, uint64_t n_samples, int n_var, int dist, double tol, double *out);
double ref_norminv(double p);
double ref_log_det(double x);

/* grids */
void ref_grid_lin(double f0, double f1, int n, double *f);
void ref_grid_log(double f0, double f1, int n, double *f);

/* synthesis, SURVEY App. B.3 / B.5 */
int ref_cheby_g(int n, double ripple_db, double *g);      /* g[0..n-1] */
int ref_butter_g(int n, double *g);
int ref_ladder_lpf(const double *g, int n, double fc, double z0, int series_first, ref_elem *out);
void ref_add_parasitics(ref_elem *e, int n, double fc, double q_l, double srf_l_mult, double esr_c, double srf_c_mult);

/* measured two-port blocks (Touchstone data handed over by the tests); s = n x 8 doubles, (re,im) of S11,S21,S12,S22 */
int ref_sblock_register(int idx, const double *f, const double *s, int n, double z0);
void ref_sblock_clear(void);

/* ---- N-port nodal analysis (qo100ref_nodal.c; pinned by util/pa-bias-simulation/pa-bias-simulation.dat) ---- */
enum { REF_NB_R = 1, REF_NB_L = 2, REF_NB_C = 3, REF_NB_VCVS = 4, REF_NB_SBLOCK = 5 };
/* R: nodes a,b p0=R | L: a,b p0=L p1=ESR p2=Cp | C: a,b p0=C p1=ESR p2=ESL | VCVS: in+,out+,out-,in- p0=gain p1=delay
 * SBLOCK: t1,t2,ref p0=registered block index p1=polar p2=z0.  Node 0 = ground. */
typedef struct { int32_t kind; int32_t node[4]; double p[4]; } ref_branch;
typedef struct { int32_t kind; int32_t row, col, pad; double f_lo, f_hi, limit; } ref_nspec;  /* kind: REF_SPEC_S21_MIN_DB / _MAX_DB on |S[row][col]| */
int ref_nodal_sweep(const ref_branch *br, int nb, int n_nodes, const int *port_node, const double *port_z0, int np,
                    const double *f, int nf, double *s_out /* [nf][np][np] (re,im), S[k][j] = b_k/a_j */);
int ref_nodal_mc_run(const ref_branch *br, int nb, int n_nodes, const int *port_node, const double *port_z0, int np,
                     const double *f, int nf, const ref_nspec *spec, int nspec, const ref_mc_cfg *cfg,
                     uint64_t *counters, double *full_s, int nthreads);
int ref_sblock_s(int idx, double f, int polar, double s[8]);

/* nominal sweep; s?? are interleaved (re,im) arrays of 2*nf doubles, nullable; gd nullable */
int ref_sweep(const ref_elem *e, int n, double rs, double rl, const double *f, int nf,
              double *s11, double *s21, double *s12, double *s22, double *gd);

/* apply sample `sample`'s perturbations to a copy of the element list */
void ref_perturb(const ref_elem *e, int n, const ref_mc_cfg *cfg, uint64_t sample, ref_elem *out);

/* Monte-Carlo yield.  counters: [0]=n_pass [1]=n_total [2..2+nspec)=fail_per_spec
 * then hist[hist_bins].  full_s (nullable): planes [4][n_samples][nf] of (re,im),
 * plane order S11,S21,S12,S22.  nthreads<=1 -> scalar loop; >1 -> OpenMP over samples. */
int ref_mc_run(const ref_elem *e, int n, double rs, double rl, const double *f, int nf,
               const ref_spec *spec, int nspec, const ref_mc_cfg *cfg,
               uint64_t *counters, double *full_s, int nthreads);
int ref_max_threads(void);

/* microstrip sub-models exposed for unit checks */
void ref_ms_quasi(double W, double h, double t, double er, double *Z, double *E, double *Weff);
void ref_ms_disp(double W, double h, double er, double Z, double E, double f, double *Zf, double *Ef);

/* QucsTranscalc CoupledMicrostrip analysis; SI units */
void ref_cpl_analyze(double w, double s, double h, double t, double er, double ht, double f, double len,
                     double *z0e, double *z0o, double *ang_e_deg, double *ang_o_deg);

#ifdef __cplusplus
}
#endif
#endif
