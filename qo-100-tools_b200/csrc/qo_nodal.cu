/*
 * qo_nodal.cu -- N-port nodal analysis on the GPU, sm_100a.
 *
 * Replaces qucsator's S-parameter analysis for networks that are not a cascade: the reference's bias networks
 * util/pa-bias-simulation/pa-bias-simulation.sch:19-72 (result: pa-bias-simulation.dat:1-85035).
 *
 * One thread owns one (sample, frequency) point: it stamps the modified-nodal matrix A (node voltages plus
 * one branch current per VCVS; port terminations 1/Z_k included), factorises it in place (LU, partial
 * pivoting) and back-substitutes once per port:  S[k][j] = 2 sqrt(Z_j/Z_k) V_k - delta_kj  with 1/Z_j
 * injected at port j.  A lives in thread-local memory (the hardware interleaves it across the warp, so a
 * warp's accesses to "its" element (i,j) are coalesced and L1-resident); its leading dimension is a template
 * parameter so that small networks do not pay for the largest.  Per sample the block first derives the
 * perturbed branch parameters (the shared bit-exact Philox stream) into shared memory.  Measured two-port
 * blocks carry no tolerances: their 2x2 admittance per grid point is tabulated on the host at plan time with
 * the same SPfile interpolation the cascade path uses.
 *
 * Work per point ~ (8/3) n^3 real FMAs (n = unknowns): FP64-bound, no reuse across points -- like the cascade
 * kernels this is latency/FP64-pipe bound, not memory bound.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "qo_ctx_internal.h"
#include "qo_stream.h"

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

__device__ __forceinline__ double2 c_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 c_inv(double2 b) { const double d = 1.0 / (b.x * b.x + b.y * b.y); return make_double2(b.x * d, -b.y * d); }
/* a -= l * u (complex fused multiply-subtract, 4 DFMA) */
__device__ __forceinline__ double2 c_fms(double2 a, double2 l, double2 u)
{
    a.x = fma(-l.x, u.x, a.x); a.x = fma(l.y, u.y, a.x);
    a.y = fma(-l.x, u.y, a.y); a.y = fma(-l.y, u.x, a.y);
    return a;
}

template <int LD> __device__ __forceinline__ void stamp(double2 *A, int a, int b, double2 y)
{
    if (a) A[(a - 1) * LD + (a - 1)] = c_add(A[(a - 1) * LD + (a - 1)], y);
    if (b) A[(b - 1) * LD + (b - 1)] = c_add(A[(b - 1) * LD + (b - 1)], y);
    if (a && b) {
        A[(a - 1) * LD + (b - 1)] = c_sub(A[(a - 1) * LD + (b - 1)], y);
        A[(b - 1) * LD + (a - 1)] = c_sub(A[(b - 1) * LD + (a - 1)], y);
    }
}

/*
 * unit u = sample * nchunks + chunk; a block works on one unit at a time: QN_TPB consecutive grid points per
 * pass.  Reduce-only jobs use nchunks == 1 (the block walks the whole grid of its sample and then reduces).
 */
template <int LD>
__global__ void __launch_bounds__(QN_TPB)
qo_nodal_kernel(const NodalProg *__restrict__ prog, const double *__restrict__ fgrid, const unsigned char *__restrict__ mask,
                const double2 *__restrict__ yblk, int nf, int chunk_len, int nchunks, unsigned long long sample_offset,
                unsigned long long nsamples, unsigned long long *__restrict__ counters, double2 *__restrict__ s_out)
{
    __shared__ double s_p[QO_NODAL_MAX_BR][4];
    __shared__ double s_x[QN_MAX_VAR];
    __shared__ double s_trk[QN_MAX_SPEC][QN_TPB / 32];
    __shared__ unsigned int s_cnt[2 + QN_MAX_SPEC + 1024];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = prog->n_unk, np = prog->np, nb = prog->nb, nspec = prog->nspec, n_nodes = prog->n_nodes;
    const int full = prog->full;
    const int ncnt = 2 + nspec + (prog->hist_bins > 0 ? prog->hist_bins : 0);
    for (int i = tid; i < ncnt; i += QN_TPB) s_cnt[i] = 0;
    __syncthreads();

    const unsigned long long n_units = nsamples * (unsigned long long)nchunks;
    for (unsigned long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        const unsigned long long s = u / (unsigned long long)nchunks;
        const int chunk = (int)(u - s * (unsigned long long)nchunks);
        /* per-sample branch parameters */
        for (int v = tid; v < prog->n_var; v += QN_TPB) s_x[v] = qo_stream_variate(prog->seed, sample_offset + s, (uint32_t)v, prog->dist);
        __syncthreads();
        for (int i = tid; i < nb * 4; i += QN_TPB) {
            const int b = i >> 2, k = i & 3;
            double v = prog->nom[b][k];
            const int tv = prog->tvar[b][k];
            if (tv >= 0) v = qo_stream_apply(v, prog->ttol[b][k], s_x[tv], prog->tmode[b][k]);
            s_p[b][k] = v;
        }
        __syncthreads();

        double trk[QN_MAX_SPEC];
#pragma unroll
        for (int sp = 0; sp < QN_MAX_SPEC; sp++) trk[sp] = -1e300;       /* running max of (+-)|S|^2 */
        const int k_lo = chunk * chunk_len, k_hi = min(nf, k_lo + chunk_len);
        for (int k = k_lo + tid; k < k_hi; k += QN_TPB) {
            double2 A[LD * LD];
            int perm[LD];
            const double w = 6.283185307179586476925286766559 * fgrid[k];
            for (int i = 0; i < n; i++) {
                perm[i] = i;
                for (int j = 0; j < n; j++) A[i * LD + j] = make_double2(0.0, 0.0);
            }
            int extra = n_nodes, iblk = 0;
            for (int b = 0; b < nb; b++) {
                const int kind = prog->kind[b];
                const int n0 = prog->node[b][0], n1 = prog->node[b][1], n2 = prog->node[b][2], n3 = prog->node[b][3];
                const double p0 = s_p[b][0], p1 = s_p[b][1], p2 = s_p[b][2];
                if (kind == QO_NB_R) stamp<LD>(A, n0, n1, make_double2(1.0 / p0, 0.0));
                else if (kind == QO_NB_L) {                /* Y = (1 - w^2 L Cp + j w R Cp) / (R + jwL) */
                    const double2 num = make_double2(1.0 - w * w * p0 * p2, w * p1 * p2);
                    stamp<LD>(A, n0, n1, c_mul(num, c_inv(make_double2(p1, w * p0))));
                } else if (kind == QO_NB_C) {              /* Y = 1 / (R + j(w Ls - 1/(wC))) */
                    stamp<LD>(A, n0, n1, c_inv(make_double2(p1, w * p2 - 1.0 / (w * p0))));
                } else if (kind == QO_NB_VCVS) {
                    const int kx = extra++;
                    double sn, cs;
                    sincos(w * p1, &sn, &cs);
                    const double2 g = make_double2(p0 * cs, -p0 * sn);
                    if (n1) { A[(n1 - 1) * LD + kx].x += 1.0; A[kx * LD + (n1 - 1)].x += 1.0; }
                    if (n2) { A[(n2 - 1) * LD + kx].x -= 1.0; A[kx * LD + (n2 - 1)].x -= 1.0; }
                    if (n0) A[kx * LD + (n0 - 1)] = c_sub(A[kx * LD + (n0 - 1)], g);
                    if (n3) A[kx * LD + (n3 - 1)] = c_add(A[kx * LD + (n3 - 1)], g);
                } else if (kind == QO_NB_SBLOCK) {
                    /* tabulated 2x2 admittance of the block at this grid point; indefinite form with reference n2 */
                    const double2 *y = yblk + ((size_t)iblk * (size_t)nf + (size_t)k) * 4;
                    iblk++;
                    const double2 y11 = y[0], y12 = y[1], y21 = y[2], y22 = y[3];
                    const int t[3] = { n0, n1, n2 };
                    double2 Y3[3][3];
                    Y3[0][0] = y11; Y3[0][1] = y12; Y3[1][0] = y21; Y3[1][1] = y22;
                    Y3[0][2] = make_double2(-(y11.x + y12.x), -(y11.y + y12.y));
                    Y3[1][2] = make_double2(-(y21.x + y22.x), -(y21.y + y22.y));
#pragma unroll
                    for (int c = 0; c < 3; c++) Y3[2][c] = make_double2(-(Y3[0][c].x + Y3[1][c].x), -(Y3[0][c].y + Y3[1][c].y));
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            if (t[r] && t[c]) A[(t[r] - 1) * LD + (t[c] - 1)] = c_add(A[(t[r] - 1) * LD + (t[c] - 1)], Y3[r][c]);
                }
            }
            for (int p = 0; p < np; p++) {
                const int pn = prog->port_node[p] - 1;
                A[pn * LD + pn].x += 1.0 / prog->port_z0[p];
            }
            /* LU, partial pivoting (rows are swapped physically: n <= 32).  Nodal matrices are sparse and every
             * point of a job shares one sparsity pattern, so the elimination walks bit masks of the non-zero
             * columns instead of full rows (warp-coherent: the masks are the same in all lanes unless a pivot
             * choice differs); lmask/umask record the factors' patterns for the substitutions. */
            bool singular = false;
            unsigned int lmask[LD], umask[LD];
            for (int c = 0; c < n; c++) {
                int piv = c;
                double best = A[c * LD + c].x * A[c * LD + c].x + A[c * LD + c].y * A[c * LD + c].y;
                for (int r = c + 1; r < n; r++) {
                    const double2 v = A[r * LD + c];
                    const double m = v.x * v.x + v.y * v.y;
                    if (m > best) { best = m; piv = r; }
                }
                if (!(best > 0.0)) { singular = true; break; }
                if (piv != c) {
                    for (int j = 0; j < n; j++) { const double2 t = A[c * LD + j]; A[c * LD + j] = A[piv * LD + j]; A[piv * LD + j] = t; }
                    const int t = perm[c]; perm[c] = perm[piv]; perm[piv] = t;
                }
                unsigned int nz = 0;
                for (int j = c + 1; j < n; j++) { const double2 v = A[c * LD + j]; if (v.x != 0.0 || v.y != 0.0) nz |= 1u << j; }
                umask[c] = nz;
                const double2 inv = c_inv(A[c * LD + c]);
                for (int r = c + 1; r < n; r++) {
                    const double2 a = A[r * LD + c];
                    if (a.x == 0.0 && a.y == 0.0) continue;
                    const double2 l = c_mul(a, inv);
                    A[r * LD + c] = l;
                    for (unsigned int m = nz; m; m &= m - 1) {
                        const int j = __ffs((int)m) - 1;
                        A[r * LD + j] = c_fms(A[r * LD + j], l, A[c * LD + j]);
                    }
                }
            }
            if (!singular)
                for (int i = 0; i < n; i++) {
                    unsigned int lm = 0;
                    for (int q = 0; q < i; q++) { const double2 v = A[i * LD + q]; if (v.x != 0.0 || v.y != 0.0) lm |= 1u << q; }
                    lmask[i] = lm;
                }
            const size_t obase = (((size_t)s * (size_t)nf + (size_t)k) * (size_t)np) * (size_t)np;
            const unsigned int mb = mask[k];
            for (int j = 0; j < np; j++) {
                double2 x[LD];
                const int src = prog->port_node[j] - 1;
                for (int i = 0; i < n; i++) x[i] = make_double2(perm[i] == src ? 1.0 / prog->port_z0[j] : 0.0, 0.0);
                if (singular) { for (int i = 0; i < n; i++) x[i] = make_double2(nan(""), nan("")); }
                else {
                    for (int i = 1; i < n; i++) {
                        double2 acc = x[i];
                        for (unsigned int m = lmask[i]; m; m &= m - 1) { const int q = __ffs((int)m) - 1; acc = c_fms(acc, A[i * LD + q], x[q]); }
                        x[i] = acc;
                    }
                    for (int i = n - 1; i >= 0; i--) {
                        double2 acc = x[i];
                        for (unsigned int m = umask[i]; m; m &= m - 1) { const int q = __ffs((int)m) - 1; acc = c_fms(acc, A[i * LD + q], x[q]); }
                        x[i] = c_mul(acc, c_inv(A[i * LD + i]));
                    }
                }
                for (int kk = 0; kk < np; kk++) {
                    const double sc = 2.0 * sqrt(prog->port_z0[j] / prog->port_z0[kk]);
                    const double2 v = x[prog->port_node[kk] - 1];
                    const double2 sv = make_double2(sc * v.x - (kk == j ? 1.0 : 0.0), sc * v.y);
                    if (full) s_out[obase + (size_t)kk * np + j] = sv;
#pragma unroll
                    for (int sp = 0; sp < QN_MAX_SPEC; sp++)
                        if (sp < nspec && ((mb >> sp) & 1u) && prog->spec_row[sp] == kk && prog->spec_col[sp] == j) {
                            const double m2 = sv.x * sv.x + sv.y * sv.y;
                            const double c = prog->spec_min[sp] ? -m2 : m2;
                            if (c > trk[sp] || m2 != m2) trk[sp] = m2 != m2 ? 1e300 : c;
                        }
                }
            }
        }
        if (!full) {
            /* per-sample verdict: block max of every tracker */
#pragma unroll
            for (int sp = 0; sp < QN_MAX_SPEC; sp++) {
                if (sp < nspec) {
                    double v = trk[sp];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, v, off); v = o > v ? o : v; }
                    if (lane == 0) s_trk[sp][warp] = v;
                }
            }
            __syncthreads();
            if (tid == 0) {
                unsigned int fail = 0;
                double hist_worst = 0.0;
                for (int sp = 0; sp < nspec; sp++) {
                    double v = s_trk[sp][0];
                    for (int q = 1; q < QN_TPB / 32; q++) v = s_trk[sp][q] > v ? s_trk[sp][q] : v;
                    if (v == -1e300) continue;                          /* the spec covers no grid point */
                    const double thr = prog->spec_min[sp] ? -prog->spec_thr[sp] : prog->spec_thr[sp];
                    if (v > thr) fail |= 1u << sp;
                    if (sp == prog->hist_spec) hist_worst = fabs(v);
                }
                atomicAdd(&s_cnt[0], fail == 0 ? 1u : 0u);
                atomicAdd(&s_cnt[1], 1u);
                for (int sp = 0; sp < nspec; sp++) if ((fail >> sp) & 1u) atomicAdd(&s_cnt[2 + sp], 1u);
                if (prog->hist_bins > 0) {
                    const double v = 10.0 * log10(hist_worst);
                    const double xb = (v - prog->hist_lo) / (prog->hist_hi - prog->hist_lo) * (double)prog->hist_bins;
                    long long b = (long long)floor(xb);
                    if (!(xb >= 0.0)) b = 0;
                    if (b >= prog->hist_bins) b = prog->hist_bins - 1;
                    atomicAdd(&s_cnt[2 + nspec + (int)b], 1u);
                }
            }
        }
        __syncthreads();
    }
    if (!full) {
        for (int i = tid; i < ncnt; i += QN_TPB)
            if (s_cnt[i]) atomicAdd(&counters[i], (unsigned long long)s_cnt[i]);
    }
}

/* ---- host side ------------------------------------------------------------------------------------------ */
static qo_c64 h_mul(qo_c64 a, qo_c64 b) { qo_c64 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static qo_c64 h_div(qo_c64 a, qo_c64 b)
{
    const double d = b.re * b.re + b.im * b.im;
    qo_c64 r = { (a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d };
    return r;
}

/* 2x2 admittance of a measured block: Y = (I - S)(I + S)^-1 / z0; y = { y11, y12, y21, y22 } */
static void s_to_y(const qo_c64 s[4], double z0, qo_c64 y[4])
{
    const qo_c64 s11 = s[0], s21 = s[1], s12 = s[2], s22 = s[3];
    const qo_c64 a = { 1.0 + s11.re, s11.im }, d = { 1.0 + s22.re, s22.im }, m11 = { 1.0 - s11.re, -s11.im }, m22 = { 1.0 - s22.re, -s22.im };
    const qo_c64 p = h_mul(s12, s21), ad = h_mul(a, d), det = { ad.re - p.re, ad.im - p.im };
    qo_c64 t, u;
    t = h_mul(m11, d); t.re += p.re; t.im += p.im; y[0] = h_div(t, det);
    t = h_mul(s12, a); u = h_mul(m11, s12); t.re = -t.re - u.re; t.im = -t.im - u.im; y[1] = h_div(t, det);
    t = h_mul(s21, d); u = h_mul(m22, s21); t.re = -t.re - u.re; t.im = -t.im - u.im; y[2] = h_div(t, det);
    t = h_mul(m22, a); t.re += p.re; t.im += p.im; y[3] = h_div(t, det);
    for (int i = 0; i < 4; i++) { y[i].re /= z0; y[i].im /= z0; }
}

static int nodal_run(qo_ctx *ctx, const qo_nodal *nd, const double *f, int nf, const qo_nspec *spec, int nspec,
                     const qo_mc_cfg *cfg, qo_mc_result *res, qo_c64 *full_s_host)
{
    if (!ctx || !nd || !f || nf <= 0 || !cfg || nspec < 0 || nspec > QN_MAX_SPEC || (nspec && !spec)) { qo_set_error("bad arguments"); return QO_ERR_ARG; }
    if (nd->np < 1) { qo_set_error("the netlist has no ports"); return QO_ERR_ARG; }
    for (int k = 0; k < nf; k++) if (!(f[k] > 0.0) || !isfinite(f[k])) { qo_set_error("frequency %d is not a positive finite number", k); return QO_ERR_ARG; }
    const int full = cfg->mode == QO_MODE_FULL_S;
    if (full && !full_s_host) { qo_set_error("FULL_S needs an output buffer"); return QO_ERR_ARG; }
    if (!full && !res) return QO_ERR_ARG;
    std::vector<NodalProg> hpv(1);
    NodalProg *hp = &hpv[0];
    memset(hp, 0, sizeof *hp);
    hp->n_nodes = nd->n_nodes; hp->nb = nd->nb; hp->np = nd->np; hp->full = full;
    hp->seed = cfg->seed; hp->dist = cfg->dist;
    if (hp->dist != QO_DIST_UNIFORM && hp->dist != QO_DIST_GAUSS3S) { qo_set_error("unknown distribution %d", hp->dist); return QO_ERR_ARG; }
    int n_unk = nd->n_nodes, n_sb = 0;
    for (int b = 0; b < nd->nb; b++) {
        hp->kind[b] = nd->br[b].kind;
        for (int k = 0; k < 4; k++) { hp->node[b][k] = nd->br[b].node[k]; hp->nom[b][k] = nd->br[b].p[k]; hp->tvar[b][k] = -1; }
        if (nd->br[b].kind == QO_NB_VCVS) n_unk++;
        if (nd->br[b].kind == QO_NB_SBLOCK) n_sb++;
    }
    if (n_unk > QO_NODAL_MAX_UNK) { qo_set_error("%d unknowns, limit %d", n_unk, QO_NODAL_MAX_UNK); return QO_ERR_RANGE; }
    hp->n_unk = n_unk;
    for (int p = 0; p < nd->np; p++) { hp->port_node[p] = nd->port_node[p]; hp->port_z0[p] = nd->port_z0[p]; }
    int nvar = 0;
    if (cfg->n_tol < 0 || (cfg->n_tol > 0 && !cfg->tol)) return QO_ERR_ARG;
    for (int i = 0; i < cfg->n_tol; i++) {
        const qo_tol *t = &cfg->tol[i];
        if (t->elem < 0 || t->elem >= nd->nb || t->param < 0 || t->param >= 4) { qo_set_error("tolerance %d: branch/param out of range", i); return QO_ERR_ARG; }
        if (t->var < 0 || t->var >= QN_MAX_VAR) { qo_set_error("tolerance %d: random variable index must be in [0,%d)", i, QN_MAX_VAR); return QO_ERR_RANGE; }
        if (nd->br[t->elem].kind == QO_NB_SBLOCK) { qo_set_error("tolerance %d: measured blocks carry no tolerances", i); return QO_ERR_ARG; }
        hp->tvar[t->elem][t->param] = (int16_t)t->var; hp->ttol[t->elem][t->param] = t->tol; hp->tmode[t->elem][t->param] = (uint8_t)(t->mode == QO_TOL_ABS);
        if (t->var + 1 > nvar) nvar = t->var + 1;
    }
    hp->n_var = nvar;
    hp->nspec = nspec;
    std::vector<unsigned char> mask((size_t)nf, 0);
    for (int s = 0; s < nspec; s++) {
        if (spec[s].kind != QO_SPEC_S21_MIN_DB && spec[s].kind != QO_SPEC_S21_MAX_DB) { qo_set_error("nodal spec %d: kind must be QO_SPEC_S21_MIN_DB or QO_SPEC_S21_MAX_DB", s); return QO_ERR_ARG; }
        if (spec[s].row < 0 || spec[s].row >= nd->np || spec[s].col < 0 || spec[s].col >= nd->np) { qo_set_error("nodal spec %d: S entry out of range", s); return QO_ERR_ARG; }
        hp->spec_min[s] = spec[s].kind == QO_SPEC_S21_MIN_DB; hp->spec_row[s] = spec[s].row; hp->spec_col[s] = spec[s].col;
        hp->spec_thr[s] = pow(10.0, spec[s].limit / 10.0); hp->spec_limit_db[s] = spec[s].limit;
        for (int k = 0; k < nf; k++) if (f[k] >= spec[s].f_lo && f[k] <= spec[s].f_hi) mask[k] |= (unsigned char)(1u << s);
    }
    hp->hist_bins = 0; hp->hist_spec = -1;
    if (!full && cfg->hist_bins > 0) {
        if (cfg->hist_bins > 1024 || cfg->hist_spec < 0 || cfg->hist_spec >= nspec || !(cfg->hist_hi > cfg->hist_lo)) { qo_set_error("bad histogram configuration"); return QO_ERR_ARG; }
        hp->hist_bins = cfg->hist_bins; hp->hist_spec = cfg->hist_spec; hp->hist_lo = cfg->hist_lo; hp->hist_hi = cfg->hist_hi;
    }
    /* measured blocks: admittance per (block branch, grid point), in branch order */
    std::vector<double2> yb((size_t)n_sb * nf * 4);
    int ib = 0;
    for (int b = 0; b < nd->nb; b++) {
        if (nd->br[b].kind != QO_NB_SBLOCK) continue;
        const qo_s2p *blk = nd->blk[(int)nd->br[b].p[0]];
        for (int k = 0; k < nf; k++) {
            qo_c64 sv[4], y[4];
            qo_s2p_eval(blk, f[k], nd->br[b].p[1] != 0.0, sv);
            s_to_y(sv, nd->br[b].p[2], y);
            for (int q = 0; q < 4; q++) yb[((size_t)ib * nf + k) * 4 + q] = make_double2(y[q].re, y[q].im);
        }
        ib++;
    }
    const int ncnt = 2 + nspec + hp->hist_bins;
    const unsigned long long N = cfg->n_samples;
    if (N == 0) { if (res) { res->n_pass = res->n_total = 0; } return QO_OK; }

    DevCtx *dc = &ctx->d[0];                 /* device 0 of the ctx: the nodal path is not sharded yet */
    CU(cudaSetDevice(dc->device));
    NodalProg *dprog = NULL;
    double *dfr = NULL;
    unsigned char *dmask = NULL;
    double2 *dy = NULL, *ds = NULL;
    unsigned long long *dcnt = NULL;
    int rc = QO_OK;
    const size_t s_elems = full ? (size_t)N * nf * nd->np * nd->np : 0;
#define CUN(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { qo_set_error("%s -> %s", #call, cudaGetErrorString(e_)); rc = e_ == cudaErrorMemoryAllocation ? QO_ERR_NOMEM : QO_ERR_CUDA; goto out; } } while (0)
    {
        CUN(cudaMalloc(&dprog, sizeof(NodalProg)));
        CUN(cudaMalloc(&dfr, (size_t)nf * sizeof(double)));
        CUN(cudaMalloc(&dmask, (size_t)nf));
        CUN(cudaMalloc(&dy, (yb.size() ? yb.size() : 1) * sizeof(double2)));
        CUN(cudaMalloc(&dcnt, (size_t)ncnt * sizeof(unsigned long long)));
        if (full) CUN(cudaMalloc(&ds, s_elems * sizeof(double2)));
        CUN(cudaMemcpyAsync(dprog, hp, sizeof(NodalProg), cudaMemcpyHostToDevice, dc->stream));
        CUN(cudaMemcpyAsync(dfr, f, (size_t)nf * sizeof(double), cudaMemcpyHostToDevice, dc->stream));
        CUN(cudaMemcpyAsync(dmask, mask.data(), (size_t)nf, cudaMemcpyHostToDevice, dc->stream));
        if (!yb.empty()) CUN(cudaMemcpyAsync(dy, yb.data(), yb.size() * sizeof(double2), cudaMemcpyHostToDevice, dc->stream));
        CUN(cudaMemsetAsync(dcnt, 0, (size_t)ncnt * sizeof(unsigned long long), dc->stream));
        /* FULL_S: cut the grid into QN_TPB-point chunks so that a nominal sweep fills the GPU */
        int chunk_len = nf, nchunks = 1;
        if (full) { chunk_len = QN_TPB; nchunks = (nf + chunk_len - 1) / chunk_len; }
        const unsigned long long units = N * (unsigned long long)nchunks;
        const unsigned long long cap = (unsigned long long)dc->sm_count * 16;
        const int grid = (int)(units < cap ? units : cap);
        cudaEventRecord(dc->ev0, dc->stream);
#define QN_LAUNCH(LDV) qo_nodal_kernel<LDV><<<grid, QN_TPB, 0, dc->stream>>>(dprog, dfr, dmask, dy, nf, chunk_len, nchunks, cfg->sample_offset, N, dcnt, ds)
        if (n_unk <= 8) QN_LAUNCH(8);
        else if (n_unk <= 16) QN_LAUNCH(16);
        else if (n_unk <= 24) QN_LAUNCH(24);
        else QN_LAUNCH(32);
#undef QN_LAUNCH
        cudaEventRecord(dc->ev1, dc->stream);
        CUN(cudaGetLastError());
        CUN(cudaStreamSynchronize(dc->stream));
        if (full) CUN(cudaMemcpy(full_s_host, ds, s_elems * sizeof(double2), cudaMemcpyDeviceToHost));
        if (res) {
            std::vector<unsigned long long> h((size_t)ncnt);
            CUN(cudaMemcpy(h.data(), dcnt, (size_t)ncnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            float ms = 0;
            cudaEventElapsedTime(&ms, dc->ev0, dc->ev1);
            res->n_pass = full ? 0 : h[0];
            res->n_total = full ? N : h[1];
            if (res->fail_per_spec) for (int s = 0; s < nspec; s++) res->fail_per_spec[s] = h[2 + s];
            if (res->hist) for (int b = 0; b < hp->hist_bins; b++) res->hist[b] = h[2 + nspec + b];
            res->seconds = ms * 1e-3;
            res->evals_per_s = ms > 0 ? (double)N * nf / (ms * 1e-3) : 0.0;
            res->flops_per_eval = (8.0 / 3.0) * n_unk * n_unk * n_unk + 8.0 * nd->np * n_unk * n_unk;   /* LU + substitutions, real flops */
        }
    }
out:
#undef CUN
    cudaFree(dprog); cudaFree(dfr); cudaFree(dmask); cudaFree(dy); cudaFree(dcnt); cudaFree(ds);
    return rc;
}

extern "C" int qo_nodal_mc_run(qo_ctx *ctx, const qo_nodal *nd, const double *f, int nf, const qo_nspec *spec, int nspec,
                               const qo_mc_cfg *cfg, qo_mc_result *res, qo_c64 *full_s)
{
    qo_clear_error();
    return nodal_run(ctx, nd, f, nf, spec, nspec, cfg, res, full_s);
}

extern "C" int qo_nodal_sweep(qo_ctx *ctx, const qo_nodal *nd, const double *f, int nf, qo_c64 *s)
{
    qo_clear_error();
    if (!s) return QO_ERR_ARG;
    qo_mc_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_samples = 1; cfg.mode = QO_MODE_FULL_S; cfg.precision = 64;
    return nodal_run(ctx, nd, f, nf, NULL, 0, &cfg, NULL, s);
}
