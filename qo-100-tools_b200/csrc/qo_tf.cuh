/*
 * qo_tf.cuh -- transfer-function Monte-Carlo kernel for lumped ladders (sm_100a, FP64).
 *
 * Same job as the straight-line ladder kernel of qo_ladder.cuh (reduce-only |S21| yield of the pcb/generic-filter
 * ladder family, reference pcb/generic-filter/README.md:13, qo-100-generic-filter.sch:1450-1488,1703-1995, and of the
 * rf-tools ladders util/if-bandpass-filter/schematic.svg:191-213, docs/gpsdo-filters/*.svg:195-241), restructured so
 * that the per-POINT work no longer grows with a complex chain step per element:
 *
 *   every lumped branch is a ratio of real polynomials in s (qo_tf_core.h), so the cascade
 *       [P(s); Q(s)] / D(s) = M1(s) M2(s) ... MN(s) [Rl; 1]
 *   has REAL polynomial entries.  Per SAMPLE the warp expands them once -- lane i holds the coefficient of s^i,
 *   one element is a 3-tap convolution done with two warp shuffles per polynomial -- and per POINT each thread
 *   evaluates them at s = j x by Horner in y = -x^2 on the even / odd coefficients:
 *       den = (P + Rs Q) / D          |S21|^2 = 4 Rs Rl / |den|^2
 *   i.e. 4 real Horner chains of K = deg/2 + 1 coefficients (|Num|, |D|) -- 44 DFMA for the 11-element ladder with
 *   ESR/SRF parasitics (degree 22) instead of its ~145 chain instructions, and all of them FMAs.
 *   With the coupled-line block in front (BASELINE config 5) P and Q stay separate (6 chains) and are contracted
 *   with the block's row vector [1 Rs] k M_cpl of qo_ladder.cuh::lad_cpl_first.
 *
 * Accuracy: the monomial basis loses log10(kappa) digits, kappa = sum |c_k| x^k / |Num(jx)| (5e3 on the 0.1 dB
 * Chebyshev 11th-order pass-band edge: 4e-13 relative on |den|^2, measured against a 60-digit evaluation).  The
 * plan builder (qo_tf.cu) evaluates the nominal network and its tolerance-box corners both ways on every in-band
 * grid point and only selects this kernel when they agree to 1e-10 (north_star asks 1e-9); otherwise the job
 * runs on the chain kernels.  QO100NET_KERNEL=ladder|interp selects those explicitly.
 */
#pragma once
#include "qo_ladder.cuh"
#include "qo_tf_core.h"

#define QO_TF_MAXK 15            /* coefficients per Horner chain: degree <= 29 (lanes 30, 31 stay zero: free shuffle wrap-around) */
#define QO_TF_MAXEL 24           /* lumped elements */
#define QO_TF_REC 8              /* doubles per element record: N0 N1 N2 D0 D1 D2 series - */

enum { QO_TF_S21 = 0 /* Num, D */, QO_TF_S21_NOD = 1 /* D == 1: ideal L/C/R ladders */, QO_TF_CPL = 2 /* P, Q, D + coupler block */ };

struct TfParams {
    const DevProg *prog;
    const double2 *xt;                       /* normalised w / wref per grid point, two points per entry */
    const double2 *wt;                       /* w (coupler block) */
    const uchar2 *m2;
    const double2 *cse, *cce, *cso, *cco;    /* coupler: sin/cos of the nominal mode angles (cpl_fast) */
    unsigned long long *counters, *ticket;
    unsigned long long sample_offset, nsamples, seed;
    double rs, rl, k21, hist_lo, hist_hi, wref, zn, zni;
    double thr[QO_LAD_NSPEC];
    int neg[QO_LAD_NSPEC];
    int npairs, n_var, n_el, el0, nspec, dist, hist_spec, hist_bins, hist_kind;
    int cpl_fast, cpl_same, cpl_op;
    const double *cplms;
};

template <int MODE> struct TfChains { static constexpr int n = MODE == QO_TF_S21 ? 4 : MODE == QO_TF_S21_NOD ? 2 : 6; };

/* per-sample element record: perturbed parameters -> N, D normalised to zn = sqrt(Rs Rl) */
__device__ __forceinline__ void tf_derive(const DevProg *__restrict__ prog, int e, const double *__restrict__ x, double wr, double zn, double zni,
                                          double *rec)
{
    double p[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        p[k] = prog->nom[e][k];
        const int tv = prog->tvar[e][k];
        if (tv >= 0) p[k] = qo_stream_apply(p[k], prog->ttol[e][k], x[tv], prog->tmode[e][k]);
    }
    double nd[6];
    const int series = qo_tf_element(prog->opcode[e], p, wr, nd);
    const double sc = series ? zni : zn;
    rec[0] = nd[0] * sc; rec[1] = nd[1] * sc; rec[2] = nd[2] * sc; rec[3] = nd[3]; rec[4] = nd[4]; rec[5] = nd[5];
    rec[6] = series ? 1.0 : 0.0; rec[7] = 0.0;
}

__device__ __forceinline__ double tf_up(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <int K, int MODE, int PP, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) qo_mc_tf_kernel(const __grid_constant__ TfParams P)
{
    constexpr int PTS = 2 * PP;
    constexpr int WARPS = TPB / 32;
    constexpr int NCH = TfChains<MODE>::n;
    constexpr bool CPL = MODE == QO_TF_CPL;
    __shared__ __align__(16) double s_poly[WARPS][K * NCH];
    __shared__ __align__(16) double s_el[WARPS][QO_TF_MAXEL * QO_TF_REC];
    __shared__ __align__(16) double s_cpl[WARPS][CPL ? QO_LAD_CPL + 2 : 2];
    __shared__ double s_x[WARPS][QO_MAX_VAR];
    __shared__ unsigned int s_cnt[2 + QO_NSPEC_MAX + QO_MAX_HIST];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncnt = 2 + P.nspec + (P.hist_bins > 0 ? P.hist_bins : 0);
    for (int i = threadIdx.x; i < ncnt; i += TPB) s_cnt[i] = 0;
    __syncthreads();

    double *polyw = s_poly[warp], *elw = s_el[warp], *xw = s_x[warp];
    const unsigned int polys = (unsigned int)__cvta_generic_to_shared(polyw);
    const unsigned int cpls = (unsigned int)__cvta_generic_to_shared(s_cpl[warp]);
    const int npairs = P.npairs;
    const double rs = P.rs;
    const int up1 = (lane + 31) & 31, up2 = (lane + 30) & 31;

    const unsigned long long total_warps = (unsigned long long)gridDim.x * WARPS;
    unsigned long long s = (unsigned long long)blockIdx.x * WARPS + warp;
    while (s < P.nsamples) {
        unsigned long long s_next = 0;
        if (lane == 0) s_next = total_warps + atomicAdd(P.ticket, 1ull);
        /* 1. the sample's random variables and element records */
        for (int v = lane; v < P.n_var; v += 32) xw[v] = qo_stream_variate(P.seed, P.sample_offset + s, (uint32_t)v, P.dist);
        __syncwarp();
        if (lane < P.n_el) tf_derive(P.prog, P.el0 + lane, xw, P.wref, P.zn, P.zni, elw + lane * QO_TF_REC);
        if (CPL && lane == 31) {
            double nom_k[2];
            lad_derive<double>(P.prog, P.cpl_op, xw, s_cpl[warp], P.cplms ? P.cplms + 4 * s : NULL, nom_k);
        }
        __syncwarp();
        /* 2. expand [P; Q] and D from the load end: lane i holds the coefficient of sn^i */
        double p = lane == 0 ? P.rl : 0.0, q = lane == 0 ? P.zn : 0.0, d = lane == 0 ? 1.0 : 0.0;
        for (int e = P.n_el - 1; e >= 0; e--) {
            const double2 n01 = *(const double2 *)(elw + e * QO_TF_REC), n2d0 = *(const double2 *)(elw + e * QO_TF_REC + 2),
                          d12 = *(const double2 *)(elw + e * QO_TF_REC + 4);
            const bool series = elw[e * QO_TF_REC + 6] != 0.0;
            /* series Z = N/D: P <- D P + N Q, Q <- D Q;  shunt Y = N/D: Q <- D Q + N P, P <- D P */
            double a = series ? p : q, b = series ? q : p;
            const double a1 = tf_up(a, up1), a2 = tf_up(a, up2), b1 = tf_up(b, up1), b2 = tf_up(b, up2);
            double na = fma(n2d0.y, a, fma(d12.x, a1, d12.y * a2));
            na = fma(n01.x, b, fma(n01.y, b1, fma(n2d0.x, b2, na)));
            const double nb = fma(n2d0.y, b, fma(d12.x, b1, d12.y * b2));
            p = series ? na : nb; q = series ? nb : na;
            if (MODE != QO_TF_S21_NOD) {
                const double d1 = tf_up(d, up1), d2 = tf_up(d, up2);
                d = fma(n2d0.y, d, fma(d12.x, d1, d12.y * d2));
            }
        }
        /* chain table: step k holds the coefficients of sn^(2k) and sn^(2k+1) of every polynomial */
        if (lane < 2 * K) {
            const int k = lane >> 1, par = lane & 1;
            if (CPL) { polyw[k * NCH + par] = p; polyw[k * NCH + 2 + par] = q; polyw[k * NCH + 4 + par] = d; }
            else {
                polyw[k * NCH + par] = fma(rs * P.zni, q, p);
                if (MODE == QO_TF_S21) polyw[k * NCH + 2 + par] = d;
            }
        }
        __syncwarp();

        /* 3. frequency loop */
        double trk[QO_LAD_NSPEC];
#pragma unroll
        for (int sp = 0; sp < QO_LAD_NSPEC; sp++) trk[sp] = -1.7e308;
        for (int jb = 0; jb < npairs; jb += 32 * PP) {
            const int j0 = jb + lane;
            double x[PTS], y[PTS];
            unsigned int mk[PTS];
#pragma unroll
            for (int qq = 0; qq < PP; qq++) {
                const int j = j0 + 32 * qq;
                const int jc = j < npairs ? j : npairs - 1;
                const double2 a = P.xt[jc];
                const uchar2 m = P.m2[jc];
                x[2 * qq] = a.x; x[2 * qq + 1] = a.y;
                mk[2 * qq] = j < npairs ? m.x : 0u; mk[2 * qq + 1] = j < npairs ? m.y : 0u;
            }
            QO_PTS y[p] = -x[p] * x[p];
            double r[NCH][PTS];
            {
#pragma unroll
                for (int c = 0; c < NCH; c += 2) {
                    const LadV2<double> cc = lad_lds2(polys + ((K - 1) * NCH + c) * 8u, 0.0);
                    QO_PTS { r[c][p] = cc.x; r[c + 1][p] = cc.y; }
                }
            }
#pragma unroll
            for (int k = K - 2; k >= 0; k--) {
#pragma unroll
                for (int c = 0; c < NCH; c += 2) {
                    const LadV2<double> cc = lad_lds2(polys + (k * NCH + c) * 8u, 0.0);
                    QO_PTS { r[c][p] = fma(r[c][p], y[p], cc.x); r[c + 1][p] = fma(r[c + 1][p], y[p], cc.y); }
                }
            }
            double den2[PTS];
            if (!CPL) {
                QO_PTS { const double ni = r[1][p] * x[p]; den2[p] = fma(r[0][p], r[0][p], ni * ni); }
                if (MODE == QO_TF_S21) {
                    double dd[PTS], rd[PTS];
                    QO_PTS { const double di = r[3][p] * x[p]; dd[p] = fma(r[2][p], r[2][p], di * di); }
                    lad_rcp_batch<PTS>(dd, rd);
                    QO_PTS den2[p] *= rd[p];
                }
            } else {
                double w[PTS], tse[PTS], tce[PTS], tso[PTS], tco[PTS], cscale[PTS];
                QO_PTS { tse[p] = 0.0; tce[p] = 1.0; tso[p] = 0.0; tco[p] = 1.0; }
#pragma unroll
                for (int qq = 0; qq < PP; qq++) {
                    const int j = j0 + 32 * qq;
                    const int jc = j < npairs ? j : npairs - 1;
                    const double2 a = P.wt[jc];
                    w[2 * qq] = a.x; w[2 * qq + 1] = a.y;
                    if (P.cpl_fast) {
                        const double2 se = P.cse[jc], ce = P.cce[jc];
                        tse[2 * qq] = se.x; tse[2 * qq + 1] = se.y; tce[2 * qq] = ce.x; tce[2 * qq + 1] = ce.y;
                        if (!P.cpl_same) {
                            const double2 so = P.cso[jc], co = P.cco[jc];
                            tso[2 * qq] = so.x; tso[2 * qq + 1] = so.y; tco[2 * qq] = co.x; tco[2 * qq + 1] = co.y;
                        }
                    }
                }
                LadRow<double, PTS, 1> u;
                if (P.cpl_fast) lad_cpl_first<double, PTS, 1, true>(cpls, w, tse, tce, tso, tco, rs, u, cscale);
                else lad_cpl_first<double, PTS, 1, false>(cpls, w, tse, tce, tso, tco, rs, u, cscale);
                double dd[PTS], rd[PTS];
                const double zni = P.zni;
                QO_PTS {
                    const double pi_ = r[1][p] * x[p], qr = r[2][p] * zni, qi = r[3][p] * x[p] * zni, di = r[5][p] * x[p];
                    const double nr = fma(u.ar[0][p], r[0][p], fma(-u.ai[0][p], pi_, fma(u.br[0][p], qr, -u.bi[0][p] * qi)));
                    const double ni = fma(u.ar[0][p], pi_, fma(u.ai[0][p], r[0][p], fma(u.br[0][p], qi, u.bi[0][p] * qr)));
                    den2[p] = fma(nr, nr, ni * ni) * cscale[p];
                    dd[p] = fma(r[4][p], r[4][p], di * di);
                }
                lad_rcp_batch<PTS>(dd, rd);
                QO_PTS den2[p] *= rd[p];
            }
            /* trackers: as in qo_ladder.cuh (bands are contiguous: warp vote picks the select-free path) */
            unsigned int m_or = mk[0], m_and = mk[0];
#pragma unroll
            for (int p = 1; p < PTS; p++) { m_or |= mk[p]; m_and &= mk[p]; }
            const unsigned int any = __reduce_or_sync(0xffffffffu, m_or), all = __reduce_and_sync(0xffffffffu, m_and);
            if (any == all) {
#pragma unroll
                for (int sp = 0; sp < QO_LAD_NSPEC; sp++) {
                    if ((all >> sp) & 1u) {
                        if (P.neg[sp]) { QO_PTS { const double c = -den2[p]; trk[sp] = c > trk[sp] ? c : trk[sp]; } }
                        else { QO_PTS trk[sp] = den2[p] > trk[sp] ? den2[p] : trk[sp]; }
                    }
                }
            } else {
#pragma unroll
                for (int sp = 0; sp < QO_LAD_NSPEC; sp++) {
                    if ((any >> sp) & 1u) {
                        QO_PTS {
                            const double val = P.neg[sp] ? -den2[p] : den2[p];
                            const double cand = ((mk[p] >> sp) & 1u) ? val : -1.7e308;
                            trk[sp] = cand > trk[sp] ? cand : trk[sp];
                        }
                    }
                }
            }
        }

        /* 4. per-sample verdict */
        unsigned int fail = 0;
#pragma unroll
        for (int sp = 0; sp < QO_LAD_NSPEC; sp++) {
            if (sp < P.nspec) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double o = __shfl_xor_sync(0xffffffffu, trk[sp], off);
                    trk[sp] = o > trk[sp] ? o : trk[sp];
                }
                if (trk[sp] > P.thr[sp]) fail |= 1u << sp;
            }
        }
        if (lane == 0) {
            atomicAdd(&s_cnt[0], fail == 0 ? 1u : 0u);
            atomicAdd(&s_cnt[1], 1u);
            for (int sp = 0; sp < P.nspec; sp++)
                if ((fail >> sp) & 1u) atomicAdd(&s_cnt[2 + sp], 1u);
            if (P.hist_spec >= 0) {
                double worst = 0.0;
#pragma unroll
                for (int sp = 0; sp < QO_LAD_NSPEC; sp++) if (sp == P.hist_spec) worst = fabs(trk[sp]);
                const double k21 = P.k21;
                const double lin = P.hist_kind == SK_DEN2_MAX ? k21 * k21 / worst : k21 * k21 * (1.0 / worst);
                const double v = 10.0 * log10(lin);
                const double xb = (v - P.hist_lo) / (P.hist_hi - P.hist_lo) * (double)P.hist_bins;
                long long b = (long long)floor(xb);
                if (!(xb >= 0.0)) b = 0;
                if (b >= P.hist_bins) b = P.hist_bins - 1;
                atomicAdd(&s_cnt[2 + P.nspec + (int)b], 1u);
            }
        }
        s = __shfl_sync(0xffffffffu, s_next, 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncnt; i += TPB)
        if (s_cnt[i]) atomicAdd(&P.counters[i], (unsigned long long)s_cnt[i]);
}
