/*
 * qo_tf_fs.cuh -- transfer-function kernel, FULL_S flavour (sm_100a, FP64): all four S-parameters of every (sample,
 * frequency) point written to HBM (BASELINE config 4: util/gpsdo-ouput-filters/10M/schematic.svg:197-231 and
 * docs/gpsdo-filters/<name>.svg:195-241 swept with tolerances, 64 bytes per eval).
 *
 * The opcode interpreter spends its time in the per-element dispatch loop (ncu r01j: 72 % of the stall samples, DRAM at
 * 71 % of peak) -- FULL_S through it is latency-bound, not HBM-bound.  Here the cascade is expanded per sample into real
 * polynomials exactly as in qo_tf.cuh, for BOTH load vectors [Rl; 1] and [-Rl; 1]:
 *     [P ; Q ] = M [ Rl; 1] / D        den = (P + Rs Q)/D     S21 = S12 = 2 sqrt(Rs Rl) D / (P + Rs Q)
 *     [P'; Q'] = M [-Rl; 1] / D        S11 = (P - Rs Q)/(P + Rs Q)       S22 = (P' + Rs Q')/(P + Rs Q)
 * so a point costs four complex Horner evaluations (Num, N11, N22, D: eight real chains), one reciprocal and three
 * complex products -- independent of the number of branches -- and the kernel runs at the speed of its stores
 * (each lane owns two adjacent points = 32 contiguous bytes per plane, one 256-bit store per plane).
 */
#pragma once
#include "qo_tf.cuh"

struct TfFsParams {
    const DevProg *prog;
    const double2 *yt, *xt;                  /* -(w/wref)^2 and w/wref per grid point, two points per entry, padded */
    double2 *s11, *s21, *s12, *s22;          /* planes [nsamples][nf] (any may be NULL) */
    unsigned long long *ticket;
    unsigned long long sample_offset, nsamples, seed;
    double rs, rl, k21, wref, zn, zni;
    int nf, niter, kn, kd, n_var, n_el, el0, dist, planes_al32;
    int nfac;                                /* DMODE 2: branches whose denominator is not 1 ... */
    unsigned char fac[QO_TF_MAXEL];          /* ... and their indices (relative to el0) */
};

/* per-sample element record for one lane (same stream, same normalisation as qo_tf.cuh) */
__device__ __forceinline__ void tf_fs_derive(const DevProg *__restrict__ prog, int e, const double *__restrict__ x, double wr, double zn, double zni,
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

/* DMODE  0: every branch denominator is 1 (D == 1)
 *        1: D(jx) from its expanded polynomial (two more Horner chains)
 *        2: D(jx) as the PRODUCT of the branch denominators d0 + d2 y + j d1 x.  Next to the notch of a trap (|D| -> 0) the
 *           expanded polynomial loses the digits that the product keeps: its rounding error scales with prod(|d0| + |d2| x^2)
 *           while the product's scales with |D| of the resonating branch alone.  The plan picks this form whenever a branch
 *           resonates inside the grid (elliptic filters), S21 next to a notch then matches the per-element chain to rounding. */
template <int DMODE, int PP, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) qo_fs_tf_kernel(const __grid_constant__ TfFsParams P)
{
    constexpr int PTS = 2 * PP;
    constexpr int WARPS = TPB / 32;
    constexpr bool HASD = DMODE == 1;          /* D rides along as Horner chains */
    constexpr int NCH = HASD ? 8 : 6;
    __shared__ __align__(16) double s_tab[WARPS][QO_TF_MAXK * NCH];      /* row k: Num, N11, N22 (, D) coefficients of sn^(2k), sn^(2k+1) */
    __shared__ __align__(16) double s_el[WARPS][QO_TF_MAXEL * 8];
    __shared__ double s_x[WARPS][QO_MAX_VAR];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *tabw = s_tab[warp], *elw = s_el[warp], *xw = s_x[warp];
    const unsigned int tabs = (unsigned int)__cvta_generic_to_shared(tabw);
    const int up1 = (lane + 31) & 31, up2 = (lane + 30) & 31;
    const int kn = P.kn, kd = HASD ? P.kd : 0, krows = kn > kd ? kn : kd;
    const int nf = P.nf;
    const double zq = P.rs * P.zni, k21 = P.k21;

    const unsigned long long total_warps = (unsigned long long)gridDim.x * WARPS;
    unsigned long long s = (unsigned long long)blockIdx.x * WARPS + warp;
    while (s < P.nsamples) {
        unsigned long long s_next = 0;
        if (lane == 0) s_next = total_warps + atomicAdd(P.ticket, 1ull);
        for (int v = lane; v < P.n_var; v += 32) xw[v] = qo_stream_variate(P.seed, P.sample_offset + s, (uint32_t)v, P.dist);
        __syncwarp();
        if (lane < P.n_el) tf_fs_derive(P.prog, P.el0 + lane, xw, P.wref, P.zn, P.zni, elw + lane * 8);
        __syncwarp();
        /* expansion from the load end for both load vectors; lane i holds the coefficient of sn^i */
        double p = lane == 0 ? P.rl : 0.0, q = lane == 0 ? P.zn : 0.0, p2 = -p, q2 = q, d = lane == 0 ? 1.0 : 0.0;
        for (int e = P.n_el - 1; e >= 0; e--) {
            const double2 n01 = *(const double2 *)(elw + e * 8), n2d0 = *(const double2 *)(elw + e * 8 + 2), d12 = *(const double2 *)(elw + e * 8 + 4);
            const bool series = elw[e * 8 + 6] != 0.0;
            const double pa = tf_up(p, up1), pb = tf_up(p, up2), qa = tf_up(q, up1), qb = tf_up(q, up2);
            const double ra = tf_up(p2, up1), rb = tf_up(p2, up2), sa = tf_up(q2, up1), sb = tf_up(q2, up2);
            const double dp = fma(n2d0.y, p, fma(d12.x, pa, d12.y * pb)), dq = fma(n2d0.y, q, fma(d12.x, qa, d12.y * qb));
            const double dr = fma(n2d0.y, p2, fma(d12.x, ra, d12.y * rb)), ds = fma(n2d0.y, q2, fma(d12.x, sa, d12.y * sb));
            if (series) {
                p = fma(n01.x, q, fma(n01.y, qa, fma(n2d0.x, qb, dp))); q = dq;
                p2 = fma(n01.x, q2, fma(n01.y, sa, fma(n2d0.x, sb, dr))); q2 = ds;
            } else {
                q = fma(n01.x, p, fma(n01.y, pa, fma(n2d0.x, pb, dq))); p = dp;
                q2 = fma(n01.x, p2, fma(n01.y, ra, fma(n2d0.x, rb, ds))); p2 = dr;
            }
            if (HASD) {
                const double d1 = tf_up(d, up1), d2 = tf_up(d, up2);
                d = fma(n2d0.y, d, fma(d12.x, d1, d12.y * d2));
            }
        }
        if (lane < 2 * krows) {
            const int k = lane >> 1, par = lane & 1;
            const bool nk = k < kn;
            tabw[k * NCH + par] = nk ? fma(zq, q, p) : 0.0;             /* Num = P + Rs Q */
            tabw[k * NCH + 2 + par] = nk ? fma(-zq, q, p) : 0.0;        /* N11 = P - Rs Q */
            tabw[k * NCH + 4 + par] = nk ? fma(zq, q2, p2) : 0.0;       /* N22 = P' + Rs Q' */
            if (HASD) tabw[k * NCH + 6 + par] = k < kd ? d : 0.0;
        }
        __syncwarp();

        for (int it = 0; it < P.niter; it++) {
            const int j0 = it * (32 * PP) + lane;
            double y[PTS], x[PTS];
#pragma unroll
            for (int qq = 0; qq < PP; qq++) {
                const double2 a = P.yt[j0 + 32 * qq], b = P.xt[j0 + 32 * qq];
                y[2 * qq] = a.x; y[2 * qq + 1] = a.y; x[2 * qq] = b.x; x[2 * qq + 1] = b.y;
            }
            double r[NCH][PTS];
            unsigned int a = tabs + (unsigned int)(krows - 1) * (NCH * 8u);
#pragma unroll
            for (int c = 0; c < NCH; c += 2) {
                const LadV2<double> cc = lad_lds2(a + c * 8u, 0.0);
                QO_PTS { r[c][p] = cc.x; r[c + 1][p] = cc.y; }
            }
#pragma unroll 2
            for (int k = krows - 2; k >= 0; k--) {
                a -= NCH * 8u;
#pragma unroll
                for (int c = 0; c < NCH; c += 2) {
                    const LadV2<double> cc = lad_lds2(a + c * 8u, 0.0);
                    QO_PTS { r[c][p] = fma(r[c][p], y[p], cc.x); r[c + 1][p] = fma(r[c + 1][p], y[p], cc.y); }
                }
            }
            double2 o11[PTS], o21[PTS], o22[PTS];
            double n2[PTS], rn[PTS];
            QO_PTS { r[1][p] *= x[p]; r[3][p] *= x[p]; r[5][p] *= x[p]; if (HASD) r[7][p] *= x[p]; n2[p] = fma(r[0][p], r[0][p], r[1][p] * r[1][p]); }
            lad_rcp_batch<PTS>(n2, rn);
            QO_PTS {
                const double ir = r[0][p] * rn[p], ii = -r[1][p] * rn[p];                /* 1 / Num */
                o11[p] = make_double2(fma(r[2][p], ir, -r[3][p] * ii), fma(r[2][p], ii, r[3][p] * ir));
                o22[p] = make_double2(fma(r[4][p], ir, -r[5][p] * ii), fma(r[4][p], ii, r[5][p] * ir));
                if (HASD) o21[p] = make_double2(k21 * fma(r[6][p], ir, -r[7][p] * ii), k21 * fma(r[6][p], ii, r[7][p] * ir));
                else if (DMODE == 0) o21[p] = make_double2(k21 * ir, k21 * ii);
            }
            if (DMODE == 2) {
                double dre[PTS], dim[PTS];
                QO_PTS { dre[p] = k21; dim[p] = 0.0; }
                for (int t = 0; t < P.nfac; t++) {
                    const double *rec = elw + (int)P.fac[t] * 8;
                    const double d0 = rec[3], d1 = rec[4], d2 = rec[5];
                    QO_PTS {
                        const double fr = fma(d2, y[p], d0), fi = d1 * x[p];
                        const double nr = fma(dre[p], fr, -dim[p] * fi), ni = fma(dre[p], fi, dim[p] * fr);
                        dre[p] = nr; dim[p] = ni;
                    }
                }
                QO_PTS {
                    const double ir = r[0][p] * rn[p], ii = -r[1][p] * rn[p];
                    o21[p] = make_double2(fma(dre[p], ir, -dim[p] * ii), fma(dre[p], ii, dim[p] * ir));
                }
            }
#pragma unroll
            for (int qq = 0; qq < PP; qq++) {
                const int k = 2 * (j0 + 32 * qq);
                if (k >= nf) continue;
                const size_t o = (size_t)s * (size_t)nf + (size_t)k;
                if (k + 1 < nf && ((o & 1) == 0) && P.planes_al32) {
                    if (P.s21) qo_st256(P.s21 + o, o21[2 * qq], o21[2 * qq + 1]);
                    if (P.s11) qo_st256(P.s11 + o, o11[2 * qq], o11[2 * qq + 1]);
                    if (P.s22) qo_st256(P.s22 + o, o22[2 * qq], o22[2 * qq + 1]);
                    if (P.s12) qo_st256(P.s12 + o, o21[2 * qq], o21[2 * qq + 1]);
                } else {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (k + h < nf) {
                            if (P.s21) P.s21[o + h] = o21[2 * qq + h];
                            if (P.s11) P.s11[o + h] = o11[2 * qq + h];
                            if (P.s22) P.s22[o + h] = o22[2 * qq + h];
                            if (P.s12) P.s12[o + h] = o21[2 * qq + h];
                        }
                    }
                }
            }
        }
        s = __shfl_sync(0xffffffffu, s_next, 0);
    }
}
