/*
 * qo_ladder.cuh -- the straight-line Monte-Carlo kernel for the pcb/generic-filter
 * ladder family (reference: pcb/generic-filter/README.md:13 "up to 11th order",
 * qo-100-generic-filter.sch:1450-1488,1703-1995 -- alternating series / shunt
 * branches), optionally preceded by the coupled-line through section of
 * util/directional-couplers/*.trc (BASELINE configs 2 and 5).  sm_100a only.
 *
 * Why a second kernel next to the opcode interpreter of qo_lumped.cuh: ncu on the
 * interpreter (profiles/r01a_*) shows the FP64 pipe 47 % busy with 53 % of the
 * issued instructions being dispatch overhead (BRA/ISETP/IMAD.MOV/LDS) and the
 * dominant stall a fixed-latency wait on the serial immittance -> reciprocal ->
 * chain dependency of ONE element.  Here the element sequence is a template
 * parameter, so the whole ladder is one basic block: ptxas overlaps the
 * reciprocal of element k+1 with the chain update of element k, and the only
 * non-FP64 work left is the broadcast LDS of the per-sample coefficients.
 *
 * Arithmetic (all FP64, one thread = PTS (sample, frequency) points of ONE sample):
 *   |S21|^2 = 4 Rs Rl / |den|^2,  den = [1 Rs] . M1 M2 ... MN . [Rl 1]^T
 * Reduce-only specs on |S21| need den only, so the chain carries the ROW VECTOR
 * u = [1 Rs] . M1 ... Mk = (a, b) instead of the 2x2 product (half the FMAs):
 *   series Z: b += a Z        shunt Y: a += b Y        den = a Rl + b
 * Series lossy inductor (R + jwL) || 1/(jwCp), with D = 1 - w^2 L Cp:
 *   Z = (R + j w (L D - R^2 Cp)) / (D^2 + w^2 (R Cp)^2)       [numerator real part is exactly R]
 * Shunt lossy capacitor R + jwLs + 1/(jwC), X = w Ls - (1/w)(1/C):
 *   Y = (R - jX) / (R^2 + X^2)
 * Executed FP64 instructions per eval for the 11th-order ladder: 6*15 + 5*12 - 3
 * (first step) + 4 (den, |den|^2) + nspec compares = 151 + nspec.
 *
 * Spec bookkeeping: per spec one running extreme of |den|^2 over the spec's band
 * (sign-flipped for "max dB" specs so that every tracker is a running MAX),
 * compared with the threshold once per sample; the histogram variable is the
 * tracker of its spec, so it costs nothing extra.
 */
#pragma once
#include "qo_lumped.cuh"

#define QO_LAD_MAXN 11
#define QO_LAD_NSPEC 4               /* trackers kept in registers; more specs -> interpreter */
#define QO_LAD_STRIDE 6              /* doubles per element record (16-byte aligned) */
#define QO_LAD_CPL 8                 /* doubles of the coupler record that precedes the ladder records */
#define QO_LAD_NEG_HUGE_HI 0xFFEFFFFFu   /* high word of a huge negative finite double: "no point seen yet" */

template <int PTS> struct LadRow { double ar[PTS], ai[PTS], br[PTS], bi[PTS]; };

/* Broadcast read of two coefficients from the warp's record table.  The records are invariant over
 * the frequency loop, and left to itself ptxas hoists all 55 doubles of an 11-element ladder out of
 * the loop and spills them to local memory; an asm volatile load stays where it is written (one
 * conflict-free LDS.64x2 wavefront per element per iteration). */
__device__ __forceinline__ double2 lad_lds2(unsigned int saddr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ double lad_lds1(unsigned int saddr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
    return v;
}

#define QO_PTS _Pragma("unroll") for (int p = 0; p < PTS; p++)

/* 1/q for the PTS points of one element with ONE reciprocal (Montgomery's trick): the MUFU.RCP64H seed
 * and the register move that zeroes its low word are not FP64-pipe instructions, and tools/pipe_probe2.cu
 * shows each such instruction costs about one FP64 issue cycle when it does not land in a DFMA's shadow.
 * FP64 work is unchanged (9 DMUL + 3 DFMA per four points = 3 per point); each 1/q carries <= 3 roundings.
 * The product of four |immittance|^2 values stays far inside the double range for any physical element. */
template <int PTS> __device__ __forceinline__ void lad_rcp_batch(const double (&q)[PTS], double (&s)[PTS])
{
    static_assert(PTS == 2 || PTS == 4, "two or four points per thread");
    if (PTS == 2) {
        const double r = qrcp(q[0] * q[1]);
        s[0] = r * q[1]; s[1] = r * q[0];
    } else {
        const double p01 = q[0] * q[1], p23 = q[2] * q[3];
        const double r = qrcp(p01 * p23);
        const double r01 = r * p23, r23 = r * p01;
        s[0] = r01 * q[1]; s[1] = r01 * q[0]; s[2] = r23 * q[3]; s[3] = r23 * q[2];
    }
}

/* series lossy inductor; record = { L*Cp, (R*Cp)^2, L, R^2*Cp, R, - } */
template <int PTS, bool FIRSTSTEP>
__device__ __forceinline__ void lad_ser_lossy_l(unsigned int cf, const double (&w)[PTS], const double (&w2)[PTS],
                                                double rs, LadRow<PTS> &u)
{
    const double2 c01 = lad_lds2(cf), c23 = lad_lds2(cf + 16);
    const double R = lad_lds1(cf + 32);
    double dre[PTS], q[PTS], s[PTS];
    QO_PTS { dre[p] = fma(-w2[p], c01.x, 1.0); q[p] = fma(dre[p], dre[p], w2[p] * c01.y); }
    lad_rcp_batch<PTS>(q, s);
    QO_PTS {
        const double g = fma(c23.x, dre[p], -c23.y);
        const double zr = R * s[p], zi = (w[p] * s[p]) * g;
        if (FIRSTSTEP) {                 /* u = [1 Rs]: b = Rs + Z */
            u.br[p] = rs + zr; u.bi[p] = zi;
        } else {
            u.br[p] = fma(u.ar[p], zr, u.br[p]); u.br[p] = fma(-u.ai[p], zi, u.br[p]);
            u.bi[p] = fma(u.ar[p], zi, u.bi[p]); u.bi[p] = fma(u.ai[p], zr, u.bi[p]);
        }
    }
}

/* shunt lossy capacitor; record = { 1/C, Ls, R, R^2 } */
template <int PTS, bool FIRSTSTEP>
__device__ __forceinline__ void lad_shunt_lossy_c(unsigned int cf, const double (&w)[PTS], const double (&wi)[PTS],
                                                  double rs, LadRow<PTS> &u)
{
    const double2 c01 = lad_lds2(cf), c23 = lad_lds2(cf + 16);
    double x[PTS], q[PTS], s[PTS];
    QO_PTS { x[p] = fma(w[p], c01.y, -wi[p] * c01.x); q[p] = fma(x[p], x[p], c23.y); }
    lad_rcp_batch<PTS>(q, s);
    QO_PTS {
        const double yr = c23.x * s[p], yi = -x[p] * s[p];
        if (FIRSTSTEP) {                 /* u = [1 Rs]: a = 1 + Rs Y */
            u.ar[p] = fma(rs, yr, 1.0); u.ai[p] = rs * yi;
        } else {
            u.ar[p] = fma(u.br[p], yr, u.ar[p]); u.ar[p] = fma(-u.bi[p], yi, u.ar[p]);
            u.ai[p] = fma(u.br[p], yi, u.ai[p]); u.ai[p] = fma(u.bi[p], yr, u.ai[p]);
        }
    }
}

/* coupled-line through section as the FIRST block: u = [1 Rs] . [A B; C A]
 * record = { z0e/zt + zt/z0e, z0e/zt - zt/z0e, (same for odd), te/w, to/w, zt, 1/zt }   */
template <int PTS>
__device__ __forceinline__ void lad_cpl_first(unsigned int cf, const double (&w)[PTS], double rs, LadRow<PTS> &u)
{
    const double2 c01 = lad_lds2(cf), c23 = lad_lds2(cf + 16), c45 = lad_lds2(cf + 32), c67 = lad_lds2(cf + 48);
    const double cE = c01.x, dE = c01.y, cO = c23.x, dO = c23.y, ke = c45.x, ko = c45.y, zt = c67.x, yt = c67.y;
    double se[PTS], ce[PTS], so[PTS], co[PTS], er_[PTS], ei_[PTS], or_[PTS], oi_[PTS], qe[PTS], qo[PTS], re[PTS], ro[PTS];
    QO_PTS {
        sincos(ke * w[p], &se[p], &ce[p]);
        if (ke == ko) { so[p] = se[p]; co[p] = ce[p]; } else sincos(ko * w[p], &so[p], &co[p]);
        er_[p] = ce[p] + ce[p]; ei_[p] = se[p] * cE; or_[p] = co[p] + co[p]; oi_[p] = so[p] * cO;
        qe[p] = fma(er_[p], er_[p], ei_[p] * ei_[p]); qo[p] = fma(or_[p], or_[p], oi_[p] * oi_[p]);
    }
    lad_rcp_batch<PTS>(qe, re);
    lad_rcp_batch<PTS>(qo, ro);
    double s21r[PTS], s21i[PTS], s11r[PTS], s11i[PTS], dr_[PTS], di_[PTS], qd[PTS], rd[PTS];
    QO_PTS {
        const double ier = er_[p] * re[p], iei = -ei_[p] * re[p], ior = or_[p] * ro[p], ioi = -oi_[p] * ro[p];
        s21r[p] = ier + ior; s21i[p] = iei + ioi;
        const double ge = 0.5 * se[p] * dE, go = 0.5 * so[p] * dO;
        s11r[p] = -(ge * iei + go * ioi); s11i[p] = ge * ier + go * ior;
        dr_[p] = s21r[p] + s21r[p]; di_[p] = s21i[p] + s21i[p];
        qd[p] = fma(dr_[p], dr_[p], di_[p] * di_[p]);
    }
    lad_rcp_batch<PTS>(qd, rd);
    QO_PTS {
        const double q2r = s21r[p] * s21r[p] - s21i[p] * s21i[p], q2i = 2.0 * s21r[p] * s21i[p];
        const double p2r = s11r[p] * s11r[p] - s11i[p] * s11i[p], p2i = 2.0 * s11r[p] * s11i[p];
        const double idr = dr_[p] * rd[p], idi = -di_[p] * rd[p];
        const double nar = 1.0 - p2r + q2r, nai = q2i - p2i;
        const double nbr = 1.0 + s11r[p] + s11r[p] + p2r - q2r, nbi = s11i[p] + s11i[p] + p2i - q2i;
        const double ncr = 1.0 - s11r[p] - s11r[p] + p2r - q2r, nci = -s11i[p] - s11i[p] + p2i - q2i;
        const double Ar = nar * idr - nai * idi, Ai = nar * idi + nai * idr;
        const double Br = zt * (nbr * idr - nbi * idi), Bi = zt * (nbr * idi + nbi * idr);
        const double Cr = yt * (ncr * idr - nci * idi), Ci = yt * (ncr * idi + nci * idr);
        u.ar[p] = fma(rs, Cr, Ar); u.ai[p] = fma(rs, Ci, Ai);
        u.br[p] = fma(rs, Ar, Br); u.bi[p] = fma(rs, Ai, Bi);
    }
}

/* per-sample coefficient records, one lane per element (perturbation is the shared bit-exact stream) */
__device__ __forceinline__ void lad_derive(const DevProg *__restrict__ prog, int e, const double *__restrict__ x, double *out)
{
    double p[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        p[k] = prog->nom[e][k];
        const int tv = prog->tvar[e][k];
        if (tv >= 0) p[k] = qo_stream_apply(p[k], prog->ttol[e][k], x[tv], prog->tmode[e][k]);
    }
    switch (prog->opcode[e]) {
    case OP_SER_LOSSY_L: case OP_SER_L: {              /* p = L, R, Cp */
        const double rcp_ = p[1] * p[2];
        out[0] = p[0] * p[2]; out[1] = rcp_ * rcp_; out[2] = p[0]; out[3] = p[1] * rcp_; out[4] = p[1]; out[5] = 0.0;
        break;
    }
    case OP_SHUNT_LOSSY_C: case OP_SHUNT_C:            /* p = C, R, Ls */
        out[0] = 1.0 / p[0]; out[1] = p[2]; out[2] = p[1]; out[3] = p[1] * p[1]; out[4] = 0.0; out[5] = 0.0;
        break;
    case OP_CPL: {
        const double a = p[0] / p[5], b = p[1] / p[5];
        out[0] = a + 1.0 / a; out[1] = a - 1.0 / a; out[2] = b + 1.0 / b; out[3] = b - 1.0 / b;
        out[4] = p[2] / (360.0 * p[4]); out[5] = p[3] / (360.0 * p[4]); out[6] = p[5]; out[7] = 1.0 / p[5];
        break;
    }
    default: break;
    }
}

/* everything warp-uniform travels as a kernel parameter (constant bank -> uniform registers, no
 * vector registers held across the frequency loop) */
struct LadParams {
    const DevProg *prog;
    const double2 *wt, *wit, *wsqt;          /* w, 1/w, w^2 per grid point, two points per entry */
    const uchar2 *m2;                        /* per-point spec bit masks */
    unsigned long long *counters;
    unsigned long long *ticket;              /* next unclaimed sample of this launch (zeroed on the stream before it) */
    unsigned long long sample_offset, nsamples, seed;
    double rs, rl, k21, hist_lo, hist_hi;
    double thr[QO_LAD_NSPEC];                /* sign-adjusted thresholds: FAIL iff tracker > thr */
    unsigned int sgn[QO_LAD_NSPEC];          /* 0x80000000 for "max dB" specs (tracker holds -|den|^2) */
    int npairs, n_var, n_ops, nspec, dist, hist_spec, hist_bins, hist_kind;
};

/*
 * N      ladder elements (1..11)        FIRST  0 = series element first, 1 = shunt first
 * CPL    coupled-line block in front    PP     frequency pairs per thread per iteration (PTS = 2*PP points)
 * TPB / MINB  block size and resident blocks per SM (register budget = 65536 / (TPB*MINB))
 * One warp = one sample at a time; lane l owns pairs l, l+32, ... of that sample's grid.
 * Samples are handed out through a global ticket counter: with a static split ncu showed every
 * SMSP averaging 3.06 of its 4 warps (the issue scheduler is not fair, favoured warps finished
 * their share early and the FP64 pipe drained); tickets keep all warps busy until the pool is empty.
 */
template <int N, int FIRST, bool CPL, int PP, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) qo_mc_ladder_kernel(const __grid_constant__ LadParams P)
{
    constexpr int PTS = 2 * PP;
    constexpr int WARPS = TPB / 32;
    constexpr int NREC = (CPL ? QO_LAD_CPL : 0) + N * QO_LAD_STRIDE;
    __shared__ __align__(16) double s_coef[WARPS][NREC];
    __shared__ double s_x[WARPS][QO_MAX_VAR];
    __shared__ unsigned int s_cnt[2 + QO_NSPEC_MAX + QO_MAX_HIST];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncnt = 2 + P.nspec + (P.hist_bins > 0 ? P.hist_bins : 0);
    for (int i = threadIdx.x; i < ncnt; i += TPB) s_cnt[i] = 0;
    __syncthreads();

    double *coefw = s_coef[warp];
    double *xw = s_x[warp];
    const unsigned int coefs = (unsigned int)__cvta_generic_to_shared(coefw);
    const int npairs = P.npairs;
    const double rs = P.rs;

    /* the first gridDim.x*WARPS samples are claimed by position, the rest by ticket; the next ticket
     * is requested before the frequency loop so that the atomic's latency is never exposed */
    const unsigned long long total_warps = (unsigned long long)gridDim.x * WARPS;
    unsigned long long s = (unsigned long long)blockIdx.x * WARPS + warp;
    while (s < P.nsamples) {
        unsigned long long s_next = 0;
        if (lane == 0) s_next = total_warps + atomicAdd(P.ticket, 1ull);
        /* 1. the sample's random variables (Philox, counter-based) and coefficient records */
        for (int v = lane; v < P.n_var; v += 32) xw[v] = qo_stream_variate(P.seed, P.sample_offset + s, (uint32_t)v, P.dist);
        __syncwarp();
        if (lane < P.n_ops) {
            const int rec = CPL ? (lane == 0 ? 0 : QO_LAD_CPL + (lane - 1) * QO_LAD_STRIDE) : lane * QO_LAD_STRIDE;
            lad_derive(P.prog, lane, xw, coefw + rec);
        }
        __syncwarp();

        /* 2. frequency loop */
        double trk[QO_LAD_NSPEC];
#pragma unroll
        for (int sp = 0; sp < QO_LAD_NSPEC; sp++) trk[sp] = __hiloint2double((int)QO_LAD_NEG_HUGE_HI, 0);
        /* warp-uniform trip count (the tracker vote below is a full-warp collective): lanes past the end of
         * the grid re-evaluate the last pair with all-zero masks */
        for (int jb = 0; jb < npairs; jb += 32 * PP) {
            const int j0 = jb + lane;
            double w[PTS], wi[PTS], w2[PTS];
            unsigned int mk[PTS];
#pragma unroll
            for (int q = 0; q < PP; q++) {
                const int j = j0 + 32 * q;
                const int jc = j < npairs ? j : npairs - 1;
                const double2 a = P.wt[jc], b = P.wit[jc], c = P.wsqt[jc];
                const uchar2 m = P.m2[jc];
                w[2 * q] = a.x; w[2 * q + 1] = a.y; wi[2 * q] = b.x; wi[2 * q + 1] = b.y; w2[2 * q] = c.x; w2[2 * q + 1] = c.y;
                mk[2 * q] = j < npairs ? m.x : 0u; mk[2 * q + 1] = j < npairs ? m.y : 0u;
            }
            LadRow<PTS> u;
            QO_PTS { u.ar[p] = 1.0; u.ai[p] = 0.0; u.br[p] = rs; u.bi[p] = 0.0; }
            if (CPL) lad_cpl_first<PTS>(coefs, w, rs, u);
            const unsigned int lad = coefs + (CPL ? QO_LAD_CPL : 0) * 8u;
#pragma unroll
            for (int e = 0; e < N; e++) {
                const bool series = ((e + FIRST) & 1) == 0;
                const unsigned int cf = lad + e * QO_LAD_STRIDE * 8u;
                if (series) {
                    if (e == 0 && !CPL) lad_ser_lossy_l<PTS, true>(cf, w, w2, rs, u);
                    else lad_ser_lossy_l<PTS, false>(cf, w, w2, rs, u);
                } else {
                    if (e == 0 && !CPL) lad_shunt_lossy_c<PTS, true>(cf, w, wi, rs, u);
                    else lad_shunt_lossy_c<PTS, false>(cf, w, wi, rs, u);
                }
            }
            const double rl = P.rl;
            double den2[PTS];
            QO_PTS {
                const double den_r = fma(u.ar[p], rl, u.br[p]), den_i = fma(u.ai[p], rl, u.bi[p]);
                den2[p] = fma(den_r, den_r, den_i * den_i);
            }
            /* Trackers.  Spec bands are contiguous in frequency, so nearly every warp-iteration sees ONE
             * mask value on all its points: a warp vote picks the fast path (a plain running max per active
             * spec, no per-point selects); iterations that straddle a band edge take the general path. */
            unsigned int m_or = mk[0], m_and = mk[0];
#pragma unroll
            for (int p = 1; p < PTS; p++) { m_or |= mk[p]; m_and &= mk[p]; }
            const unsigned int any = __reduce_or_sync(0xffffffffu, m_or), all = __reduce_and_sync(0xffffffffu, m_and);
            if (any == all) {
#pragma unroll
                for (int sp = 0; sp < QO_LAD_NSPEC; sp++) {
                    if ((all >> sp) & 1u) {
                        if (P.sgn[sp]) { QO_PTS { const double c = -den2[p]; trk[sp] = c > trk[sp] ? c : trk[sp]; } }
                        else { QO_PTS trk[sp] = den2[p] > trk[sp] ? den2[p] : trk[sp]; }
                    }
                }
            } else {
#pragma unroll
                for (int sp = 0; sp < QO_LAD_NSPEC; sp++) {
                    if ((any >> sp) & 1u) {
                        QO_PTS {
                            const unsigned int hi = (unsigned int)__double2hiint(den2[p]);
                            const unsigned int chi = ((mk[p] >> sp) & 1u) ? (hi ^ P.sgn[sp]) : QO_LAD_NEG_HUGE_HI;
                            const double cand = __hiloint2double((int)chi, __double2loint(den2[p]));
                            trk[sp] = cand > trk[sp] ? cand : trk[sp];
                        }
                    }
                }
            }
        }

        /* 3. per-sample verdict: warp max of every tracker, then shared-memory counters */
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
                double worst = 0.0;      /* |den|^2 extreme of the histogram spec's band */
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
