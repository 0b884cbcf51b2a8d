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
 * Jobs with |S11| specs add the second row vector v = [1 -Rs] . M1 ... Mk (NROWS = 2):
 *   S11 = (v . [Rl 1]^T) / den
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
#define QO_LAD_CPL 10                /* doubles of the coupler record that precedes the ladder records */
#define QO_LAD_NEG_HUGE_HI 0xFFEFFFFFu   /* high word of a huge negative finite double: "no point seen yet" */

/* NROWS = 1: the row vector u = [1 Rs] M (enough for |S21|); NROWS = 2 adds v = [1 -Rs] M, whose contraction
 * v . [Rl 1]^T is the numerator of S11, for jobs with |S11| specs */
template <int PTS, int NROWS> struct LadRow { double ar[NROWS][PTS], ai[NROWS][PTS], br[NROWS][PTS], bi[NROWS][PTS]; };
#define QO_ROWS _Pragma("unroll") for (int r = 0; r < NROWS; r++)

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
template <int PTS, int NROWS, bool FIRSTSTEP>
__device__ __forceinline__ void lad_ser_lossy_l(unsigned int cf, const double (&w)[PTS], const double (&w2)[PTS],
                                                double rs, LadRow<PTS, NROWS> &u)
{
    const double2 c01 = lad_lds2(cf), c23 = lad_lds2(cf + 16);
    const double R = lad_lds1(cf + 32);
    double dre[PTS], q[PTS], s[PTS];
    QO_PTS { dre[p] = fma(-w2[p], c01.x, 1.0); q[p] = fma(dre[p], dre[p], w2[p] * c01.y); }
    lad_rcp_batch<PTS>(q, s);
    QO_PTS {
        const double g = fma(c23.x, dre[p], -c23.y);
        const double zr = R * s[p], zi = (w[p] * s[p]) * g;
        QO_ROWS {
            if (FIRSTSTEP) {             /* u = [1 +-Rs]: b = +-Rs + Z */
                u.br[r][p] = (r ? -rs : rs) + zr; u.bi[r][p] = zi;
            } else {
                u.br[r][p] = fma(u.ar[r][p], zr, u.br[r][p]); u.br[r][p] = fma(-u.ai[r][p], zi, u.br[r][p]);
                u.bi[r][p] = fma(u.ar[r][p], zi, u.bi[r][p]); u.bi[r][p] = fma(u.ai[r][p], zr, u.bi[r][p]);
            }
        }
    }
}

/* shunt lossy capacitor; record = { 1/C, Ls, R, R^2 } */
template <int PTS, int NROWS, bool FIRSTSTEP>
__device__ __forceinline__ void lad_shunt_lossy_c(unsigned int cf, const double (&w)[PTS], const double (&wi)[PTS],
                                                  double rs, LadRow<PTS, NROWS> &u)
{
    const double2 c01 = lad_lds2(cf), c23 = lad_lds2(cf + 16);
    double x[PTS], q[PTS], s[PTS];
    QO_PTS { x[p] = fma(w[p], c01.y, -wi[p] * c01.x); q[p] = fma(x[p], x[p], c23.y); }
    lad_rcp_batch<PTS>(q, s);
    QO_PTS {
        const double yr = c23.x * s[p], yi = -x[p] * s[p];
        QO_ROWS {
            if (FIRSTSTEP) {             /* u = [1 +-Rs]: a = 1 +- Rs Y */
                const double rr = r ? -rs : rs;
                u.ar[r][p] = fma(rr, yr, 1.0); u.ai[r][p] = rr * yi;
            } else {
                u.ar[r][p] = fma(u.br[r][p], yr, u.ar[r][p]); u.ar[r][p] = fma(-u.bi[r][p], yi, u.ar[r][p]);
                u.ai[r][p] = fma(u.br[r][p], yi, u.ai[r][p]); u.ai[r][p] = fma(u.bi[r][p], yr, u.ai[r][p]);
            }
        }
    }
}

/* Coupled-line through section as the FIRST block, division-free.
 * Mode lines normalised to Zt: D_m = 2 cos(t_m) + j sin(t_m) (z_m + 1/z_m), n_m = j sin(t_m) (z_m - 1/z_m);
 * S21 = Sg/Pi, S11 = Nu/Pi with Pi = D_e D_o, Sg = D_e + D_o, Nu = (n_e D_o + n_o D_e)/2.  Symmetric S -> ABCD:
 *   A = (Pi^2 - Nu^2 + Sg^2)/k,  B = Zt ((Pi + Nu)^2 - Sg^2)/k,  C = ((Pi - Nu)^2 - Sg^2)/(Zt k),  k = 2 Sg Pi.
 * The complex scalar 1/k multiplies the whole matrix, hence den, so the chain runs on k*[A B; C A] and
 * |den|^2 is rescaled by 1/|k|^2 at the end (one batched real reciprocal instead of three complex divides).
 * record = { z_e + 1/z_e, (z_e - 1/z_e)/2, (odd: same two), te/w, to/w, Zt, Rs/Zt, te/w - nominal, to/w - nominal }
 *
 * Angles: t_m = k_m w.  When the plan could bound the perturbation |k_m - k_m,nominal| w_max <= 0.05 rad
 * (FAST), sin/cos of the NOMINAL angle come from per-frequency tables and the sample's small rotation from
 * short Taylor polynomials (|d|^10/10! < 3e-20, |d|^9/9! < 6e-18) -- 14 FP64 instructions instead of the ~30 of
 * a general sincos; otherwise sincos() is called. */
template <int PTS, int NROWS, bool FAST>
__device__ __forceinline__ void lad_cpl_first(unsigned int cf, const double (&w)[PTS], const double (&tse)[PTS], const double (&tce)[PTS],
                                              const double (&tso)[PTS], const double (&tco)[PTS], double rs, LadRow<PTS, NROWS> &u,
                                              double (&scale)[PTS])
{
    const double2 c01 = lad_lds2(cf), c23 = lad_lds2(cf + 16), c45 = lad_lds2(cf + 32), c67 = lad_lds2(cf + 48), c89 = lad_lds2(cf + 64);
    const double cE = c01.x, hE = c01.y, cO = c23.x, hO = c23.y, ke = c45.x, ko = c45.y, zt = c67.x, rz = c67.y;
    const bool same = ke == ko;
    double se[PTS], ce[PTS], so[PTS], co[PTS], kap2[PTS];
    QO_PTS {
        if (FAST) {
            {
                const double d = c89.x * w[p], z = d * d;
                const double cd = fma(fma(fma(fma(1.0 / 40320.0, z, -1.0 / 720.0), z, 1.0 / 24.0), z, -0.5), z, 1.0);
                const double sd = d * fma(fma(fma(-1.0 / 5040.0, z, 1.0 / 120.0), z, -1.0 / 6.0), z, 1.0);
                se[p] = fma(tse[p], cd, tce[p] * sd); ce[p] = fma(tce[p], cd, -tse[p] * sd);
            }
            if (same) { so[p] = se[p]; co[p] = ce[p]; }
            else {
                const double d = c89.y * w[p], z = d * d;
                const double cd = fma(fma(fma(fma(1.0 / 40320.0, z, -1.0 / 720.0), z, 1.0 / 24.0), z, -0.5), z, 1.0);
                const double sd = d * fma(fma(fma(-1.0 / 5040.0, z, 1.0 / 120.0), z, -1.0 / 6.0), z, 1.0);
                so[p] = fma(tso[p], cd, tco[p] * sd); co[p] = fma(tco[p], cd, -tso[p] * sd);
            }
        } else {
            sincos(ke * w[p], &se[p], &ce[p]);
            if (same) { so[p] = se[p]; co[p] = ce[p]; } else sincos(ko * w[p], &so[p], &co[p]);
        }
    }
    QO_PTS {
        const double a1 = ce[p] + ce[p], b1 = se[p] * cE, a2 = co[p] + co[p], b2 = so[p] * cO;        /* D_e, D_o */
        const double Pr = fma(a1, a2, -b1 * b2), Pi = fma(a1, b2, a2 * b1);                              /* Pi */
        const double Sr = a1 + a2, Si = b1 + b2;                                                         /* Sg */
        const double pe = se[p] * hE, po = so[p] * hO;
        const double Nr = -fma(pe, b2, po * b1), Ni = fma(pe, a2, po * a1);                              /* Nu */
        const double Xr = fma(Pr, Pr, -Pi * Pi), Xi = (Pr + Pr) * Pi;                                    /* Pi^2 */
        const double Yr = fma(Nr, Nr, -Ni * Ni), Yi = (Nr + Nr) * Ni;                                    /* Nu^2 */
        const double Wr = fma(Sr, Sr, -Si * Si), Wi = (Sr + Sr) * Si;                                    /* Sg^2 */
        const double Vr = fma(Pr, Nr, -Pi * Ni), Vi = fma(Pr, Ni, Pi * Nr);                              /* Pi Nu */
        const double Tr = Xr + Yr - Wr, Ti = Xi + Yi - Wi;
        const double Ar = Xr - Yr + Wr, Ai = Xi - Yi + Wi;                                               /* k A */
        const double Br = fma(2.0, Vr, Tr), Bi = fma(2.0, Vi, Ti);                                       /* k B / Zt */
        const double Cr = fma(-2.0, Vr, Tr), Ci = fma(-2.0, Vi, Ti);                                     /* k C Zt */
        const double Kr = fma(Sr, Pr, -Si * Pi), Ki = fma(Sr, Pi, Si * Pr);                              /* k / 2 */
        kap2[p] = fma(Kr, Kr, Ki * Ki);
        QO_ROWS {
            const double sg = r ? -1.0 : 1.0;
            u.ar[r][p] = fma(sg * rz, Cr, Ar); u.ai[r][p] = fma(sg * rz, Ci, Ai);                        /* k (A +- Rs C) */
            u.br[r][p] = fma(sg * rs, Ar, zt * Br); u.bi[r][p] = fma(sg * rs, Ai, zt * Bi);              /* k (B +- Rs A) */
        }
    }
    lad_rcp_batch<PTS>(kap2, scale);
    QO_PTS scale[p] *= 0.25;                                                                             /* 1/|k|^2 */
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
        out[0] = a + 1.0 / a; out[1] = 0.5 * (a - 1.0 / a); out[2] = b + 1.0 / b; out[3] = 0.5 * (b - 1.0 / b);
        out[4] = p[2] / (360.0 * p[4]); out[5] = p[3] / (360.0 * p[4]); out[6] = p[5]; out[7] = prog->rs / p[5];
        out[8] = out[4] - prog->nom[e][2] / (360.0 * prog->nom[e][4]); out[9] = out[5] - prog->nom[e][3] / (360.0 * prog->nom[e][4]);
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
    const double2 *cse, *cce, *cso, *cco;    /* coupler: sin/cos of the NOMINAL even/odd angle per grid point (cpl_fast) */
    unsigned long long *counters;
    unsigned long long *ticket;              /* next unclaimed sample of this launch (zeroed on the stream before it) */
    unsigned long long sample_offset, nsamples, seed;
    double rs, rl, k21, hist_lo, hist_hi;
    double thr[QO_LAD_NSPEC];                /* sign-adjusted thresholds: FAIL iff tracker > thr */
    unsigned int sgn[QO_LAD_NSPEC];          /* 0x80000000 for "max dB" specs (tracker holds -|den|^2) */
    int is_s11[QO_LAD_NSPEC];                /* the spec's tracker holds |S11|^2 = |n11|^2 / |den|^2 (NROWS == 2 kernels) */
    int npairs, n_var, n_ops, nspec, dist, hist_spec, hist_bins, hist_kind;
    int cpl_fast, cpl_same;                  /* small-angle table path usable; nominal even and odd angles identical */
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
template <int N, int FIRST, bool CPL, int NROWS, int PP, int TPB, int MINB>
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
            LadRow<PTS, NROWS> u;
            double cscale[PTS];
            QO_PTS { cscale[p] = 1.0; QO_ROWS { u.ar[r][p] = 1.0; u.ai[r][p] = 0.0; u.br[r][p] = r ? -rs : rs; u.bi[r][p] = 0.0; } }
            if (CPL) {
                double tse[PTS], tce[PTS], tso[PTS], tco[PTS];
                QO_PTS { tse[p] = 0.0; tce[p] = 1.0; tso[p] = 0.0; tco[p] = 1.0; }
                if (P.cpl_fast) {
#pragma unroll
                    for (int q = 0; q < PP; q++) {
                        const int j = j0 + 32 * q;
                        const int jc = j < npairs ? j : npairs - 1;
                        const double2 a = P.cse[jc], b = P.cce[jc];
                        tse[2 * q] = a.x; tse[2 * q + 1] = a.y; tce[2 * q] = b.x; tce[2 * q + 1] = b.y;
                        if (!P.cpl_same) {
                            const double2 c = P.cso[jc], d = P.cco[jc];
                            tso[2 * q] = c.x; tso[2 * q + 1] = c.y; tco[2 * q] = d.x; tco[2 * q + 1] = d.y;
                        }
                    }
                    lad_cpl_first<PTS, NROWS, true>(coefs, w, tse, tce, tso, tco, rs, u, cscale);
                } else {
                    lad_cpl_first<PTS, NROWS, false>(coefs, w, tse, tce, tso, tco, rs, u, cscale);
                }
            }
            const unsigned int lad = coefs + (CPL ? QO_LAD_CPL : 0) * 8u;
#pragma unroll
            for (int e = 0; e < N; e++) {
                const bool series = ((e + FIRST) & 1) == 0;
                const unsigned int cf = lad + e * QO_LAD_STRIDE * 8u;
                if (series) {
                    if (e == 0 && !CPL) lad_ser_lossy_l<PTS, NROWS, true>(cf, w, w2, rs, u);
                    else lad_ser_lossy_l<PTS, NROWS, false>(cf, w, w2, rs, u);
                } else {
                    if (e == 0 && !CPL) lad_shunt_lossy_c<PTS, NROWS, true>(cf, w, wi, rs, u);
                    else lad_shunt_lossy_c<PTS, NROWS, false>(cf, w, wi, rs, u);
                }
            }
            const double rl = P.rl;
            double den2[PTS], s11m[PTS];
            QO_PTS {
                const double den_r = fma(u.ar[0][p], rl, u.br[0][p]), den_i = fma(u.ai[0][p], rl, u.bi[0][p]);
                den2[p] = fma(den_r, den_r, den_i * den_i);
            }
            if (NROWS == 2) {
                /* |S11|^2 = |v . [Rl 1]|^2 / |den|^2 (the coupler's common factor 1/k cancels in the ratio) */
                double rd[PTS];
                lad_rcp_batch<PTS>(den2, rd);
                QO_PTS {
                    const double n_r = fma(u.ar[1][p], rl, u.br[1][p]), n_i = fma(u.ai[1][p], rl, u.bi[1][p]);
                    s11m[p] = fma(n_r, n_r, n_i * n_i) * rd[p];
                }
            }
            if (CPL) { QO_PTS den2[p] *= cscale[p]; }
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
                        if (NROWS == 2 && P.is_s11[sp]) { QO_PTS trk[sp] = s11m[p] > trk[sp] ? s11m[p] : trk[sp]; }
                        else if (P.sgn[sp]) { QO_PTS { const double c = -den2[p]; trk[sp] = c > trk[sp] ? c : trk[sp]; } }
                        else { QO_PTS trk[sp] = den2[p] > trk[sp] ? den2[p] : trk[sp]; }
                    }
                }
            } else {
#pragma unroll
                for (int sp = 0; sp < QO_LAD_NSPEC; sp++) {
                    if ((any >> sp) & 1u) {
                        QO_PTS {
                            const double val = (NROWS == 2 && P.is_s11[sp]) ? s11m[p] : den2[p];
                            const unsigned int hi = (unsigned int)__double2hiint(val);
                            const unsigned int chi = ((mk[p] >> sp) & 1u) ? (hi ^ P.sgn[sp]) : QO_LAD_NEG_HUGE_HI;
                            const double cand = __hiloint2double((int)chi, __double2loint(val));
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
                const double lin = P.hist_kind == SK_S11_MAX ? worst : P.hist_kind == SK_DEN2_MAX ? k21 * k21 / worst : k21 * k21 * (1.0 / worst);
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
