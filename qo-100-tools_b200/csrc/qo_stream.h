/*
 * qo_stream.h -- counter-based Philox4x32-10 perturbation stream, compiled for
 * both the host (C-ABI twins qo_philox4x32_10 / qo_variate / qo_perturb_factor)
 * and the device (Monte-Carlo kernels).  Every floating-point step is a single
 * correctly-rounded IEEE operation spelled explicitly (no contraction on either
 * side), so host and device produce identical bits.  Stream contract:
 *   key = (seed lo32, seed hi32); ctr = (sample lo32, sample hi32, var>>1, 0);
 *   words (x0,x1) feed an even var, (x2,x3) an odd var;
 *   k = ((hi<<32 | lo) >> 11);  uniform: x = fma(2, k*2^-53, -1);
 *   gauss3s: x = clamp(norminv(fma(k, 2^-53, 2^-54)), -3, 3) / 3.
 */
#ifndef QO_STREAM_H
#define QO_STREAM_H
#include <stdint.h>

#if defined(__CUDACC__)
#define QO_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define QO_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define QO_MUL(a, b) __dmul_rn((a), (b))
#define QO_ADD(a, b) __dadd_rn((a), (b))
#define QO_SUB(a, b) __dadd_rn((a), -(b))
#define QO_DIV(a, b) __ddiv_rn((a), (b))
#define QO_FMA(a, b, c) __fma_rn((a), (b), (c))
#define QO_SQRT(a) __dsqrt_rn(a)
#define QO_MULHI(a, b) __umulhi((a), (b))
#define QO_U2D(k) __ull2double_rn(k)
#else
#define QO_MUL(a, b) ((a) * (b))
#define QO_ADD(a, b) ((a) + (b))
#define QO_SUB(a, b) ((a) - (b))
#define QO_DIV(a, b) ((a) / (b))
#define QO_FMA(a, b, c) fma((a), (b), (c))
#define QO_SQRT(a) sqrt(a)
#define QO_MULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#define QO_U2D(k) ((double)(k))
#endif

QO_HD void qo_philox_rounds(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = QO_MULHI(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = QO_MULHI(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

QO_HD uint64_t qo_stream_bits53(uint64_t seed, uint64_t sample, uint32_t var)
{
    uint32_t c[4] = { (uint32_t)sample, (uint32_t)(sample >> 32), var >> 1, 0u };
    qo_philox_rounds(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t w = (var & 1u) ? (((uint64_t)c[3] << 32) | c[2]) : (((uint64_t)c[1] << 32) | c[0]);
    return w >> 11;
}

/* natural log from + - * / fma only (atanh series on m in [sqrt(1/2), sqrt(2))) */
QO_HD double qo_log_det(double x)
{
#if defined(__CUDA_ARCH__)
    uint64_t b = (uint64_t)__double_as_longlong(x);
#else
    union { double d; uint64_t u; } cv; cv.d = x; uint64_t b = cv.u;
#endif
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
#if defined(__CUDA_ARCH__)
    double m = __longlong_as_double((long long)b);
#else
    cv.u = b; double m = cv.d;
#endif
    if (m > 1.4142135623730951) { m = QO_MUL(m, 0.5); e += 1; }
    double s = QO_DIV(QO_SUB(m, 1.0), QO_ADD(m, 1.0));
    double z = QO_MUL(s, s);
    double p = 1.0 / 23.0;
    p = QO_FMA(p, z, 1.0 / 21.0);
    p = QO_FMA(p, z, 1.0 / 19.0);
    p = QO_FMA(p, z, 1.0 / 17.0);
    p = QO_FMA(p, z, 1.0 / 15.0);
    p = QO_FMA(p, z, 1.0 / 13.0);
    p = QO_FMA(p, z, 1.0 / 11.0);
    p = QO_FMA(p, z, 1.0 / 9.0);
    p = QO_FMA(p, z, 1.0 / 7.0);
    p = QO_FMA(p, z, 1.0 / 5.0);
    p = QO_FMA(p, z, 1.0 / 3.0);
    double q = QO_FMA(QO_MUL(s, z), p, s);
    return QO_FMA((double)e, 0.6931471805599453, QO_ADD(q, q));
}

QO_HD double qo_horner6(double x, double c0, double c1, double c2, double c3, double c4, double c5)
{
    double r = c0;
    r = QO_FMA(r, x, c1); r = QO_FMA(r, x, c2); r = QO_FMA(r, x, c3);
    r = QO_FMA(r, x, c4); r = QO_FMA(r, x, c5);
    return r;
}

/* Acklam's rational inverse normal CDF, deterministic */
QO_HD double qo_norminv(double p)
{
    const double plow = 0.02425;
    if (p < plow || p > QO_SUB(1.0, plow)) {
        int upper = p > 0.5;
        double pp = upper ? QO_SUB(1.0, p) : p;
        double q = QO_SQRT(QO_MUL(-2.0, qo_log_det(pp)));
        double num = qo_horner6(q, -7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00);
        double den = 7.784695709041462e-03;
        den = QO_FMA(den, q, 3.224671290700398e-01);
        den = QO_FMA(den, q, 2.445134137142996e+00);
        den = QO_FMA(den, q, 3.754408661907416e+00);
        den = QO_FMA(den, q, 1.0);
        double x = QO_DIV(num, den);
        return upper ? -x : x;
    }
    double q = QO_SUB(p, 0.5), r = QO_MUL(q, q);
    double num = qo_horner6(r, -3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                            1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00);
    double den = qo_horner6(r, -5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                            6.680131188771972e+01, -1.328068155288572e+01, 1.0);
    return QO_DIV(QO_MUL(num, q), den);
}

/* x in [-1, 1] from the 53 stream bits of one variable */
QO_HD double qo_stream_from_bits53(uint64_t k, int dist)
{
    if (dist == 1) {
        double p = QO_FMA(QO_U2D(k), 0x1p-53, 0x1p-54);
        double z = qo_norminv(p);
        if (z > 3.0) z = 3.0;
        if (z < -3.0) z = -3.0;
        return QO_DIV(z, 3.0);
    }
    return QO_FMA(2.0, QO_MUL(QO_U2D(k), 0x1p-53), -1.0);
}

/* x in [-1, 1] for random variable `var` of sample `sample` */
QO_HD double qo_stream_variate(uint64_t seed, uint64_t sample, uint32_t var, int dist)
{
    return qo_stream_from_bits53(qo_stream_bits53(seed, sample, var), dist);
}

/* perturbed parameter value: REL nominal*fma(tol,x,1) ; ABS fma(tol,x,nominal) */
QO_HD double qo_stream_apply(double nominal, double tol, double x, int mode_abs)
{
    return mode_abs ? QO_FMA(tol, x, nominal) : QO_MUL(nominal, QO_FMA(tol, x, 1.0));
}
#endif
