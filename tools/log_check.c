#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
static double mylog(double x)
{
    uint64_t u; memcpy(&u, &x, 8);
    int hx = (int)(u >> 32); unsigned lx = (unsigned)u;
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int i = (hx + 0x95f64) & 0x100000;
    u = ((uint64_t)(unsigned)(hx | (i ^ 0x3ff00000)) << 32) | lx; memcpy(&x, &u, 8);
    k += i >> 20;
    double f = x - 1.0;
    double s = f / (2.0 + f);
    double dk = (double)k;
    double z = s * s, w = z * z;
    double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}
int main(void)
{
    double worst = 0; srand(1);
    for (int t = 0; t < 20000000; t++) {
        double e = (rand() / (double)RAND_MAX) * 60.0 - 30.0, m = 1.0 + rand() / (double)RAND_MAX;
        double x = ldexp(m, (int)e);
        if (t % 3 == 0) x = 1.0 + (rand() / (double)RAND_MAX - 0.5) * 1e-3 * (t % 7);
        double a = mylog(x); long double b = logl((long double)x);
        double ulp = fabs(nextafter((double)b, INFINITY) - (double)b);
        if (ulp == 0) continue;
        double err = (double)(fabsl((long double)a - b) / ulp);
        if (err > worst) { worst = err; }
    }
    printf("worst error %.3f ulp\n", worst);
    return 0;
}
