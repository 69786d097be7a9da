// Host twin of log10_ge1 (jaero_b200/csrc/cfe.cu) against glibc log10 on 5 M arguments in [1, 1e40]: prints the largest distance in ulp (2).
// gcc -O2 -ffp-contract=off -o log10_test log10_test.c -lm
#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static double log10_ge1(double x)
{
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    const double ivln10 = 4.34294481903251816668e-01, log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
    uint64_t u; memcpy(&u, &x, 8);
    int hi = (int)(u >> 32); unsigned lo = (unsigned)u;
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    int i = (hi + 0x95f64) & 0x100000;
    hi |= (i ^ 0x3ff00000);
    k += (i >> 20);
    u = ((uint64_t)(unsigned)hi << 32) | lo; double m; memcpy(&m, &u, 8);
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s, w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    double lg = f - (hfsq - s * (hfsq + R));
    double dk = (double)k;
    return (dk * log10_2lo + ivln10 * lg) + dk * log10_2hi;
}
int main(){
    double maxulp=0; srand(1);
    for (int n=0;n<5000000;n++){
        double e = (rand()/(double)RAND_MAX)*40.0;     // 1 .. 1e40
        double x = pow(10.0, e) * (1.0 + rand()/(double)RAND_MAX*1e-3);
        if (n < 1000) x = 1.0 + n * 1e-6;
        double a = log10_ge1(x), b = log10(x);
        double ulp = b != 0 ? fabs(a-b)/ (nextafter(fabs(b), INFINITY)-fabs(b)) : fabs(a);
        if (ulp>maxulp){maxulp=ulp; if (ulp>2) printf("x=%.17g a=%.17g b=%.17g ulp=%g\n",x,a,b,ulp);}
    }
    printf("max ulp %g\n", maxulp); printf("%g %g\n", log10_ge1(1.0), log10_ge1(2.0)-log10(2.0));
    return 0;
}
