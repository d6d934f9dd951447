/* Exhaustive accuracy check of ycge_expf (binary32 FMA form) against the binary64 evaluation ydm_exp_core rounded once:
 *   gcc -O2 -mfma -ffp-contract=off -fopenmp tools/check_expf.c -o /tmp/check_expf -lm && /tmp/check_expf
 * Walks all 2^32 bit patterns; prints the largest distance in ulps, how many inputs differ, and monotonicity breaks; also
 * that ycge_expf_nonpos returns the bits of ycge_expf for every x <= 0 and every NaN. */
#include <math.h>
#include <stdio.h>
#include "../include/ycge_detmath.h"
static unsigned ord(float f) { unsigned b; memcpy(&b, &f, 4); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
int main(void) {
    unsigned long long differ = 0, total = 0, worst = 0, nonmono = 0, nonpos_bad = 0;
    float worst_x = 0;
#pragma omp parallel for reduction(+ : differ, total, nonmono, nonpos_bad) schedule(static)
    for (long long hi = 0; hi < 65536; hi++) {
        unsigned long long lw = 0; float lx = 0;
        float prev = -1.0f;
        for (unsigned lo = 0; lo < 65536; lo++) {
            unsigned b = ((unsigned)hi << 16) | lo;
            float x; memcpy(&x, &b, 4);
            if (x != x || x <= 0.0f) { float u = ycge_expf(x), v = ycge_expf_nonpos(x); if (memcmp(&u, &v, 4)) nonpos_bad++; }
            if (x != x) continue;
            float a = ycge_expf(x);
            float e = (float)ydm_exp_core((double)x);
            if (isinf(x)) e = x > 0 ? x : 0.0f;
            total++;
            unsigned d = ord(a) > ord(e) ? ord(a) - ord(e) : ord(e) - ord(a);
            if (d) differ++;
            if (d > lw) { lw = d; lx = x; }
            /* within one 65536-block of positive floats x increases with the bit pattern (decreases for negative ones) */
            if (lo && prev >= 0.0f && ((b & 0x80000000u) ? a > prev : a < prev)) nonmono++;
            prev = a;
        }
#pragma omp critical
        if (lw > worst) { worst = lw; worst_x = lx; }
    }
    printf("inputs %llu, differ from the rounded binary64 evaluation: %llu (%.4f %%), worst %llu ulp at x = %.9g, monotonicity breaks %llu\n",
           total, differ, 100.0 * differ / total, worst, worst_x, nonmono);
    printf("ycge_expf_nonpos != ycge_expf on x <= 0 or NaN: %llu inputs\n", nonpos_bad);
    return worst > 1 || nonpos_bad != 0;
}
