/* markstein_check.c -- CPU brute force behind div_by<T, true> (lineprofiles_b200/csrc/klp_kernels.cuh).
 *
 * Claim: for a divisor b whose correctly rounded reciprocal y = RN(1/b) is known,
 *     q0 = RN(a * y);  r = fma(-b, q0, a);  q = fma(r, y, q0)
 * is RN(a / b) for every a (no overflow / underflow) -- ONE residual correction, three operations.
 * The proof sketch is in the kernel file; the corner its bounds leave open (a and b within a few ulps of 2 or 1) and
 * a large random sample are checked here against the hardware division, in f32 and f64.
 *
 *   gcc -O2 -ffp-contract=off tools/markstein_check.c -lm -o /tmp/markstein_check
 *   /tmp/markstein_check quick     # ~10 s: what tests/test_abi.py runs
 *   /tmp/markstein_check full      # ~2 min: 497 divisors x all 2^23 f32 mantissas, 36e6 f64 corner pairs, 1.4e9 random pairs
 * Prints one line "bad=<n> checked=<n>" and exits 0 iff bad == 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float mkf(uint32_t m, int e) { uint32_t u = ((uint32_t)(127 + e) << 23) | (m & 0x7fffffu); float f; memcpy(&f, &u, 4); return f; }
static double mkd(uint64_t m, int e) { uint64_t u = ((uint64_t)(1023 + e) << 52) | (m & 0xfffffffffffffULL); double f; memcpy(&f, &u, 8); return f; }
static inline int bad_f(float a, float b) {
    const float y = 1.0f / b, q0 = a * y, r = fmaf(-b, q0, a);
    return fmaf(r, y, q0) != a / b;
}
static inline int bad_d(double a, double b) {
    const double y = 1.0 / b, q0 = a * y, r = fma(-b, q0, a);
    return fma(r, y, q0) != a / b;
}
static uint64_t rng_state = 88172645463325252ULL;
static inline uint64_t rng(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main(int argc, char** argv) {
    const int full = argc > 1 && strcmp(argv[1], "full") == 0;
    long bad = 0, n = 0;
    /* f32: every mantissa of a against chosen divisors (quotients in (1/2, 2): both binade cases) */
    uint32_t bm[1024];
    int nb = 0;
    const int corner = full ? 64 : 12;
    for (int k = 0; k < corner; ++k) bm[nb++] = 0x7fffffu - (uint32_t)k;   /* just below 2: the corner of the proof */
    for (int k = 0; k < corner; ++k) bm[nb++] = (uint32_t)k;                /* just above 1 */
    if (full)
        for (int k = 0; k < 23; ++k) { bm[nb++] = 1u << k; bm[nb++] = (1u << k) - 1; bm[nb++] = 0x7fffffu ^ (1u << k); }
    for (int k = 0; k < (full ? 300 : 24); ++k) bm[nb++] = (uint32_t)rng() & 0x7fffffu;
    for (int ib = 0; ib < nb; ++ib) {
        const float b = mkf(bm[ib], 0);
        for (uint32_t m = 0; m < (1u << 23); ++m) { bad += bad_f(mkf(m, 0), b); ++n; }
    }
    /* f64: corner pairs */
    const int cd = full ? 3000 : 600;
    for (int kb = 0; kb < cd; ++kb)
        for (int ka = 0; ka < cd; ++ka) {
            const uint64_t top = 0xfffffffffffffULL;
            bad += bad_d(mkd(top - ka, 0), mkd(top - kb, 0));
            bad += bad_d(mkd(ka, 0), mkd(kb, 0));
            bad += bad_d(mkd(ka, 0), mkd(top - kb, 0));
            bad += bad_d(mkd(top - ka, 0), mkd(kb, 0));
            n += 4;
        }
    /* random pairs, both precisions, a over two binades */
    const long nr = full ? 700000000L : 40000000L;
    for (long it = 0; it < nr; ++it) {
        const uint64_t s1 = rng(), s2 = s1 * 0x9E3779B97F4A7C15ULL;
        bad += bad_f(mkf((uint32_t)s1, (int)(s1 >> 60 & 1) - 1), mkf((uint32_t)(s1 >> 24), 0));
        bad += bad_d(mkd(s1, (int)(s2 & 1) - 1), mkd(s2 >> 8, 0));
        n += 2;
    }
    printf("bad=%ld checked=%ld\n", bad, n);
    return bad != 0;
}
