/* fastdiv_check.c -- host restatement of statmc_b200/csrc/smc_fastdiv.cuh (the three-operation shared-divisor division and
 * its range predicate) checked against IEEE division, to back the header's exactness argument numerically on the divisors
 * the hot paths use: n, n - 1 and n (n - 1) for sample counts n < 2^22.  Test infrastructure (tests/test_fastdiv_cpu.py). */
#include <math.h>
#include <stdint.h>
#include <string.h>

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

typedef struct { float b, y; int fast; } divisor;

static divisor make_divisor(float b) {
    divisor d;
    d.b = b;
    d.y = 1.0f / b; /* __frcp_rn: correctly rounded reciprocal */
    d.fast = ((f2u(b) & 3u) == 0u) && b >= 9.5367432e-7f && b <= 67108864.f;
    return d;
}

static float div_fast(float x, const divisor *d, int *took_fast) {
    const float ax = fabsf(x);
    const int in_range = (ax >= 8.6736174e-19f && ax <= 1.1529215e18f) || ax == 0.f;
    *took_fast = d->fast && in_range;
    if (!*took_fast) return x / d->b;
    const float q0 = x * d->y;
    const float e = fmaf(-d->b, q0, x);
    return fmaf(e, d->y, q0);
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 16);
}

/* returns the number of mismatches against IEEE division; *fast_taken = how many quotients took the fast path */
long long fastdiv_check(long long n_divisors, int x_per_divisor, long long *fast_taken) {
    long long bad = 0, fast = 0;
    for (long long i = 0; i < n_divisors; i++) {
        /* sample counts spread over [2, 2^22): small ones densely, large ones at random */
        const uint32_t n = i < 70000 ? (uint32_t)(i + 2) : 2u + rnd() % 4194302u;
        const float cands[3] = {(float)n, (float)n - 1.0f, (float)n * ((float)n - 1.0f)};
        for (int c = 0; c < 3; c++) {
            const divisor d = make_divisor(cands[c]);
            for (int k = 0; k < x_per_divisor; k++) {
                float x;
                if (k == 0) x = 0.f;
                else if (k == 1) x = cands[c];                          /* quotient exactly 1 */
                else if (k == 2) x = u2f(f2u(cands[c]) + 1u);           /* just above 1 */
                else {
                    /* random sign and significand, exponent in [-70, 70]: straddles the [2^-60, 2^60] fast range */
                    const int ex = (int)(rnd() % 141u) - 70;
                    x = ldexpf(1.0f + (float)(rnd() & 0x7fffffu) / 8388608.0f, ex) * ((rnd() & 1u) ? -1.f : 1.f);
                }
                int tf;
                const float q = div_fast(x, &d, &tf);
                fast += tf;
                const float r = x / d.b;
                if (f2u(q) != f2u(r)) bad++;
            }
        }
    }
    *fast_taken = fast;
    return bad;
}
