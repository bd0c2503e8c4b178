// smc_fastdiv.cuh -- IEEE-exact float32 division by a divisor that is shared by many quotients.
//
// The hot paths divide many values by the same number (every channel of d / n and filmD / n in the moment update,
// estimator.h:162-226; m2 / (n-1), m3 / n, t^2 m2 / (n (n-1)) in the prepass, stat_denoiser.cu:114-123, :205).  A correctly
// rounded division costs ~10 issue slots and an XU-pipe op each; with y = RN(1 / b) computed once,
//     q0 = RN(x * y);   e = x - b * q0  (exact, one FMA);   q = RN(q0 + e * y)
// is the correctly rounded quotient RN(x / b) whenever
//   (1) the two low significand bits of b are zero (b = B * 2^k with B < 2^22): the exact quotient then stays at least
//       1/(2B) > 1.1e-7 ulp away from every rounding boundary, while the error of q0 + e*y against x/b is below
//       |e/b| * 2^-24 <= 1.5 * 2^-24 = 9e-8 ulp (q0 is within 1.5 ulp of x/b), and a quotient never lies exactly on a
//       boundary; integer sample counts n < 2^22 and their products n (n-1) < 2^24 always qualify;
//   (2) nothing under- or overflows: 2^-60 <= |x| <= 2^60 or x == 0, and 2^-20 <= b <= 2^26.
// Everything else (and only that) takes the IEEE division, so results are bit-identical to `x / b` for every input;
// tests/test_moments_gpu.py and tests/test_denoiser_gpu.py compare the planes bit for bit with CPU division.
#pragma once
#include <cuda_runtime.h>

struct SmcDivisor {
    float b, y;
    bool fast;
};

__device__ __forceinline__ SmcDivisor smc_divisor(float b) {
    SmcDivisor d;
    d.b = b;
    d.y = __frcp_rn(b);
    d.fast = ((__float_as_uint(b) & 3u) == 0u) && b >= 9.5367432e-7f && b <= 67108864.f;
    return d;
}

__device__ __forceinline__ float smc_div(float x, const SmcDivisor &d) {
    const float ax = fabsf(x);
    const bool in_range = (ax >= 8.6736174e-19f && ax <= 1.1529215e18f) || ax == 0.f;
    if (__builtin_expect(!(d.fast && in_range), 0)) return __fdiv_rn(x, d.b);
    const float q0 = __fmul_rn(x, d.y);
    const float e = __fmaf_rn(-d.b, q0, x);
    return __fmaf_rn(e, d.y, q0);
}
