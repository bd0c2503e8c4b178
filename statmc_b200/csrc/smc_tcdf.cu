// smc_tcdf.cu -- float32 Student-t CDF on the device.
//
// The reference never evaluates a t CDF: its test compares against a fixed quantile table (stat_denoiser.cu:56,
// 200-205).  BASELINE.json asks for "a fast, accuracy-checked Student-t CDF"; here it serves to validate quantile
// tables on the device (cdf(table[i], i + 1) == 1 - alpha/2) and as the building block of soft membership.
//
//   P(T <= t) = 1 - 0.5 * I_x(nu/2, 1/2),  x = nu / (nu + t^2)               (t >= 0; symmetric otherwise)
// evaluated with the modified-Lentz continued fraction in float32, with the two float32 hazards removed:
//   * x -> 1 (small t): switch to 1 - I_{1-x}(1/2, nu/2) with 1 - x = t^2 / (nu + t^2) formed directly;
//   * ln B(nu/2, 1/2) for large nu: lgamma differences cancel catastrophically in float32, so the ratio
//     Gamma(a + 1/2) / Gamma(a) comes from its asymptotic series for a >= 8 and from lgammaf below.
// Accuracy (tests/test_tcdf.py, against scipy in float64): |error| <= 5e-6 for nu in [1, 2048].
#include "smc_internal.h"

namespace {

__device__ float betacf(float a, float b, float x) {
    const float tiny = 1e-30f, eps = 3e-7f;
    const float qab = a + b, qap = a + 1.f, qam = a - 1.f;
    float c = 1.f, d = 1.f - qab * x / qap;
    if (fabsf(d) < tiny) d = tiny;
    d = 1.f / d;
    float h = d;
    for (int m = 1; m <= 300; m++) {
        const float m2 = 2.f * m;
        float aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.f + aa * d;
        if (fabsf(d) < tiny) d = tiny;
        c = 1.f + aa / c;
        if (fabsf(c) < tiny) c = tiny;
        d = 1.f / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.f + aa * d;
        if (fabsf(d) < tiny) d = tiny;
        c = 1.f + aa / c;
        if (fabsf(c) < tiny) c = tiny;
        d = 1.f / d;
        const float del = d * c;
        h *= del;
        if (fabsf(del - 1.f) < eps) break;
    }
    return h;
}

// ln( Gamma(a + 1/2) / Gamma(a) )
__device__ float ln_gamma_ratio_half(float a) {
    if (a < 8.f) return lgammaf(a + 0.5f) - lgammaf(a);
    const float ia = 1.f / a, ia2 = ia * ia;
    // 0.5 ln a - 1/(8a) + 1/(192 a^3) - 1/(640 a^5)
    return 0.5f * logf(a) + ia * (-0.125f + ia2 * (1.f / 192.f - ia2 * (1.f / 640.f)));
}

__device__ float t_cdf(float t, float nu) {
    if (!(nu > 0.f) || t != t) return nanf("");
    const float at = fabsf(t);
    if (at == INFINITY) return t > 0 ? 1.f : 0.f;
    const float a = 0.5f * nu, b = 0.5f;
    const float t2 = at * at;
    const float x = nu / (nu + t2);   // -> 1 for small t
    const float y = t2 / (nu + t2);   // = 1 - x, no cancellation
    // ln of x^a (1-x)^b / B(a, b),  B(a, 1/2) = Gamma(a) sqrt(pi) / Gamma(a + 1/2)
    const float lbt = ln_gamma_ratio_half(a) - 0.5723649429247001f + a * log1pf(-y) + b * logf(y);
    const float bt = y > 0.f ? expf(lbt) : 0.f;
    float tail;  // P(T > |t|)
    if (x < (a + 1.f) / (a + b + 2.f))
        tail = 0.5f * bt * betacf(a, b, x) / a;
    else
        tail = 0.5f * (1.f - bt * betacf(b, a, y) / b);
    return t >= 0.f ? 1.f - tail : tail;
}

__global__ void t_cdf_kernel(const float *__restrict__ t, const float *__restrict__ df, float *__restrict__ out,
                             size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = t_cdf(t[i], df[i]);
}

}  // namespace

extern "C" int smc_student_t_cdf(smc_context *ctx, const float *t, const float *df, float *out, size_t count) {
    if (!ctx || !t || !df || !out) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    if (count == 0) return SMC_OK;
    SMC_CUDA(cudaSetDevice(ctx->device));
    const unsigned blocks = (unsigned)((count + 255) / 256);
    t_cdf_kernel<<<blocks, 256, 0, ctx->stream>>>(t, df, out, count);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}
