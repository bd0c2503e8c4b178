// smc_filter_math.cuh -- per-tap arithmetic shared by the generic and the streaming filter kernels, so that
// both produce bit-identical results.  Restates the loop body of filter_kernel (stat_denoiser.cu:247-268, :316-338):
//   membership  is_not_discriminated (:81-88):  discC + discI <= 2.f * meanC * meanI   for every channel
//   weight      expf(dS2 * dSFactor + dr2(...)) (:261), evaluated as 2^(sw - a) with
//                 sw = dS2 * dSFactor * log2(e)  (host-built table, -inf outside the window), and
//                 a  = sum_k (g'_I,k - g'_C,k)^2  over G-buffer channels pre-scaled by sqrt(-drFactor * log2 e)
// The membership comparison is evaluated exactly as the reference does (two rounded operands, one compare);
// the weight differs from expf() by a few ulp (ex2.approx + pre-scaling), far inside the 1e-4 parity tolerance.
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float smc_ex2(float x) {
#ifdef SMC_EXPERIMENT_NO_MUFU  // timing experiment only (wrong results): what the kernel costs without the SFU op
    return x * 0.001f;
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}

// Two-lane fp32 arithmetic.  Default: Blackwell packed FADD2 / FMUL2 / FFMA2 (one instruction, two IEEE-rounded
// lanes).  -DSMC_SCALAR_MATH=1 builds the same arithmetic from scalar ops (bit-identical results) for A/B timing.
#ifndef SMC_SCALAR_MATH
#define SMC_SCALAR_MATH 0
#endif
__device__ __forceinline__ float2 smc_add2(float2 a, float2 b) {
#if SMC_SCALAR_MATH
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
#else
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
#endif
}
__device__ __forceinline__ float2 smc_mul2(float2 a, float2 b) {
#if SMC_SCALAR_MATH
    return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#else
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
#endif
}
__device__ __forceinline__ float2 smc_fma2(float2 a, float2 b, float2 c) {
#if SMC_SCALAR_MATH
    return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
#else
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
#endif
}

// What a thread keeps in registers about one centre pixel.
template <int C, int NG>
struct SmcCentre {
    float2 t01;   // 2*m.x, 2*m.y   (Welch)   |  unused (Moon)
    float tz;     // 2*m.z
    float2 d01;   // d.x, d.y
    float dz;
    float2 lo01, hi01;  // Moon: m - se, m + se
    float loz, hiz;
    float2 g[3];  // NEGATED pre-scaled G channels (0,1) (2,3) (4,5)
    float go;     // NEGATED odd leftover channel
};

// slots of a record held as four float4 chunks c0..c3 (layout in smc_internal.h)
struct SmcRec {
    float4 c0, c1, c2, c3;
};

template <int NG>
__device__ __forceinline__ float smc_rec_odd_g(const SmcRec &r) {
    // leftover channel when NG is odd: NG=1 -> slot 10, 3 -> slot 12, 5 -> slot 14, 7 -> slot 7
    return NG == 1 ? r.c2.z : NG == 3 ? r.c3.x : NG == 5 ? r.c3.z : r.c1.w;
}

template <int C, int NG, int MODE>
__device__ __forceinline__ void smc_make_centre(const SmcRec &r, SmcCentre<C, NG> &c) {
    const float mx = r.c0.x, my = r.c0.y, mz = r.c1.x, dx = r.c0.z, dy = r.c0.w, dz = r.c1.y;
    if (MODE == 0) {
        c.t01 = make_float2(__fmul_rn(2.f, mx), __fmul_rn(2.f, my));  // 2.f * meanC, exact
        c.tz = __fmul_rn(2.f, mz);
        c.d01 = make_float2(dx, dy);
        c.dz = dz;
    } else {
        // moon_mean_test, stat_denoiser.cu:132-143: meanI >= meanC - se && meanI <= meanC + se
        c.lo01 = make_float2(__fsub_rn(mx, dx), __fsub_rn(my, dy));
        c.hi01 = make_float2(__fadd_rn(mx, dx), __fadd_rn(my, dy));
        c.loz = __fsub_rn(mz, dz);
        c.hiz = __fadd_rn(mz, dz);
    }
    c.g[0] = make_float2(-r.c2.z, -r.c2.w);
    c.g[1] = make_float2(-r.c3.x, -r.c3.y);
    c.g[2] = make_float2(-r.c3.z, -r.c3.w);
    c.go = -smc_rec_odd_g<NG>(r);
}

template <int C, int NG, int MODE>
__device__ __forceinline__ bool smc_member(const SmcCentre<C, NG> &c, const SmcRec &r) {
#ifdef SMC_EXPERIMENT_NO_MEMBER  // timing experiment only (wrong results)
    return r.c0.x > -1e30f;
#endif
    if (MODE == 0) {
        if (C == 3) {
            const float2 s = smc_add2(c.d01, make_float2(r.c0.z, r.c0.w));
            const float2 p = smc_mul2(c.t01, make_float2(r.c0.x, r.c0.y));
            const float sz = __fadd_rn(c.dz, r.c1.y);
            const float pz = __fmul_rn(c.tz, r.c1.x);
            return (s.x <= p.x) & (s.y <= p.y) & (sz <= pz);
        } else {
            return __fadd_rn(c.d01.x, r.c0.z) <= __fmul_rn(c.t01.x, r.c0.x);
        }
    } else {
        if (C == 3)
            return (r.c0.x >= c.lo01.x) & (r.c0.x <= c.hi01.x) & (r.c0.y >= c.lo01.y) & (r.c0.y <= c.hi01.y) &
                   (r.c1.x >= c.loz) & (r.c1.x <= c.hiz);
        else
            return (r.c0.x >= c.lo01.x) & (r.c0.x <= c.hi01.x);
    }
}

// 2^(sw - a): sw from the spatial table, a from the G-buffers
template <int C, int NG>
__device__ __forceinline__ float smc_weight(const SmcCentre<C, NG> &c, const SmcRec &r, float sw) {
    float a = 0.f;
    if (NG >= 2) {
        float2 e = smc_add2(make_float2(r.c2.z, r.c2.w), c.g[0]);
        float2 acc = smc_mul2(e, e);
        if (NG >= 4) {
            e = smc_add2(make_float2(r.c3.x, r.c3.y), c.g[1]);
            acc = smc_fma2(e, e, acc);
        }
        if (NG >= 6) {
            e = smc_add2(make_float2(r.c3.z, r.c3.w), c.g[2]);
            acc = smc_fma2(e, e, acc);
        }
        a = __fadd_rn(acc.x, acc.y);
    }
    if (NG & 1) {
        const float e = __fadd_rn(smc_rec_odd_g<NG>(r), c.go);
        a = __fmaf_rn(e, e, a);
    }
    return smc_ex2(__fsub_rn(sw, a));
}
