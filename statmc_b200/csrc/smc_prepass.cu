// smc_prepass.cu -- fused prepass of the denoiser: Johnson mean correction + Welch discriminator (or Moon CI
// half-width) + packing of everything the filter needs about a pixel into one 64-byte record, written into a
// border-replicated, line-padded record array (layout: smc_internal.h).
//
// Replaces johnson_mean_corrs_kernel (stat_denoiser.cu:162-182, johnson_mean_corr :114-123) and
// mean_discriminators_kernel (:184-206), and moves every clamp-to-edge read of the filter (BrdReplicate,
// SET_INNER stat_denoiser.cu:30-37) out of the inner loop: the padded array already holds the replicated values.
// HBM-bound: reads 4 + 4*12 + 12 + G bytes per pixel once, writes 64 B per (padded) pixel once.
//
// Float semantics: every operation is a single IEEE round-to-nearest op (__f*_rn) in the reference's order; the
// reference's SASS has no fused multiply-add on these values (checked with cuobjdump on oracle/_ref), so
// mean-corr and discriminator planes are bit-identical to the reference kernels' output.
#include <cfloat>

#include "smc_internal.h"

namespace {

__device__ __forceinline__ const float *rowf(const SmcPtrStepSz &p, int y) {
    return (const float *)(p.data + (size_t)y * p.step);
}
__device__ __forceinline__ float *rowf_w(const SmcPtrStepSz &p, int y) { return (float *)(p.data + (size_t)y * p.step); }
__device__ __forceinline__ const int *rowi(const SmcPtrStepSz &p, int y) {
    return (const int *)(p.data + (size_t)y * p.step);
}

__device__ __forceinline__ float lut_t(const float *__restrict__ lut, int idx) {
    // stat_denoiser.cu:200-202 (signed compare).  idx < 0 (n < 2) is an out-of-bounds read in the reference;
    // we read entry 0: for n == 1 the result is NaN anyway (0/0 below), for n <= 0 it is implementation-defined.
    idx = idx < 0 ? 0 : idx;
    return __ldg(lut + (idx < SMC_T_LUT_ENTRIES ? idx : SMC_T_LUT_ENTRIES - 1));
}

template <int C>
__global__ void __launch_bounds__(256) prepass_kernel(SmcPrepassParams p) {
    const int col_blocks = (p.rec_pitch + blockDim.x - 1) / blockDim.x;
    const int pc = (blockIdx.x % col_blocks) * blockDim.x + threadIdx.x;  // padded column
    const int pr = p.pr_begin + blockIdx.x / col_blocks;                   // padded row: y = pr - radius
    const int z = blockIdx.y;
    if (pc >= p.rec_pitch) return;
    const int yy = pr - p.radius;
    if ((yy < 0 && p.skip_top) || (yy >= p.H && p.skip_bottom)) return;
    const int y = min(max(yy, 0), p.H - 1);
    const int x = min(max(pc - p.padX, 0), p.W - 1);
    const bool own = (yy == y) && (pc - p.padX == x);  // not a replicated copy: also writes the API-visible planes

    const int n = rowi(p.n[z], y)[x];
    const float nF = __int2float_rn(n);
    const float nm1 = __fsub_rn(nF, 1.f);
    const float *meanp = rowf(p.mean[z], y) + x * C;
    const float *m2p = rowf(p.m2[z], y) + x * C;

    float rec[SMC_REC_FLOATS];
#pragma unroll
    for (int i = 0; i < SMC_REC_FLOATS; i++) rec[i] = 0.f;

    float m[C], d[C];
    if (p.mode == SMC_MEMBER_WELCH) {
        const float *m3p = rowf(p.m3[z], y) + x * C;
        const float t = lut_t(p.lut, 2 * n - 3);
        const float tt = __fmul_rn(t, t);
        const float nn1 = __fmul_rn(nF, nm1);
#pragma unroll
        for (int c = 0; c < C; c++) {
            const float m2 = m2p[c];
            const float s2 = __fdiv_rn(m2, nm1);  // stat_denoiser.cu:179
            // johnson_mean_corr, stat_denoiser.cu:114-116: s2 > FLT_EPSILON ? (m3 / nF) / (6.f * s2 * nF) : 0.f
            float corr = 0.f;
            if (s2 > FLT_EPSILON) corr = __fdiv_rn(__fdiv_rn(m3p[c], nF), __fmul_rn(__fmul_rn(6.f, s2), nF));
            m[c] = __fadd_rn(meanp[c], corr);  // :181
            // :205  mean*mean - t*t*m2 / (nF*(nF-1.f))
            d[c] = __fsub_rn(__fmul_rn(m[c], m[c]), __fdiv_rn(__fmul_rn(tt, m2), nn1));
        }
        if (own) {
            if (p.mean_corr && p.mean_corr[z].data) {
                float *o = rowf_w(p.mean_corr[z], y) + x * C;
#pragma unroll
                for (int c = 0; c < C; c++) o[c] = m[c];
            }
            if (p.disc && p.disc[z].data) {
                float *o = rowf_w(p.disc[z], y) + x * C;
#pragma unroll
                for (int c = 0; c < C; c++) o[c] = d[c];
            }
        }
    } else {
        // Moon et al. CI test, stat_denoiser.cu:125-131: t index n-2, se = t * sqrtf(m2 / (nF * (nF - 1.f)))
        const float t = lut_t(p.lut, n - 2);
        const float nn1 = __fmul_rn(nF, nm1);
#pragma unroll
        for (int c = 0; c < C; c++) {
            m[c] = meanp[c];
            d[c] = __fmul_rn(t, __fsqrt_rn(__fdiv_rn(m2p[c], nn1)));
        }
    }

    if (C == 3) {
        rec[0] = m[0]; rec[1] = m[1]; rec[4] = m[C - 1];
        rec[2] = d[0]; rec[3] = d[1]; rec[5] = d[C - 1];
        const float *v = (p.denoise_film && z == 0) ? rowf(p.film, y) + x * 3 : rowf(p.film_ptrs[z], y) + x * 3;
        rec[8] = v[0]; rec[9] = v[1]; rec[6] = v[2];
    } else {
        rec[0] = m[0];
        rec[2] = d[0];
        rec[4] = rowf(p.film_ptrs[z], y)[x];
        if (p.denoise_film && z == 0) {
            const float *v = rowf(p.film, y) + x * 3;
            rec[8] = v[0]; rec[9] = v[1]; rec[6] = v[2];
        }
    }

    // G-buffers, flattened and pre-scaled so that  sum_k (g'_C - g'_I)^2 = -log2(e) * sum_g drFactor_g |g_C - g_I|^2
    // (dr2, stat_denoiser.cu:90-112); the filter then needs one subtraction and one FMA per channel and no factor.
    const int slots[7] = {10, 11, 12, 13, 14, 15, 7};
    int k = 0;
    for (int g = 0; g < p.n_gbufs; g++) {
        const int gc = p.gbuf_channels[g];
        const float scale = sqrtf(-p.gbuf_dr_factors[g] * 1.4426950408889634f);
        const float *gp = rowf(p.gbufs[g], y) + x * gc;
        for (int c = 0; c < gc; c++, k++) {
            const float v = __fmul_rn(gp[c], scale);
#pragma unroll
            for (int s = 0; s < 7; s++)
                if (k == s) rec[slots[s]] = v;
        }
    }

    unsigned char *row = p.rec + (size_t)z * p.rec_image_stride + (size_t)pr * smc_rec_row_bytes(p.rec_pitch);
#pragma unroll
    for (int c = 0; c < 4; c++)
        *(float4 *)(row + smc_rec_chunk_offset(pc, c)) = make_float4(rec[4 * c], rec[4 * c + 1], rec[4 * c + 2], rec[4 * c + 3]);
}

}  // namespace

int smc_launch_prepass(smc_context *ctx, const SmcPrepassParams &p) {
    if (p.pr_end <= p.pr_begin) return SMC_OK;
    const dim3 block(256), grid(((p.rec_pitch + 255) / 256) * (p.pr_end - p.pr_begin), p.ptr_count);
    if (p.C == 3) prepass_kernel<3><<<grid, block, 0, ctx->stream>>>(p);
    else prepass_kernel<1><<<grid, block, 0, ctx->stream>>>(p);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}
