// smc_filter_fixup.cu -- non-finite values in the filtered plane, after any of the filter kernels.
//
// The reference's loop `continue`s over rejected taps and over taps outside the disc (stat_denoiser.cu:247-268): a NaN or
// +-Inf radiance reaches exactly the centres it is a member tap of (and its own pixel through the centre tap).  The streaming
// kernels give excluded taps weight 0 instead of branching, and 0 * Inf = NaN would reach every centre of the window.  So the
// prepass stores 0 for such a value and lists the position (SmcNfEntry, smc_prepass.cu); this kernel then walks the few
// listed positions and adds  w * value  to the outputs of the centres whose reference loop would have added it:
//     out = (finite sums) / den  +  w * value      ==      (finite sums + w * value) / den      for a non-finite value
// (+-Inf keeps its sign since den > 0; NaN, Inf - Inf and 0 * Inf give NaN either way).  Rare path: plain global loads, one
// block per listed position, nothing to do -- one read of the counters -- on a clean frame.
#include <algorithm>
#include <cmath>

#include "smc_filter_math.cuh"
#include "smc_internal.h"

namespace {

__device__ __forceinline__ void load_rec16(const unsigned char *row, int pcol, float (&r)[SMC_REC_FLOATS]) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float4 v = __ldg((const float4 *)(row + smc_rec_chunk_offset(pcol, c)));
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
}

// membership of the tap record `q` in the confidence test of the centre record `c`, statistic channel ch (slots: mean 0 / 1 / 4,
// discriminator or standard error 2 / 3 / 5): is_not_discriminated (stat_denoiser.cu:81-88) or moon_mean_test (:132-143)
__device__ __forceinline__ bool member_channel(const float (&c)[SMC_REC_FLOATS], const float (&q)[SMC_REC_FLOATS], int ch, int mode) {
    const int ms = ch == 2 ? 4 : ch, ds = ch == 2 ? 5 : 2 + ch;
    if (mode == SMC_MEMBER_WELCH) return __fadd_rn(c[ds], q[ds]) <= __fmul_rn(__fmul_rn(2.f, c[ms]), q[ms]);
    return q[ms] >= __fsub_rn(c[ms], c[ds]) && q[ms] <= __fadd_rn(c[ms], c[ds]);
}

__global__ void __launch_bounds__(256) nonfinite_fixup_kernel(const SmcFilterParams p, const SmcNfSources src) {
    int n[3], total = 0;
#pragma unroll
    for (int s = 0; s < 3; s++) {
        n[s] = src.list[s] ? min(*(volatile const int *)&src.list[s]->count, SMC_NF_CAP) : 0;
        total += n[s];
    }
    const int r = p.radius;
    const size_t row_bytes = smc_rec_row_bytes(p.rec_pitch);
    const int side = 2 * r;  // centres (y, x) with the listed position at (dy, dx) in [-r, r): y in (qy - r, qy + r]
    for (int i = blockIdx.x; i < total; i += gridDim.x) {
        // entry i of the concatenated lists (constant indices: the arrays live in registers / the parameter bank)
        const int s = i < n[0] ? 0 : i < n[0] + n[1] ? 1 : 2;
        const int k = s == 0 ? i : s == 1 ? i - n[0] : i - n[0] - n[1];
        const SmcNfList *lst = s == 0 ? src.list[0] : s == 1 ? src.list[1] : src.list[2];
        const int lo = s == 0 ? src.row_lo[0] : s == 1 ? src.row_lo[1] : src.row_lo[2];
        const int hi = s == 0 ? src.row_hi[0] : s == 1 ? src.row_hi[1] : src.row_hi[2];
        const int shift = s == 0 ? src.row_shift[0] : s == 1 ? src.row_shift[1] : src.row_shift[2];
        const SmcNfEntry e = lst->e[k];
        if (e.pr < lo || e.pr >= hi) continue;
        const int pr = e.pr + shift;
        const int qy = pr - r, qx = e.pc - p.padX;
        const unsigned char *img = p.rec + (size_t)e.z * p.rec_image_stride;
        float q[SMC_REC_FLOATS];
        load_rec16(img + (size_t)pr * row_bytes, e.pc, q);
        for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
            const int y = qy - r + 1 + t / side, x = qx - r + 1 + t % side;
            if (y < p.row_begin || y >= p.row_end || x < 0 || x >= p.W) continue;
            const int dy = qy - y, dx = qx - x;
            const float sw = __ldg(p.sw + (size_t)(dy + r + p.sw_margin_y) * p.sw_stride + (r + p.sw_margin_x) + dx);
            if (sw == -INFINITY) continue;  // dS2 > rad2 (stat_denoiser.cu:36)
            float c[SMC_REC_FLOATS];
            load_rec16(img + (size_t)(y + r) * row_bytes, x + p.padX, c);
            const bool centre = dy == 0 && dx == 0;
            float w = 1.f;  // is_center (stat_denoiser.cu:78)
            if (!centre) {
                float a = 0.f;
#pragma unroll
                for (int g = 0; g < SMC_REC_GBUF_CHANNELS; g++) {
                    constexpr int slots[SMC_REC_GBUF_CHANNELS] = {10, 11, 12, 13, 14, 15, 7};
                    const float d = __fsub_rn(q[slots[g]], c[slots[g]]);
                    if (g < p.NG) a = __fmaf_rn(d, d, a);
                }
                if (p.NGX > 0) {
                    const size_t gx_row = (size_t)p.rec_pitch * p.gext_stride;
                    const float *gq = p.gext + (size_t)pr * gx_row + (size_t)e.pc * p.gext_stride;
                    const float *gc = p.gext + (size_t)(y + r) * gx_row + (size_t)(x + p.padX) * p.gext_stride;
                    for (int g = 0; g < p.NGX; g++) {
                        const float d = __fsub_rn(__ldg(gq + g), __ldg(gc + g));
                        a = __fmaf_rn(d, d, a);
                    }
                }
                w = exp2f(__fsub_rn(sw, a));  // like the reference's expf: denormal weights are kept (0 * Inf would be NaN)
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!(e.mask >> j & 1)) continue;
                // which test gates value j: RGB statistics -- all three channels; one scalar image -- its channel 0 (the scalar
                // value j = 3 and, for image 0 with denoiseFilm, the RGB film j = 0..2); a triple -- image j's own channel
                bool ok = centre;
                if (!ok) {
                    if (p.C == 3) ok = member_channel(c, q, 0, p.mode) && member_channel(c, q, 1, p.mode) && member_channel(c, q, 2, p.mode);
                    else if (p.tri) ok = member_channel(c, q, j < 3 ? j : 0, p.mode);
                    else ok = member_channel(c, q, 0, p.mode);
                }
                if (!ok) continue;
                const float add = __fmul_rn(w, e.v[j]);
                const bool film_out = p.denoise_film && e.z == 0;
                if (p.C == 3) {
                    const SmcPtrStepSz o = film_out ? p.film_filtered : p.out_ptrs[e.z];
                    atomicAdd((float *)(o.data + (size_t)y * o.step) + x * 3 + j, add);
                } else if (p.tri) {
                    if (3 * e.z + j < p.images) {
                        const SmcPtrStepSz o = p.out_ptrs[3 * e.z + j];
                        atomicAdd((float *)(o.data + (size_t)y * o.step) + x, add);
                    }
                } else if (j == 3) {
                    const SmcPtrStepSz o = p.out_ptrs[e.z];
                    atomicAdd((float *)(o.data + (size_t)y * o.step) + x, add);
                } else if (film_out) {
                    atomicAdd((float *)(p.film_filtered.data + (size_t)y * p.film_filtered.step) + x * 3 + j, add);
                }
            }
        }
    }
    // multi-GPU: the halo rows are no longer read -- the last block to retire tells the neighbours (SmcHaloSync)
    if (p.halo.signal0 || p.halo.signal1) {
        __syncthreads();
        if (threadIdx.x == 0) smc_halo_signal_last(p.halo, gridDim.x);
    }
}

}  // namespace

int smc_launch_nonfinite_fixup(smc_context *ctx, const SmcFilterParams &p, const SmcNfSources &src) {
    if (p.row_end <= p.row_begin) return SMC_OK;
    nonfinite_fixup_kernel<<<2 * std::max(ctx->sm_count, 1), 256, 0, ctx->stream>>>(p, src);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}
