// smc_filter_generic.cu -- the any-configuration filter kernel: one thread per output pixel walking its window
// over the packed record array in global memory (L1/L2-served).  Same per-tap arithmetic and summation order as the
// streaming kernel (smc_filter_math.cuh), so the two are bit-identical; this one covers what the streaming kernel
// does not instantiate (scalar statistics, odd G-buffer sets, radius > 64) and serves as its on-device cross-check.
//
// Restates filter_kernel<T> / filter_kernel<float3> (stat_denoiser.cu:208-345): half-open window, dS2 <= r^2 cut
// (via the host-built spatial table: -inf marks excluded offsets), replicated borders (already materialised in the
// padded record array by the prepass), centre tap weight 1, and the denoiseFilm / image-0 routing of :251-253,
// :263-265, :271-273 (scalar: both outputs) and :319-344 (RGB: film only).
#include <cmath>

#include "smc_filter_math.cuh"
#include "smc_internal.h"

namespace {

__device__ __forceinline__ SmcRec load_rec(const unsigned char *row, int pcol) {
    SmcRec r;
    r.c0 = __ldg((const float4 *)(row + smc_rec_chunk_offset(pcol, 0)));
    r.c1 = __ldg((const float4 *)(row + smc_rec_chunk_offset(pcol, 1)));
    r.c2 = __ldg((const float4 *)(row + smc_rec_chunk_offset(pcol, 2)));
    r.c3 = __ldg((const float4 *)(row + smc_rec_chunk_offset(pcol, 3)));
    return r;
}

template <int C, int NG, int MODE>
__global__ void __launch_bounds__(256) filter_generic_kernel(SmcFilterParams p) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = p.row_begin + blockIdx.y * 8 + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= p.W || y >= p.row_end) return;
    const int r = p.radius;
    const unsigned char *img = p.rec + (size_t)z * p.rec_image_stride;
    const size_t row_bytes = smc_rec_row_bytes(p.rec_pitch);
    const bool film_out = p.denoise_film && z == 0;

    const SmcRec rc = load_rec(img + (size_t)(y + r) * row_bytes, x + p.padX);
    SmcCentre<C, NG> c;
    smc_make_centre<C, NG, MODE>(rc, c);

    // G-buffer channels beyond the record's seven (SmcFilterParams::gext)
    constexpr int kMaxX = SMC_MAX_GBUF_CHANNELS - SMC_REC_GBUF_CHANNELS;
    float gxc[kMaxX];
    const size_t gx_row = (size_t)p.rec_pitch * p.gext_stride;
    if (p.NGX > 0) {
        const float *e = p.gext + (size_t)(y + r) * gx_row + (size_t)(x + p.padX) * p.gext_stride;
#pragma unroll
        for (int k = 0; k < kMaxX; k++) gxc[k] = k < p.NGX ? __ldg(e + k) : 0.f;
    }

    float n0 = 0.f, n1 = 0.f, n2 = 0.f, ns = 0.f, den = 0.f;
    int accepted = 0;
    for (int dy = -r; dy < r; dy++) {
        const float *swrow = p.sw + (size_t)(dy + r + p.sw_margin_y) * p.sw_stride + (r + p.sw_margin_x);
        const unsigned char *row = img + (size_t)(y + dy + r) * row_bytes;
        for (int dx = -r; dx < r; dx++) {
            const float sw = __ldg(swrow + dx);
            if (sw == -INFINITY) continue;  // dS2 > rad2 (stat_denoiser.cu:36)
            const SmcRec ri = load_rec(row, x + dx + p.padX);
            float w;
            if (dx == 0 && dy == 0) {
                w = 1.f;  // is_center, stat_denoiser.cu:78, :250-254
            } else {
                if (!smc_member<C, NG, MODE>(c, ri)) continue;
                float swx = sw;
                if (p.NGX > 0) {  // the further channels' squared differences join the exponent
                    const float *e = p.gext + (size_t)(y + dy + r) * gx_row + (size_t)(x + dx + p.padX) * p.gext_stride;
                    float ax = 0.f;
#pragma unroll
                    for (int k = 0; k < kMaxX; k++)
                        if (k < p.NGX) {
                            const float dgk = __fsub_rn(__ldg(e + k), gxc[k]);
                            ax = __fmaf_rn(dgk, dgk, ax);
                        }
                    swx = __fsub_rn(sw, ax);
                }
                w = smc_weight<C, NG>(c, ri, swx);
            }
            if (C == 3 || film_out) {
                n0 = __fmaf_rn(w, ri.c2.x, n0);
                n1 = __fmaf_rn(w, ri.c2.y, n1);
                n2 = __fmaf_rn(w, ri.c1.z, n2);
            }
            if (C == 1) ns = __fmaf_rn(w, ri.c1.x, ns);
            den = __fadd_rn(den, w);
            accepted++;
        }
    }

    if (C == 3) {
        const SmcPtrStepSz o = film_out ? p.film_filtered : p.out_ptrs[z];
        float *op = (float *)(o.data + (size_t)y * o.step) + x * 3;
        op[0] = __fdiv_rn(n0, den);
        op[1] = __fdiv_rn(n1, den);
        op[2] = __fdiv_rn(n2, den);
    } else {
        const SmcPtrStepSz o = p.out_ptrs[z];
        ((float *)(o.data + (size_t)y * o.step))[x] = __fdiv_rn(ns, den);
        if (film_out) {
            float *op = (float *)(p.film_filtered.data + (size_t)y * p.film_filtered.step) + x * 3;
            op[0] = __fdiv_rn(n0, den);
            op[1] = __fdiv_rn(n1, den);
            op[2] = __fdiv_rn(n2, den);
        }
    }
    if (p.accepted && p.accepted[z].data) ((int *)(p.accepted[z].data + (size_t)y * p.accepted[z].step))[x] = accepted;
}

template <int C, int MODE>
void launch_ng(const SmcFilterParams &p, dim3 grid, dim3 block, cudaStream_t s) {
    switch (p.NG) {
        case 0: filter_generic_kernel<C, 0, MODE><<<grid, block, 0, s>>>(p); break;
        case 1: filter_generic_kernel<C, 1, MODE><<<grid, block, 0, s>>>(p); break;
        case 2: filter_generic_kernel<C, 2, MODE><<<grid, block, 0, s>>>(p); break;
        case 3: filter_generic_kernel<C, 3, MODE><<<grid, block, 0, s>>>(p); break;
        case 4: filter_generic_kernel<C, 4, MODE><<<grid, block, 0, s>>>(p); break;
        case 5: filter_generic_kernel<C, 5, MODE><<<grid, block, 0, s>>>(p); break;
        case 6: filter_generic_kernel<C, 6, MODE><<<grid, block, 0, s>>>(p); break;
        default: filter_generic_kernel<C, 7, MODE><<<grid, block, 0, s>>>(p); break;
    }
}

}  // namespace

int smc_launch_filter_generic(smc_context *ctx, const SmcFilterParams &p) {
    const int rows = p.row_end - p.row_begin;
    if (rows <= 0) return SMC_OK;
    const dim3 block(32, 8), grid((p.W + 31) / 32, (rows + 7) / 8, p.ptr_count);
    if (p.C == 3) {
        if (p.mode == SMC_MEMBER_WELCH) launch_ng<3, 0>(p, grid, block, ctx->stream);
        else launch_ng<3, 1>(p, grid, block, ctx->stream);
    } else {
        if (p.mode == SMC_MEMBER_WELCH) launch_ng<1, 0>(p, grid, block, ctx->stream);
        else launch_ng<1, 1>(p, grid, block, ctx->stream);
    }
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}
