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

#include "smc_fastdiv.cuh"
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

// One thread per padded record.  Every global load of the pixel is issued before any arithmetic (the G-buffer set is a
// compile-time unrolled list passed by value), so a thread has ~20 independent loads in flight: the kernel is bound by
// HBM bandwidth, not by load latency or the XU pipe (the twelve IEEE divisions of a pixel share three divisors, see
// smc_fastdiv.cuh).
template <int C, int NG>
__device__ __forceinline__ void prepass_pixel(const SmcPrepassParams &p, const int pc, const int pr, const int z);
template <int NG>
__device__ __forceinline__ void prepass_pixel_triple(const SmcPrepassParams &p, const int pc, const int pr, const int z);
template <int NG>
__device__ __forceinline__ void prepass_finish(const SmcPrepassParams &p, float (&rec)[SMC_REC_FLOATS], const float (&g)[NG > 0 ? NG : 1],
                                               const int pc, const int pr, const int z, const int yy, const int y, const int x);

template <int C, int NG>
__global__ void __launch_bounds__(256) prepass_kernel(SmcPrepassParams p) {
    const int col_blocks = (p.rec_pitch + blockDim.x - 1) / blockDim.x;
    const int pc = (blockIdx.x % col_blocks) * blockDim.x + threadIdx.x;  // padded column
    const int pr = p.pr_begin + blockIdx.x / col_blocks;                   // padded row: y = pr - radius
    const int z = blockIdx.y;
    const int yy = pr - p.radius;
    // multi-GPU: this block's row also goes into a neighbour's halo -> wait until that neighbour has filtered the previous step
    const bool to_up = p.peer_up_halo && yy >= 0 && yy < p.radius && yy < p.H;
    const bool to_dn = p.peer_down_halo && yy >= p.H - p.radius && yy >= 0 && yy < p.H;
    if ((to_up || to_dn) && (p.halo.wait0 || p.halo.wait1)) {
        if (threadIdx.x == 0) {
            SmcHaloSync h = p.halo;
            if (!to_up) h.wait0 = nullptr;
            if (!to_dn) h.wait1 = nullptr;
            smc_halo_wait(h);
        }
        __syncthreads();
    }
    const bool idle = pc >= p.rec_pitch || (yy < 0 && p.skip_top) || (yy >= p.H && p.skip_bottom);
    if (idle && !(to_up || to_dn)) return;
    if (!idle) {
        if (C == 1 && p.triple) prepass_pixel_triple<NG>(p, pc, pr, z);
        else prepass_pixel<C, NG>(p, pc, pr, z);
    }
    if (to_up || to_dn) {  // block-uniform
        __threadfence_system();  // this thread's peer stores before the block's signal
        __syncthreads();
        if (threadIdx.x == 0) smc_halo_signal_last(p.halo, p.halo_blocks);
    }
}

template <int C, int NG>
__device__ __forceinline__ void prepass_pixel(const SmcPrepassParams &p, const int pc, const int pr, const int z) {
    const int yy = pr - p.radius;
    const int y = min(max(yy, 0), p.H - 1);
    const int x = min(max(pc - p.padX, 0), p.W - 1);
    const bool own = (yy == y) && (pc - p.padX == x);  // not a replicated copy: also writes the API-visible planes
    const bool welch = p.mode == SMC_MEMBER_WELCH;
    const bool film_value = p.denoise_film && z == 0;

    // ---- loads -------------------------------------------------------------------------------------------------
    const int n = rowi(p.n[z], y)[x];
    float mean[C], m2[C], m3[C], val[C], film[3], g[NG > 0 ? NG : 1];
    {
        const float *meanp = rowf(p.mean[z], y) + x * C;
        const float *m2p = rowf(p.m2[z], y) + x * C;
#pragma unroll
        for (int c = 0; c < C; c++) {
            mean[c] = meanp[c];
            m2[c] = m2p[c];
            m3[c] = 0.f;
            val[c] = 0.f;
        }
        if (welch) {
            const float *m3p = rowf(p.m3[z], y) + x * C;
#pragma unroll
            for (int c = 0; c < C; c++) m3[c] = m3p[c];
        }
        if (!(C == 3 && film_value)) {
            const float *vp = rowf(p.film_ptrs[z], y) + x * C;
#pragma unroll
            for (int c = 0; c < C; c++) val[c] = vp[c];
        }
        film[0] = film[1] = film[2] = 0.f;
        if (film_value) {
            const float *fp = rowf(p.film, y) + x * 3;
            film[0] = fp[0]; film[1] = fp[1]; film[2] = fp[2];
        }
#pragma unroll
        for (int k = 0; k < NG; k++) g[k] = rowf(p.gbufs[p.g_buf[k]], y)[x * p.g_nch[k] + p.g_ch[k]];
    }
    // t-quantile: stat_denoiser.cu:200-202 (Welch, index 2n-3) / :128 (Moon, index n-2)
    const float t = lut_t(p.lut, welch ? 2 * n - 3 : n - 2);

    // ---- arithmetic ----------------------------------------------------------------------------------------------
    const float nF = __int2float_rn(n);
    const float nm1 = __fsub_rn(nF, 1.f);
    const float nn1 = __fmul_rn(nF, nm1);
    const SmcDivisor d_nn1 = smc_divisor(nn1);
    float m[C], d[C];
    if (welch) {
        const SmcDivisor d_nm1 = smc_divisor(nm1), d_n = smc_divisor(nF);
        const float tt = __fmul_rn(t, t);
#pragma unroll
        for (int c = 0; c < C; c++) {
            const float s2 = smc_div(m2[c], d_nm1);  // stat_denoiser.cu:179
            // johnson_mean_corr, stat_denoiser.cu:114-116: s2 > FLT_EPSILON ? (m3 / nF) / (6.f * s2 * nF) : 0.f
            float corr = 0.f;
            if (s2 > FLT_EPSILON) corr = __fdiv_rn(smc_div(m3[c], d_n), __fmul_rn(__fmul_rn(6.f, s2), nF));
            m[c] = __fadd_rn(mean[c], corr);  // :181
            // :205  mean*mean - t*t*m2 / (nF*(nF-1.f))
            d[c] = __fsub_rn(__fmul_rn(m[c], m[c]), smc_div(__fmul_rn(tt, m2[c]), d_nn1));
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
#pragma unroll
        for (int c = 0; c < C; c++) {
            m[c] = mean[c];
            d[c] = __fmul_rn(t, __fsqrt_rn(smc_div(m2[c], d_nn1)));
        }
    }

    float rec[SMC_REC_FLOATS];
#pragma unroll
    for (int i = 0; i < SMC_REC_FLOATS; i++) rec[i] = 0.f;
    if (C == 3) {
        rec[0] = m[0]; rec[1] = m[1]; rec[4] = m[C - 1];
        rec[2] = d[0]; rec[3] = d[1]; rec[5] = d[C - 1];
        rec[8] = film_value ? film[0] : val[0];
        rec[9] = film_value ? film[1] : val[1];
        rec[6] = film_value ? film[2] : val[C - 1];
    } else {
        rec[0] = m[0];
        rec[2] = d[0];
        rec[4] = val[0];
        rec[8] = film[0]; rec[9] = film[1]; rec[6] = film[2];
    }
    prepass_finish<NG>(p, rec, g, pc, pr, z, yy, y, x);
}

// Scalar statistics, three images per record (SmcPrepassParams::triple): image 3z + k goes where channel k of an RGB image
// would -- Johnson-corrected mean in slots 0 / 1 / 4, discriminator in 2 / 3 / 5, value in 8 / 9 / 6 -- so that the filter
// evaluates the G-buffer weight of a pair once for the three of them (the weight does not depend on the image, only the
// membership does; the reference launches one grid.z slice per image over the same G-buffers, stat_denoiser.cu:422).
template <int NG>
__device__ __forceinline__ void prepass_pixel_triple(const SmcPrepassParams &p, const int pc, const int pr, const int z) {
    const int yy = pr - p.radius;
    const int y = min(max(yy, 0), p.H - 1);
    const int x = min(max(pc - p.padX, 0), p.W - 1);
    const bool own = (yy == y) && (pc - p.padX == x);
    float g[NG > 0 ? NG : 1];
#pragma unroll
    for (int k = 0; k < NG; k++) g[k] = rowf(p.gbufs[p.g_buf[k]], y)[x * p.g_nch[k] + p.g_ch[k]];
    float rec[SMC_REC_FLOATS];
#pragma unroll
    for (int i = 0; i < SMC_REC_FLOATS; i++) rec[i] = 0.f;
    constexpr int ms[3] = {0, 1, 4}, ds[3] = {2, 3, 5}, vs[3] = {8, 9, 6};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int img = min(3 * z + k, p.images - 1);  // a triple beyond the last image repeats it; its results are dropped
        const int n = rowi(p.n[img], y)[x];
        const float mean = rowf(p.mean[img], y)[x], m2 = rowf(p.m2[img], y)[x], m3 = rowf(p.m3[img], y)[x];
        const float val = rowf(p.film_ptrs[img], y)[x];
        const float t = lut_t(p.lut, 2 * n - 3);
        const float nF = __int2float_rn(n), nm1 = __fsub_rn(nF, 1.f), nn1 = __fmul_rn(nF, nm1);
        const float s2 = __fdiv_rn(m2, nm1);  // stat_denoiser.cu:179
        float corr = 0.f;
        if (s2 > FLT_EPSILON) corr = __fdiv_rn(__fdiv_rn(m3, nF), __fmul_rn(__fmul_rn(6.f, s2), nF));  // :114-116
        const float m = __fadd_rn(mean, corr);                                                              // :181
        const float d = __fsub_rn(__fmul_rn(m, m), __fdiv_rn(__fmul_rn(__fmul_rn(t, t), m2), nn1));        // :205
        if (own && 3 * z + k < p.images) {
            if (p.mean_corr && p.mean_corr[img].data) rowf_w(p.mean_corr[img], y)[x] = m;
            if (p.disc && p.disc[img].data) rowf_w(p.disc[img], y)[x] = d;
        }
        rec[ms[k]] = m;
        rec[ds[k]] = d;
        rec[vs[k]] = val;
    }
    prepass_finish<NG>(p, rec, g, pc, pr, z, yy, y, x);
}

template <int NG>
__device__ __forceinline__ void prepass_finish(const SmcPrepassParams &p, float (&rec)[SMC_REC_FLOATS], const float (&g)[NG > 0 ? NG : 1],
                                               const int pc, const int pr, const int z, const int yy, const int y, const int x) {
    // G-buffers, flattened and pre-scaled so that  sum_k (g'_C - g'_I)^2 = -log2(e) * sum_g drFactor_g |g_C - g_I|^2
    // (dr2, stat_denoiser.cu:90-112); the filter then needs one subtraction and one FMA per channel and no factor.
    // g_scale[k] = sqrtf(-drFactor * log2(e)) is computed once on the host.
    constexpr int slots[7] = {10, 11, 12, 13, 14, 15, 7};
#pragma unroll
    for (int k = 0; k < NG; k++) rec[slots[k]] = __fmul_rn(g[k], p.g_scale[k]);
    // slot 7 is the seventh G channel; when it is free it holds 1.0f so that (V.z, 1) is a register pair for the streaming
    // filter's packed accumulation (num.z += w * V.z, den += w * 1 in one FFMA2; smc_filter_stream.cu)
    if (NG < 7) rec[7] = 1.f;

    // non-finite values: 0 in the record, listed for the fix-up pass (SmcNfEntry); when the list is full the value stays and
    // the streaming kernels spread NaN over the whole window of the pixel instead of its member taps only
    {
        constexpr int vslot[4] = {8, 9, 6, 4};
        const int nv = (p.C == 1 && !p.triple) ? 4 : 3;
        int mask = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < nv && (__float_as_uint(rec[vslot[j]]) & 0x7f800000u) == 0x7f800000u) mask |= 1 << j;
        if (mask) {
            const int idx = atomicAdd(&p.nf->count, 1);
            if (idx < SMC_NF_CAP) {
                SmcNfEntry e;
                e.pr = pr; e.pc = pc; e.z = z; e.mask = mask;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    e.v[j] = rec[vslot[j]];
                    if (mask >> j & 1) rec[vslot[j]] = 0.f;
                }
                p.nf->e[idx] = e;
            }
        }
    }

    const size_t row_bytes = smc_rec_row_bytes(p.rec_pitch);
    unsigned char *row = p.rec + (size_t)z * p.rec_image_stride + (size_t)pr * row_bytes;
#pragma unroll
    for (int c = 0; c < 4; c++)
        *(float4 *)(row + smc_rec_chunk_offset(pc, c)) = make_float4(rec[4 * c], rec[4 * c + 1], rec[4 * c + 2], rec[4 * c + 3]);
    // G-buffer channels that do not fit the record (more than seven): side array, read by the generic filter kernel
    if (p.NGX > 0 && z == 0) {
        float *e = p.gext + ((size_t)pr * p.rec_pitch + pc) * p.gext_stride;
        for (int k = 0; k < p.NGX; k++) {
            const int kk = SMC_REC_GBUF_CHANNELS + k;
            e[k] = __fmul_rn(rowf(p.gbufs[p.g_buf[kk]], y)[x * p.g_nch[kk] + p.g_ch[kk]], p.g_scale[kk]);
        }
    }
    // halo exchange fused into the producer: the same record goes to the neighbouring GPUs that need it (peer stores)
    if (yy == y) {
        unsigned char *peer = nullptr;
        if (p.peer_up_halo && yy < p.radius) peer = p.peer_up_halo + (size_t)z * p.peer_up_image_stride + (size_t)yy * row_bytes;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (peer) *(float4 *)(peer + smc_rec_chunk_offset(pc, c)) = make_float4(rec[4 * c], rec[4 * c + 1], rec[4 * c + 2], rec[4 * c + 3]);
        peer = nullptr;
        if (p.peer_down_halo && yy >= p.H - p.radius)
            peer = p.peer_down_halo + (size_t)z * p.peer_down_image_stride + (size_t)(yy - (p.H - p.radius)) * row_bytes;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (peer) *(float4 *)(peer + smc_rec_chunk_offset(pc, c)) = make_float4(rec[4 * c], rec[4 * c + 1], rec[4 * c + 2], rec[4 * c + 3]);
    }
}

// Cross-GPU flags of the halo protocol: a release store at system scope after everything this stream did before, and a
// spinning acquire load (one thread; the waiting GPU has nothing else to do on that stream).
__global__ void halo_signal_kernel(int *f0, int *f1, int value) {
    __threadfence_system();
    if (f0) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f0), "r"(value) : "memory");
    if (f1) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f1), "r"(value) : "memory");
}
__global__ void halo_wait_kernel(const int *f0, const int *f1, int value) {
    for (int k = 0; k < 2; k++) {
        const int *f = k ? f1 : f0;
        if (!f) continue;
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v < value) __nanosleep(200);
        } while (v < value);
    }
}

}  // namespace

template <int C>
static void launch_prepass_ng(const SmcPrepassParams &p, dim3 grid, dim3 block, cudaStream_t s) {
    switch (p.NG) {
        case 0: prepass_kernel<C, 0><<<grid, block, 0, s>>>(p); break;
        case 1: prepass_kernel<C, 1><<<grid, block, 0, s>>>(p); break;
        case 2: prepass_kernel<C, 2><<<grid, block, 0, s>>>(p); break;
        case 3: prepass_kernel<C, 3><<<grid, block, 0, s>>>(p); break;
        case 4: prepass_kernel<C, 4><<<grid, block, 0, s>>>(p); break;
        case 5: prepass_kernel<C, 5><<<grid, block, 0, s>>>(p); break;
        case 6: prepass_kernel<C, 6><<<grid, block, 0, s>>>(p); break;
        default: prepass_kernel<C, 7><<<grid, block, 0, s>>>(p); break;
    }
}

int smc_launch_halo_signal(smc_context *ctx, int *f0, int *f1, int value) {
    if (!f0 && !f1) return SMC_OK;
    halo_signal_kernel<<<1, 1, 0, ctx->stream>>>(f0, f1, value);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}
int smc_launch_halo_wait(smc_context *ctx, const int *f0, const int *f1, int value) {
    if (!f0 && !f1) return SMC_OK;
    halo_wait_kernel<<<1, 1, 0, ctx->stream>>>(f0, f1, value);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}

int smc_launch_prepass(smc_context *ctx, const SmcPrepassParams &p) {
    if (p.pr_end <= p.pr_begin) return SMC_OK;
    const dim3 block(256), grid(((p.rec_pitch + 255) / 256) * (p.pr_end - p.pr_begin), p.ptr_count);
    if (p.C == 3) launch_prepass_ng<3>(p, grid, block, ctx->stream);
    else launch_prepass_ng<1>(p, grid, block, ctx->stream);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}
