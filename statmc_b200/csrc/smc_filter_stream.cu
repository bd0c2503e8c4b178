// smc_filter_stream.cu -- the B200 filter kernel: membership-gated cross-bilateral filtering of RGB statistics
// (filter_kernel<float3>, stat_denoiser.cu:276-345) restructured around the SM's FP32 pipes instead of its LSUs.
//
// The reference runs one thread per pixel and re-reads ~15 floats per tap from global/L1 (2r x 2r taps).  Here:
//   * a persistent CTA (4 warps) owns a tile of 256 x PY output pixels; each thread owns 2 x PY of them and keeps
//     their centre statistics and accumulators in registers;
//   * the record rows the tile needs (y0-r .. y0+PY-1+r-1, 256+2r records wide) stream through a ring of shared
//     memory slots filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP) issued by one
//     elected thread, D-1 rows ahead of the consumers; the record array is already border-replicated and
//     line-padded by the prepass, so a slot is a verbatim copy, every LDS.128 is bank-conflict-free and its address
//     is `register + immediate`;
//   * every record a thread loads (4 x LDS.128) is used for its 2 x PY centre pixels, i.e. 0.5 (PY=4) LDS.128 per
//     pair evaluation; per pair the math is ~21 issue slots / ~24 FP32 lane-cycles using packed FADD2/FMUL2/FFMA2.
// Work is compute-bound (1255 pair evaluations per pixel at r = 20 against 64 B of record traffic), so the roofline
// is the FP32 pipe: see DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "smc_filter_math.cuh"
#include "smc_internal.h"

namespace {

#ifndef SMC_PACKED_ACC
#define SMC_PACKED_ACC 1
#endif
constexpr int kTileW = 256;
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
#ifndef SMC_STREAM_MINB
#define SMC_STREAM_MINB 2  // resident CTAs per SM the register allocation is bounded for
#endif

struct StreamGeom {
    int tiles_x, tiles_y;
    int total_tiles;
    int nr;              // record rows per tile = 2r + PY - 1
    int depth;           // ring slots
    int slot_bytes;      // bytes per ring slot (multiple of 128)
    int seg_max_rec;     // records per full segment (even)
    int sw_rows;         // rows of the spatial table
    const int2 *rowrange;  // per table row: {jlo, jhi} (device)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared-memory counter: returns the old value; acq_rel at CTA scope orders this warp's reads of a ring slot before the
// refill another warp (or this one) issues after seeing the count complete
__device__ __forceinline__ uint32_t smem_inc_acq_rel(uint32_t *p) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(p)) : "memory");
    return old;
}

__device__ __forceinline__ SmcRec lds_rec(const unsigned char *p) {
    SmcRec r;
    r.c0 = *(const float4 *)(p);
    r.c1 = *(const float4 *)(p + 16);
    r.c2 = *(const float4 *)(p + 32);
    r.c3 = *(const float4 *)(p + 48);
    return r;
}
__device__ __forceinline__ SmcRec ldg_rec(const unsigned char *row, int pcol) {
    const unsigned char *p = row + smc_rec_offset(pcol);
    SmcRec r;
    r.c0 = __ldg((const float4 *)(p));
    r.c1 = __ldg((const float4 *)(p + 16));
    r.c2 = __ldg((const float4 *)(p + 32));
    r.c3 = __ldg((const float4 *)(p + 48));
    return r;
}

// (n0, n1) and (n2, den) are 64-bit register pairs: with NG <= 6 record slot 7 holds 1.0f, so (V.z, 1) sits next to each
// other like (V.x, V.y) and a tap can be accumulated with two FFMA2 instead of three FFMA + one FADD (den + w * 1 rounds
// exactly like den + w, so the result does not change).  Measured: 8.70 -> 8.38 ms for the scalar-statistics kernel that
// also filters the film (five sums per tap), 9.90 -> 10.02 ms for the RGB kernel -- so only the former uses it.
struct Acc {
    float2 n01, n2d;
    float ns;  // scalar statistics only: the scalar value's sum
    int cnt;
};

template <int NG, int MODE, bool COUNT>
__device__ __forceinline__ void pair_eval(const SmcCentre<3, NG> &c, const SmcRec &r, float sw, Acc &a) {
    const bool ok = smc_member<3, NG, MODE>(c, r);
    const float w = smc_weight<3, NG>(c, r, sw);
    if (ok) {
        a.n01.x = __fmaf_rn(w, r.c2.x, a.n01.x);
        a.n01.y = __fmaf_rn(w, r.c2.y, a.n01.y);
        a.n2d.x = __fmaf_rn(w, r.c1.z, a.n2d.x);
        a.n2d.y = __fadd_rn(a.n2d.y, w);
        // a tap outside the window has sw = -inf -> w = 0: it adds nothing, but must not be counted
        if (COUNT) a.cnt += (sw != -INFINITY) ? 1 : 0;
    }
}

// Per-warp kernel: C = channels of the statistics (1 or 3).  C == 3 averages the RGB value (slots 8, 9, 6).  C == 1 averages the
// scalar value (slot 4) and, for the image that also filters the film (denoiseFilm, z == 0; stat_denoiser.cu:251-253,
// :263-265), the film RGB as well (FILM).
template <int C, int NG, int MODE, bool COUNT, bool FILM>
__device__ __forceinline__ void pair_eval_c(const SmcCentre<C, NG> &c, const SmcRec &r, float sw, Acc &a) {
    const bool ok = smc_member<C, NG, MODE>(c, r);
    const float w = smc_weight<C, NG>(c, r, sw);
    if (ok) {
        if (C == 1 && FILM && NG <= 6 && SMC_PACKED_ACC) {
            const float2 ww = make_float2(w, w);
            a.n01 = smc_fma2(ww, make_float2(r.c2.x, r.c2.y), a.n01);
            a.n2d = smc_fma2(ww, make_float2(r.c1.z, r.c1.w), a.n2d);  // slot 7 == 1.0f
        } else {
            if (C == 3 || FILM) {
                a.n01.x = __fmaf_rn(w, r.c2.x, a.n01.x);
                a.n01.y = __fmaf_rn(w, r.c2.y, a.n01.y);
                a.n2d.x = __fmaf_rn(w, r.c1.z, a.n2d.x);
            }
            a.n2d.y = __fadd_rn(a.n2d.y, w);
        }
        if (C == 1) a.ns = __fmaf_rn(w, r.c1.x, a.ns);
        if (COUNT) a.cnt += (sw != -INFINITY) ? 1 : 0;
    }
}

struct TileCoord {
    int z, x0, y0;
};

__device__ __forceinline__ TileCoord tile_coord(int t, const StreamGeom &g, int row_begin, int PY) {
    TileCoord c;
    const int per_img = g.tiles_x * g.tiles_y;
    c.z = t / per_img;
    const int rem = t - c.z * per_img;
    const int ty = rem / g.tiles_x;
    c.x0 = (rem - ty * g.tiles_x) * kTileW;
    c.y0 = row_begin + ty * PY;
    return c;
}

// first record (even) of the row segment a tile stages
__device__ __forceinline__ int seg_start_of(const SmcFilterParams &p, int x0) { return (x0 + p.padX - p.radius) & ~1; }

// source address / size of record row `i` (0..nr-1) of a tile; record row k of the array holds y = k - r
__device__ __forceinline__ void seg_of(const SmcFilterParams &p, const TileCoord &tc, int i, const unsigned char *&src,
                                       uint32_t &bytes, int seg_max_rec) {
    const int seg_start = seg_start_of(p, tc.x0);
    const int nrec = min(seg_max_rec, p.rec_pitch - seg_start);  // even
    src = p.rec + (size_t)tc.z * p.rec_image_stride + (size_t)(tc.y0 + i) * smc_rec_row_bytes(p.rec_pitch) +
          smc_rec_offset(seg_start);
    bytes = (uint32_t)(nrec / 2) * SMC_LINE_BYTES;
}

template <int NG, int PY, int MODE, bool COUNT>
__global__ void __launch_bounds__(kThreads, SMC_STREAM_MINB) filter_stream_kernel(const SmcFilterParams p, const StreamGeom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [ring: depth * slot_bytes][sw table: sw_rows * sw_stride floats][rowrange: sw_rows int2][barriers]
    unsigned char *ring = smem;
    float *sw = (float *)(ring + (size_t)g.depth * g.slot_bytes);
    int2 *rowrange = (int2 *)(sw + g.sw_rows * p.sw_stride);
    uint64_t *full = (uint64_t *)(rowrange + g.sw_rows);
    uint32_t *left = (uint32_t *)(full + 8);              // warps that have left each slot
    volatile int *tile_ids = (volatile int *)(left + 8);  // [4]: tiles of this CTA's sequence, indexed by (sequence number & 3)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = p.radius;
    const size_t row_bytes = smc_rec_row_bytes(p.rec_pitch);

    for (int i = threadIdx.x; i < g.sw_rows * p.sw_stride; i += kThreads) sw[i] = p.sw[i];
    for (int i = threadIdx.x; i < g.sw_rows; i += kThreads) rowrange[i] = g.rowrange[i];
    // Dynamic tile scheduling: the first tile is the CTA's index, every further one comes from a global counter, fetched
    // two tiles ahead so that the row stream never stalls at a tile boundary.  (With a static round-robin assignment the
    // CTAs sharing an SM finish far apart -- the warp scheduler favours one of them -- and the SM idles through a long
    // tail: measured 7.2 / 8.9 / 11.4 ms for the three CTAs of an SM at 4K.)
    auto next_tile = [&]() { return (int)gridDim.x + atomicAdd(p.tile_counter, 1); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < g.depth; s++) {
            mbar_init(&full[s], 1);
            left[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tile_ids[0] = (int)blockIdx.x;
        tile_ids[1] = next_tile();
    }
    __syncthreads();
    if (tile_ids[0] >= g.total_tiles) return;
    if (p.trace && threadIdx.x == 0) {
        unsigned long long t;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[4 * blockIdx.x + 0] = t;
        p.trace[4 * blockIdx.x + 2] = smid;
    }

    // issue the copy for stream position `q` (row q % nr of the CTA's tile number q / nr); nothing past the last tile
    auto issue = [&](long long q) {
        const int tl = (int)(q / g.nr), i = (int)(q - (long long)tl * g.nr);
        const int t = tile_ids[tl & 3];
        if (t >= g.total_tiles) return;
        const TileCoord tc = tile_coord(t, g, p.row_begin, PY);
        const unsigned char *src;
        uint32_t bytes;
        seg_of(p, tc, i, src, bytes, g.seg_max_rec);
        const int s = (int)(q % g.depth);
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(ring + (size_t)s * g.slot_bytes, src, bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (long long q = 0; q < g.depth; q++) issue(q);

    long long pos = 0;
    int tl = 0;
    for (;; tl++) {
        const int tile = tile_ids[tl & 3];
        if (tile >= g.total_tiles) break;
        // tile tl+1 is already known (the copies of its first rows are issued during this tile); fetch tile tl+2.  Its
        // slot last held tile tl-2, which every warp has left: warps drift by fewer than `depth` <= nr rows.
        if (threadIdx.x == 0) tile_ids[(tl + 2) & 3] = next_tile();
        const TileCoord tc = tile_coord(tile, g, p.row_begin, PY);
        const unsigned char *img = p.rec + (size_t)tc.z * p.rec_image_stride;
        const int xf = tc.x0 + warp * 64 + 2 * lane;  // first of this thread's two columns
        const bool warp_active = tc.x0 + warp * 64 < p.W;
        const int base_idx = xf + p.padX - seg_start_of(p, tc.x0);  // slot index of the record at dx = 0, column kx = 0

        SmcCentre<3, NG> cen[PY][2];
        Acc acc[PY][2];
#pragma unroll
        for (int ky = 0; ky < PY; ky++)
#pragma unroll
            for (int kx = 0; kx < 2; kx++) {
                const int yc = min(tc.y0 + ky, p.H - 1), xc = min(xf + kx, p.W - 1);
                const SmcRec rc = ldg_rec(img + (size_t)(yc + r) * row_bytes, xc + p.padX);
                smc_make_centre<3, NG, MODE>(rc, cen[ky][kx]);
                acc[ky][kx].n01.x = acc[ky][kx].n01.y = acc[ky][kx].n2d.x = acc[ky][kx].n2d.y = 0.f;
                acc[ky][kx].cnt = 0;
            }

        for (int i = 0; i < g.nr; i++, pos++) {
            const int s = (int)(pos % g.depth);
            const uint32_t parity = (uint32_t)((pos / g.depth) & 1);
            mbar_wait(&full[s], parity);
            if (warp_active) {
                // table row of centre row ky: dy = i - r - ky  ->  row index dy + r + margin_y = i - ky + margin_y
                int lo = 1 << 20, hi = -(1 << 20);
#pragma unroll
                for (int ky = 0; ky < PY; ky++) {
                    const int2 rr = rowrange[i - ky + p.sw_margin_y];
                    lo = min(lo, rr.x);
                    hi = max(hi, rr.y);
                }
                if (lo <= hi) {
                    // pointer to sw[row of ky = 0][dx = lo]; the rows of ky = 1.. are sw_stride floats lower each
                    const float *swp = sw + (i + p.sw_margin_y) * p.sw_stride + (r + p.sw_margin_x) + lo;
                    const int sws = p.sw_stride;
                    float sw_prev[PY];
#pragma unroll
                    for (int ky = 0; ky < PY; ky++) sw_prev[ky] = swp[-ky * sws - 1];
                    // records base_idx+lo, +1, ...: byte offsets alternate +64 / +80 (two records + pad per 144-B line)
                    const int first = base_idx + lo;
                    const unsigned char *rp = ring + (size_t)s * g.slot_bytes + smc_rec_offset(first);
                    const int d0 = (first & 1) ? SMC_LINE_BYTES - SMC_REC_BYTES : SMC_REC_BYTES;
                    const int d1 = SMC_LINE_BYTES - d0;
                    SmcRec cur = lds_rec(rp);
                    int j = lo;
                    for (; j + 1 <= hi; j += 2) {
                        const SmcRec nxt = lds_rec(rp + d0);
#pragma unroll
                        for (int ky = 0; ky < PY; ky++) {
                            const float sw_cur = swp[-ky * sws];
                            pair_eval<NG, MODE, COUNT>(cen[ky][0], cur, sw_cur, acc[ky][0]);       // dx = j
                            pair_eval<NG, MODE, COUNT>(cen[ky][1], cur, sw_prev[ky], acc[ky][1]);  // dx = j - 1
                            sw_prev[ky] = sw_cur;
                        }
                        rp += SMC_LINE_BYTES;
                        cur = lds_rec(rp);  // record j + 2 (one past the end stays inside the slot)
#pragma unroll
                        for (int ky = 0; ky < PY; ky++) {
                            const float sw_cur = swp[-ky * sws + 1];
                            pair_eval<NG, MODE, COUNT>(cen[ky][0], nxt, sw_cur, acc[ky][0]);       // dx = j + 1
                            pair_eval<NG, MODE, COUNT>(cen[ky][1], nxt, sw_prev[ky], acc[ky][1]);  // dx = j
                            sw_prev[ky] = sw_cur;
                        }
                        swp += 2;
                    }
                    if (j <= hi) {
#pragma unroll
                        for (int ky = 0; ky < PY; ky++) {
                            const float sw_cur = swp[-ky * sws];
                            pair_eval<NG, MODE, COUNT>(cen[ky][0], cur, sw_cur, acc[ky][0]);
                            pair_eval<NG, MODE, COUNT>(cen[ky][1], cur, sw_prev[ky], acc[ky][1]);
                        }
                    }
                    (void)d1;
                }
            }
            __syncwarp();
            // the last warp to leave the slot refills it at once: nobody ever blocks on a free slot, and a slow warp
            // does not delay the loads of the others
            if (lane == 0 && smem_inc_acq_rel(&left[s]) == kWarps - 1) {
                left[s] = 0;  // ordered before the next round of increments by the full-barrier completion
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(pos + g.depth);
            }
        }

        // write the tile (stat_denoiser.cu:341-344); centre fix-up: the reference gives the centre tap weight 1
        // unconditionally (:318-323), the loop above only if the centre passes its own test (it does unless NaN).
        const SmcPtrStepSz o = (p.denoise_film && tc.z == 0) ? p.film_filtered : p.out_ptrs[tc.z];
#pragma unroll
        for (int ky = 0; ky < PY; ky++) {
            const int y = tc.y0 + ky;
            if (y >= p.row_end) continue;
#pragma unroll
            for (int kx = 0; kx < 2; kx++) {
                const int x = xf + kx;
                if (x >= p.W) continue;
                Acc a = acc[ky][kx];
                const SmcRec rc = ldg_rec(img + (size_t)(y + r) * row_bytes, x + p.padX);
                if (!smc_member<3, NG, MODE>(cen[ky][kx], rc)) {
                    a.n01.x = __fadd_rn(a.n01.x, rc.c2.x);
                    a.n01.y = __fadd_rn(a.n01.y, rc.c2.y);
                    a.n2d.x = __fadd_rn(a.n2d.x, rc.c1.z);
                    a.n2d.y = __fadd_rn(a.n2d.y, 1.f);
                    a.cnt += 1;
                }
                float *op = (float *)(o.data + (size_t)y * o.step) + x * 3;
                op[0] = __fdiv_rn(a.n01.x, a.n2d.y);
                op[1] = __fdiv_rn(a.n01.y, a.n2d.y);
                op[2] = __fdiv_rn(a.n2d.x, a.n2d.y);
                if (COUNT && p.accepted && p.accepted[tc.z].data)
                    ((int *)(p.accepted[tc.z].data + (size_t)y * p.accepted[tc.z].step))[x] = a.cnt;
            }
        }
    }
    if (p.trace) {
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[4 * blockIdx.x + 1] = t;
            p.trace[4 * blockIdx.x + 3] = tl;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-warp streaming variant.  Same per-pair arithmetic and the same 2 x PY pixels per thread, but every WARP is an
// independent worker: it owns a 64 x PY tile, a private two-slot ring of record rows (64 + 2r records wide, filled by
// its own 1-D bulk copies) and its own tile sequence drawn from the global counter.  Nothing couples the warps of a
// CTA after the start-up barrier: no shared ring (in the CTA-wide kernel a warp that runs ahead blocks on the slot
// the slowest warp still reads: 10 % of all warp samples sat in that wait, ncu r1b), no per-row shared atomics, and
// the per-row bookkeeping is a few integer instructions (slot = row parity; no divisions).  One CTA per SM holds as
// many warps as registers (168 per thread -> 12) and shared memory allow and shares one copy of the spatial table.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWTileW = 64;
constexpr int kWMaxThreads = 384;

struct WarpGeom {
    int tiles_x, tiles_y, total_tiles;
    int nr;            // record rows per tile = 2r + PY - 1
    int slot_bytes;    // bytes per ring slot (multiple of 128)
    int seg_max_rec;   // records per full segment (even)
    int sw_rows;
    int nwarps;        // warps per CTA
    const int2 *rowrange;
};

struct WTile {
    int z, x0, y0;
    const unsigned char *src0;  // first record row of the tile's segment
    uint32_t bytes;             // bytes per row segment
};

__device__ __forceinline__ WTile wtile_make(int t, const SmcFilterParams &p, const WarpGeom &g, int PY) {
    WTile w;
    const int per_img = g.tiles_x * g.tiles_y;
    w.z = t / per_img;
    const int rem = t - w.z * per_img;
    const int ty = rem / g.tiles_x;
    w.x0 = (rem - ty * g.tiles_x) * kWTileW;
    w.y0 = p.row_begin + ty * PY;
    const int seg_start = seg_start_of(p, w.x0);
    const int nrec = min(g.seg_max_rec, p.rec_pitch - seg_start);  // even
    w.src0 = p.rec + (size_t)w.z * p.rec_image_stride + (size_t)w.y0 * smc_rec_row_bytes(p.rec_pitch) +
             smc_rec_offset(seg_start);
    w.bytes = (uint32_t)(nrec / 2) * SMC_LINE_BYTES;
    return w;
}

// One record row (already in the warp's ring slot) against the warp's 2 x PY centre pixels per lane.
template <int C, int NG, int PY, int MODE, bool COUNT, bool FILM>
__device__ __forceinline__ void warp_row(const SmcFilterParams &p, const SmcCentre<C, NG> (&cen)[PY][2], Acc (&acc)[PY][2],
                                         const int2 *rowrange, const float *sw, const unsigned char *slot, int i, int base_idx) {
    const int r = p.radius;
    {
                // table row of centre row ky: dy = i - r - ky  ->  row index i - ky + margin_y
                int lo = 1 << 20, hi = -(1 << 20);
#pragma unroll
                for (int ky = 0; ky < PY; ky++) {
                    const int2 rr = rowrange[i - ky + p.sw_margin_y];
                    lo = min(lo, rr.x);
                    hi = max(hi, rr.y);
                }
                if (lo <= hi) {
                    const float *swp = sw + (i + p.sw_margin_y) * p.sw_stride + (r + p.sw_margin_x) + lo;
                    const int sws = p.sw_stride;
                    float sw_prev[PY];
#pragma unroll
                    for (int ky = 0; ky < PY; ky++) sw_prev[ky] = swp[-ky * sws - 1];
                    const int first = base_idx + lo;
                    const unsigned char *rp = slot + smc_rec_offset(first);
                    const int d0 = (first & 1) ? SMC_LINE_BYTES - SMC_REC_BYTES : SMC_REC_BYTES;
                    SmcRec cur = lds_rec(rp);
                    int j = lo;
                    for (; j + 1 <= hi; j += 2) {
                        const SmcRec nxt = lds_rec(rp + d0);
#pragma unroll
                        for (int ky = 0; ky < PY; ky++) {
                            const float sw_cur = swp[-ky * sws];
                            pair_eval_c<C, NG, MODE, COUNT, FILM>(cen[ky][0], cur, sw_cur, acc[ky][0]);       // dx = j
                            pair_eval_c<C, NG, MODE, COUNT, FILM>(cen[ky][1], cur, sw_prev[ky], acc[ky][1]);  // dx = j - 1
                            sw_prev[ky] = sw_cur;
                        }
                        rp += SMC_LINE_BYTES;
                        cur = lds_rec(rp);  // record j + 2 (one past the end stays inside the slot)
#pragma unroll
                        for (int ky = 0; ky < PY; ky++) {
                            const float sw_cur = swp[-ky * sws + 1];
                            pair_eval_c<C, NG, MODE, COUNT, FILM>(cen[ky][0], nxt, sw_cur, acc[ky][0]);       // dx = j + 1
                            pair_eval_c<C, NG, MODE, COUNT, FILM>(cen[ky][1], nxt, sw_prev[ky], acc[ky][1]);  // dx = j
                            sw_prev[ky] = sw_cur;
                        }
                        swp += 2;
                    }
                    if (j <= hi) {
#pragma unroll
                        for (int ky = 0; ky < PY; ky++) {
                            const float sw_cur = swp[-ky * sws];
                            pair_eval_c<C, NG, MODE, COUNT, FILM>(cen[ky][0], cur, sw_cur, acc[ky][0]);
                            pair_eval_c<C, NG, MODE, COUNT, FILM>(cen[ky][1], cur, sw_prev[ky], acc[ky][1]);
                        }
                    }
                }
            }
}

template <int C, int NG, int PY, int MODE, bool COUNT>
__global__ void __launch_bounds__(kWMaxThreads, 1) filter_warp_kernel(const SmcFilterParams p, const WarpGeom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [rings: nwarps x 2 x slot_bytes][sw table][rowrange][barriers: nwarps x 2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = smem + (size_t)warp * 2 * g.slot_bytes;
    float *sw = (float *)(smem + (size_t)g.nwarps * 2 * g.slot_bytes);
    int2 *rowrange = (int2 *)(sw + g.sw_rows * p.sw_stride);
    uint64_t *full = (uint64_t *)(rowrange + g.sw_rows) + 2 * warp;

    for (int i = threadIdx.x; i < g.sw_rows * p.sw_stride; i += blockDim.x) sw[i] = p.sw[i];
    for (int i = threadIdx.x; i < g.sw_rows; i += blockDim.x) rowrange[i] = g.rowrange[i];
    if (lane == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ int warps_done;
    if (threadIdx.x == 0) {
        warps_done = 0;
        smc_halo_wait(p.halo);  // multi-GPU: the neighbours' prepasses have stored this step's halo records into our array
    }
    __syncthreads();  // the only CTA-wide synchronisation

    [&]() {  // the warp's work; returns when the queue is empty
    const int r = p.radius;
    const size_t row_bytes = smc_rec_row_bytes(p.rec_pitch);
    const int total_warps = (int)gridDim.x * g.nwarps;
    // warps that run side by side start on neighbouring tiles: the rows they share come out of L2
    int t_cur = (int)blockIdx.x * g.nwarps + warp;
    if (t_cur >= g.total_tiles) return;
    WTile ti = wtile_make(t_cur, p, g, PY);
    const uint32_t full0 = smem_u32(&full[0]);
    const uint32_t ring0 = smem_u32(ring);
    auto issue = [&](const unsigned char *src, uint32_t bytes, int s) {  // lane 0 only
        const uint32_t bar = full0 + 8u * (uint32_t)s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ring0 + (uint32_t)s * (uint32_t)g.slot_bytes),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
    };
    if (lane == 0) {
        issue(ti.src0, ti.bytes, 0);
        issue(ti.src0 + row_bytes, ti.bytes, 1);  // nr >= 3
    }

    uint32_t pos = 0;  // rows consumed so far: slot = pos & 1, phase parity = (pos >> 1) & 1
    for (;;) {
        // the tile after this one: asked for now, first needed two rows before the end of the tile
        int nxt_raw = 0;
        if (lane == 0) nxt_raw = total_warps + atomicAdd(p.tile_counter, 1);  // ~1 us, once per ~350 us tile
        int t_next = g.total_tiles;
        WTile tn = ti;

        const unsigned char *img = p.rec + (size_t)ti.z * p.rec_image_stride;
        const int xf = ti.x0 + 2 * lane;  // first of this thread's two columns
        const int base_idx = xf + p.padX - seg_start_of(p, ti.x0);  // slot index of the record at dx = 0, column kx = 0

        const bool film_out = p.denoise_film && ti.z == 0;  // warp-uniform
        SmcCentre<C, NG> cen[PY][2];
        Acc acc[PY][2];
#pragma unroll
        for (int ky = 0; ky < PY; ky++)
#pragma unroll
            for (int kx = 0; kx < 2; kx++) {
                const int yc = min(ti.y0 + ky, p.H - 1), xc = min(xf + kx, p.W - 1);
                const SmcRec rc = ldg_rec(img + (size_t)(yc + r) * row_bytes, xc + p.padX);
                smc_make_centre<C, NG, MODE>(rc, cen[ky][kx]);
                acc[ky][kx].n01.x = acc[ky][kx].n01.y = acc[ky][kx].n2d.x = acc[ky][kx].n2d.y = acc[ky][kx].ns = 0.f;
                acc[ky][kx].cnt = 0;
            }

        for (int i = 0; i < g.nr; i++, pos++) {
            const int s = (int)(pos & 1u);
            {
                const uint32_t bar = full0 + 8u * (uint32_t)s, parity = (pos >> 1) & 1u;
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "WWAIT_LOOP:\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                    "@p bra WDONE;\n"
                    "bra WWAIT_LOOP;\n"
                    "WDONE:\n"
                    "}\n" ::"r"(bar),
                    "r"(parity)
                    : "memory");
            }
            if (C == 1 && !film_out)
                warp_row<C, NG, PY, MODE, COUNT, false>(p, cen, acc, rowrange, sw, ring + (size_t)s * g.slot_bytes, i, base_idx);
            else
                warp_row<C, NG, PY, MODE, COUNT, true>(p, cen, acc, rowrange, sw, ring + (size_t)s * g.slot_bytes, i, base_idx);
            __syncwarp();  // every lane has read the slot
            if (i == g.nr - 2) {  // the next tile's first row goes into this slot
                t_next = __shfl_sync(0xffffffffu, nxt_raw, 0);
                if (t_next < g.total_tiles) tn = wtile_make(t_next, p, g, PY);
            }
            if (lane == 0) {
                const int ii = i + 2;
                if (ii < g.nr) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(ti.src0 + (size_t)ii * row_bytes, ti.bytes, s);
                } else if (t_next < g.total_tiles) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(tn.src0 + (size_t)(ii - g.nr) * row_bytes, tn.bytes, s);
                }
            }
        }

        // write the tile (stat_denoiser.cu:341-344 RGB, :263-273 scalar); centre fix-up as in the CTA-wide kernel
        const SmcPtrStepSz o = (C == 3 && film_out) ? p.film_filtered : p.out_ptrs[ti.z];
#pragma unroll
        for (int ky = 0; ky < PY; ky++) {
            const int y = ti.y0 + ky;
            if (y >= p.row_end) continue;
#pragma unroll
            for (int kx = 0; kx < 2; kx++) {
                const int x = xf + kx;
                if (x >= p.W) continue;
                Acc a = acc[ky][kx];
                const SmcRec rc = ldg_rec(img + (size_t)(y + r) * row_bytes, x + p.padX);
                if (!smc_member<C, NG, MODE>(cen[ky][kx], rc)) {
                    a.n01.x = __fadd_rn(a.n01.x, rc.c2.x);
                    a.n01.y = __fadd_rn(a.n01.y, rc.c2.y);
                    a.n2d.x = __fadd_rn(a.n2d.x, rc.c1.z);
                    a.ns = __fadd_rn(a.ns, rc.c1.x);
                    a.n2d.y = __fadd_rn(a.n2d.y, 1.f);
                    a.cnt += 1;
                }
                if (C == 3) {
                    float *op = (float *)(o.data + (size_t)y * o.step) + x * 3;
                    op[0] = __fdiv_rn(a.n01.x, a.n2d.y);
                    op[1] = __fdiv_rn(a.n01.y, a.n2d.y);
                    op[2] = __fdiv_rn(a.n2d.x, a.n2d.y);
                } else {
                    ((float *)(o.data + (size_t)y * o.step))[x] = __fdiv_rn(a.ns, a.n2d.y);
                    if (film_out) {
                        float *op = (float *)(p.film_filtered.data + (size_t)y * p.film_filtered.step) + x * 3;
                        op[0] = __fdiv_rn(a.n01.x, a.n2d.y);
                        op[1] = __fdiv_rn(a.n01.y, a.n2d.y);
                        op[2] = __fdiv_rn(a.n2d.x, a.n2d.y);
                    }
                }
                if (COUNT && p.accepted && p.accepted[ti.z].data)
                    ((int *)(p.accepted[ti.z].data + (size_t)y * p.accepted[ti.z].step))[x] = a.cnt;
            }
        }
        if (t_next >= g.total_tiles) break;
        t_cur = t_next;
        ti = tn;
    }
    }();
    // multi-GPU: when the last warp of the last CTA has read its last record, the neighbours may overwrite our halo rows
    __syncwarp();
    if (lane == 0 && atomicAdd(&warps_done, 1) == g.nwarps - 1) smc_halo_signal_last(p.halo, (int)gridDim.x);
}

template <typename K>
int launch_wk(smc_context *ctx, K k, const SmcFilterParams &p, const WarpGeom &g, size_t smem) {
    SMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<long long>((g.total_tiles + g.nwarps - 1) / g.nwarps, (long long)ctx->sm_count);
    SMC_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int), ctx->stream));
    k<<<grid, g.nwarps * 32, smem, ctx->stream>>>(p, g);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}

template <int NG, int MODE>
int launch_w(smc_context *ctx, const SmcFilterParams &p, const WarpGeom &g, size_t smem) {
    if (p.C == 1) {
        if (p.accepted != nullptr) return launch_wk(ctx, filter_warp_kernel<1, NG, 2, MODE, true>, p, g, smem);
        return launch_wk(ctx, filter_warp_kernel<1, NG, 2, MODE, false>, p, g, smem);
    }
    if (p.accepted != nullptr) return launch_wk(ctx, filter_warp_kernel<3, NG, 2, MODE, true>, p, g, smem);
    return launch_wk(ctx, filter_warp_kernel<3, NG, 2, MODE, false>, p, g, smem);
}

template <typename K>
int launch_k(smc_context *ctx, K k, const SmcFilterParams &p, const StreamGeom &g, size_t smem, int *query_per_sm) {
    SMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, smem));
    if (query_per_sm) {
        *query_per_sm = per_sm;
        return SMC_OK;
    }
    if (per_sm < 1) SMC_FAIL(SMC_ERR_CUDA, "streaming filter does not fit on an SM (smem %zu)", smem);
    // persistent grid: one wave of resident CTAs, each walking tiles b, b + grid, ... (neighbouring tiles run
    // concurrently, so the rows they share are fetched from HBM once and served from L2)
    const int grid = (int)std::min<long long>(g.total_tiles, (long long)per_sm * ctx->sm_count);
    SMC_CUDA(cudaMemsetAsync(p.tile_counter, 0, sizeof(int), ctx->stream));
    k<<<grid, kThreads, smem, ctx->stream>>>(p, g);
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}

template <int NG, int PY, int MODE>
int launch(smc_context *ctx, const SmcFilterParams &p, const StreamGeom &g, size_t smem, int *q) {
    if (p.accepted != nullptr) return launch_k(ctx, filter_stream_kernel<NG, PY, MODE, true>, p, g, smem, q);
    return launch_k(ctx, filter_stream_kernel<NG, PY, MODE, false>, p, g, smem, q);
}

}  // namespace

// geometry shared by the support check and the launcher
static bool stream_geometry(const SmcFilterParams &p, int PY, StreamGeom &g, size_t &smem) {
    const int rows = p.row_end - p.row_begin;
    g.tiles_x = (p.W + kTileW - 1) / kTileW;
    g.tiles_y = (rows + PY - 1) / PY;
    const long long total = (long long)g.tiles_x * g.tiles_y * p.ptr_count;
    if (total <= 0 || total > 0x7fffffffLL) return false;
    g.total_tiles = (int)total;
    g.nr = 2 * p.radius + PY - 1;
    // widest reach: first record x0 - r (rounded down to even), last x0 + 255 + r, plus one prefetched past the end
    g.seg_max_rec = kTileW + 2 * p.radius + 4;
    g.slot_bytes = (((g.seg_max_rec / 2) * SMC_LINE_BYTES + 127) / 128) * 128;
    g.sw_rows = 2 * p.radius + 2 * p.sw_margin_y;
    const size_t fixed = (size_t)g.sw_rows * p.sw_stride * 4 + (size_t)g.sw_rows * 8 + 8 * 8 + 8 * 4 + 4 * 4 + 64;
    // Ring depth / residency (measured on B200, 4K r=20, profiles/r1_variants.md): 3 CTAs per SM with a 3-deep ring
    // beat 2 CTAs with 4 slots when PY = 2 (166 registers per thread allow 3 CTAs); PY = 4 (250 registers) is limited to
    // 2 CTAs by the register file, where 4 slots fit.  Each CTA also reserves 1 KB of shared memory.
    smem = 0;
    int want = 0;
    if (const char *e = getenv("SMC_STREAM_DEPTH")) want = atoi(e);  // tuning knob
    const size_t s3 = 3 * (size_t)g.slot_bytes + fixed, s4 = 4 * (size_t)g.slot_bytes + fixed;
    if (PY == 2 && 3 * (s3 + 1024) <= 227 * 1024) {
        g.depth = 3;
        smem = s3;
    } else if (2 * (s4 + 1024) <= 227 * 1024) {
        g.depth = 4;
        smem = s4;
    } else {
        g.depth = 3;
        smem = s3;
    }
    if (want >= 2 && want <= 8) {
        g.depth = want;
        smem = (size_t)want * g.slot_bytes + fixed;
    }
    if (g.depth > g.nr) {  // warps drift by fewer than `depth` rows; the tile-id ring relies on that being at most one tile
        g.depth = g.nr;
        smem = (size_t)g.depth * g.slot_bytes + fixed;
    }
    return smem <= 220 * 1024;
}

// geometry of the per-warp variant: one CTA per SM with as many warps as shared memory allows (at most 12: registers)
static bool warp_geometry(const SmcFilterParams &p, int PY, WarpGeom &g, size_t &smem) {
    const int rows = p.row_end - p.row_begin;
    g.tiles_x = (p.W + kWTileW - 1) / kWTileW;
    g.tiles_y = (rows + PY - 1) / PY;
    const long long total = (long long)g.tiles_x * g.tiles_y * p.ptr_count;
    if (total <= 0 || total > 0x3fffffffLL) return false;
    g.total_tiles = (int)total;
    g.nr = 2 * p.radius + PY - 1;
    if (g.nr < 3) return false;
    g.seg_max_rec = kWTileW + 2 * p.radius + 4;
    g.slot_bytes = (((g.seg_max_rec / 2) * SMC_LINE_BYTES + 127) / 128) * 128;
    g.sw_rows = 2 * p.radius + 2 * p.sw_margin_y;
    const size_t tables = (size_t)g.sw_rows * p.sw_stride * 4 + (size_t)g.sw_rows * 8;
    const size_t avail = 227 * 1024 - 1024;
    if (tables + 64 >= avail) return false;
    int nw = (int)((avail - tables) / (2 * (size_t)g.slot_bytes + 16));
    if (const char *e = getenv("SMC_WARP_NWARPS")) nw = std::min(nw, atoi(e));  // tuning knob
    g.nwarps = std::min(nw, kWMaxThreads / 32);
    smem = (size_t)g.nwarps * 2 * g.slot_bytes + tables + (size_t)g.nwarps * 16;
    return g.nwarps >= 4;
}

// which streaming variant: per-warp rings (default when they fit) or the CTA-wide ring; SMC_STREAM_KERNEL=cta|warp
static bool use_warp_variant(const SmcFilterParams &p, int py) {
    if (py != 2 && p.C == 3) return false;  // scalar statistics always take the per-warp 2 x 2 kernel
    if (const char *e = getenv("SMC_STREAM_KERNEL")) {
        if (!strcmp(e, "cta")) return false;
    }
    WarpGeom g;
    size_t smem = 0;
    return warp_geometry(p, 2, g, smem);
}

bool smc_filter_stream_syncs_halo(const SmcFilterParams &p, int py) { return use_warp_variant(p, py); }

bool smc_filter_stream_supported(const SmcFilterParams &p, int sm_count, const char **name) {
    (void)sm_count;
    if (p.C != 3 && p.C != 1) return false;
    if (p.NGX > 0) return false;  // more G-buffer channels than the record holds: generic kernel
    if (p.radius < 1 || p.radius > 64) return false;
    if (p.sw_margin_y < 3 || p.sw_margin_x < 2) return false;
    if (!(p.NG == 0 || p.NG == 3 || p.NG == 6 || p.NG == 7)) return false;
    if (p.padX < p.radius || (p.padX & 1) || (p.rec_pitch & 1)) return false;
    StreamGeom g;
    size_t smem = 0;
    if (!stream_geometry(p, 2, g, smem) || !stream_geometry(p, 4, g, smem)) return false;
    // scalar statistics (multichannelstats = false, MIS win rates): per-warp kernel only
    if (p.C == 1 && !use_warp_variant(p, 2)) return false;
    if (name) *name = "stream";
    return true;
}

static int stream_dispatch(smc_context *ctx, const SmcFilterParams &p, const int2 *d_rowrange, int py,
                           const char **name, int *query_per_sm);

int smc_launch_filter_stream(smc_context *ctx, const SmcFilterParams &p, const int2 *d_rowrange, int py,
                             const char **name) {
    return stream_dispatch(ctx, p, d_rowrange, py, name, nullptr);
}

int smc_filter_stream_resident_ctas(const SmcFilterParams &p, int py, int sm_count, int *tile_w) {
    if (use_warp_variant(p, py)) {
        WarpGeom g;
        size_t smem = 0;
        warp_geometry(p, 2, g, smem);
        if (tile_w) *tile_w = kWTileW;
        return g.nwarps * sm_count;
    }
    if (tile_w) *tile_w = kTileW;
    int per_sm = 0;
    if (stream_dispatch(nullptr, p, nullptr, py, nullptr, &per_sm) != SMC_OK || per_sm < 1) per_sm = 1;
    return per_sm * sm_count;
}

static int stream_dispatch(smc_context *ctx, const SmcFilterParams &p, const int2 *d_rowrange, int py,
                           const char **name, int *query_per_sm) {
    int *q = query_per_sm;
    if (!q && use_warp_variant(p, py)) {
        WarpGeom wg;
        size_t wsmem = 0;
        warp_geometry(p, 2, wg, wsmem);
        wg.rowrange = d_rowrange;
        static thread_local char wnm[64];
        snprintf(wnm, sizeof(wnm), "stream-warp<%sNG=%d,PY=2,%s,W=%d>", p.C == 1 ? "C=1," : "", p.NG, p.mode ? "moon" : "welch", wg.nwarps);
        if (name) *name = wnm;
        switch (p.NG) {
            case 0: return p.mode == 0 ? launch_w<0, 0>(ctx, p, wg, wsmem) : launch_w<0, 1>(ctx, p, wg, wsmem);
            case 3: return p.mode == 0 ? launch_w<3, 0>(ctx, p, wg, wsmem) : launch_w<3, 1>(ctx, p, wg, wsmem);
            case 6: return p.mode == 0 ? launch_w<6, 0>(ctx, p, wg, wsmem) : launch_w<6, 1>(ctx, p, wg, wsmem);
            case 7: return p.mode == 0 ? launch_w<7, 0>(ctx, p, wg, wsmem) : launch_w<7, 1>(ctx, p, wg, wsmem);
            default: break;
        }
        SMC_FAIL(SMC_ERR_UNSUPPORTED, "streaming filter: NG=%d not instantiated", p.NG);
    }
    if (p.C != 3) SMC_FAIL(SMC_ERR_UNSUPPORTED, "streaming filter: scalar statistics need the per-warp variant");
    StreamGeom g;
    size_t smem = 0;
    if (!stream_geometry(p, py, g, smem)) SMC_FAIL(SMC_ERR_UNSUPPORTED, "streaming filter: geometry not supported");
    g.rowrange = d_rowrange;
    static thread_local char nm[64];
    snprintf(nm, sizeof(nm), "stream<NG=%d,PY=%d,%s,D=%d>", p.NG, py, p.mode ? "moon" : "welch", g.depth);
    if (name) *name = nm;
#define SMC_STREAM_CASE(NGv)                                                                                          \
    case NGv:                                                                                                         \
        if (py == 4) {                                                                                                \
            return p.mode == 0 ? launch<NGv, 4, 0>(ctx, p, g, smem, q) : launch<NGv, 4, 1>(ctx, p, g, smem, q); \
        } else {                                                                                                      \
            return p.mode == 0 ? launch<NGv, 2, 0>(ctx, p, g, smem, q) : launch<NGv, 2, 1>(ctx, p, g, smem, q); \
        }
    switch (p.NG) {
        SMC_STREAM_CASE(0)
        SMC_STREAM_CASE(3)
        SMC_STREAM_CASE(6)
        SMC_STREAM_CASE(7)
        default: break;
    }
#undef SMC_STREAM_CASE
    SMC_FAIL(SMC_ERR_UNSUPPORTED, "streaming filter: NG=%d not instantiated", p.NG);
}
