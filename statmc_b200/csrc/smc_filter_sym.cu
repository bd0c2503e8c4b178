// smc_filter_sym.cu -- the B200 filter kernel, symmetric form: every unordered pair of positions is evaluated ONCE.
//
// The reference (filter_kernel<float3>, stat_denoiser.cu:276-345) visits, for every pixel C, every tap I of the window and
// evaluates the membership test (:81-88) and the cross-bilateral weight (:90-112, :261).  Both are symmetric in (C, I):
//   disc_C + disc_I <= 2 mc_C mc_I        and        exp(dS2 * dSFactor + sum_g drFactor_g |g_C - g_I|^2),
// so the pair (P, Q = P + (dy, dx)) with a FORWARD offset -- dy == 0 and dx in [1, r], or dy in [1, r] -- is evaluated once and
// booked twice: w * V(Q) to P's sums and w * V(P) to Q's.  The window is half-open ([-r, r) in both axes, :247-248) and cut to
// the disc dS2 <= r^2 (:36); of all forward offsets in the disc only (0, r) and (r, 0) are not taps of P itself (their mirror
// images (0, -r), (-r, 0) are taps of Q), so those two are booked to Q only.  Replicated borders (BrdReplicate, :30-37) become
// VIRTUAL centres: P ranges over the image extended by r on the left, right and top; a virtual position carries the clamped
// pixel's record (the prepass has materialised them in the padded record array) and receives nothing.
//
// Work decomposition (one CTA per SM, every WARP an independent worker):
//   tile  = 64 x 2 centre positions: two columns x two rows per lane, centre statistics, values and forward sums in registers;
//           the tile streams the r + 2 record rows y0 .. y0 + r + 1 through a two-slot shared-memory ring filled by 1-D TMA bulk
//           copies (as smc_filter_stream.cu), one row ahead;
//   mirror sums of a streamed row live in a shared-memory row buffer (x, y, z, den per record), read-modify-written by the
//           lane that evaluates the pair (lanes of one instruction touch distinct records; the even / odd records of a row sit
//           in separate arrays so that the LDS.128 / STS.128 of a warp are conflict-free).  When the row is done its buffer is
//           flushed -- added to the partial sums the same unit left for that row one tile earlier, or stored on first touch --
//           into a scratch array private to the work unit, by the lanes themselves (each entry always by the same lane): no
//           atomics, a fixed order of summation;
//   unit  = a run of vertically adjacent tiles of one 64-column strip, processed top to bottom by one warp (units come off a
//           global atomic counter).  Only rows at unit seams and strip seams end up with more than one partial sum;
//   gather kernel: out(y, x) = (forward sums + the <= ~6 partial mirror sums that cover the pixel) / den.
// Per unordered pair: membership 7 + weight 9 + four packed FFMA2 (two sums each way) instructions, against 2 x (7 + 9 + 4) for
// the one-sided kernels.  Results differ from the one-sided kernels only in the ORDER of summation (accept / reject decisions
// are bit-identical: same operands, same roundings).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "smc_filter_math.cuh"
#include "smc_internal.h"

namespace {

constexpr int kTW = 64;        // centre columns per tile
#ifndef SMC_SYM_WARPS
#define SMC_SYM_WARPS 12       // warps per CTA the register allocation is bounded for (A/B knob)
#endif
constexpr int kSymMaxWarps = SMC_SYM_WARPS;
constexpr int kSymThreads = kSymMaxWarps * 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// What a lane keeps about one centre position: the test / weight operands (SmcCentre), its value for the mirror
// bookings, and its own (forward) sums.
template <int C, int NG>
struct SymCentre {
    SmcCentre<C, NG> c;
    float2 v01, v2o;   // RGB statistics: (V.x, V.y), (V.z, 1);  scalar statistics: v01.x = the scalar value
    float2 n01, n2d;   // forward sums: (num.x, num.y), (num.z, den);  scalar statistics: n01.x = num, n2d.y = den
    float2 d01;        // three scalar images per record (TRI): the denominators of images 0 and 1 (n2d = (num, den) of image 2)
    int cnt;
};

// mirror sums of one record, as they sit in the row buffer
struct Mir {
    float2 m01, m2d;
    float2 e;  // TRI: (num, den) of image 2; m01 = nums and m2d = dens of images 0 and 1
    int cnt;
};

// Orders the warp's shared-memory accesses for the compiler only.  The lanes of a warp run the row loop converged (its trip
// counts and branches are warp-uniform; the warp is re-converged with __syncwarp() after the barrier wait), and a warp's
// LDS / STS execute in program order, so lane l's read of the sums lane l + 1 wrote two steps earlier needs no instruction --
// only that the compiler does not hoist the read above the write.  -DSMC_SYM_SYNCWARP=1 uses __syncwarp() instead.
#ifndef SMC_SYM_SYNCWARP
#define SMC_SYM_SYNCWARP 0
#endif
__device__ __forceinline__ void sym_order() {
#if SMC_SYM_SYNCWARP
    __syncwarp();
#else
    asm volatile("" ::: "memory");
#endif
}

// 2^(sw - a), a = sum_k (g'_I,k - g'_C,k)^2, with the table holding (-sw, 0) so that the first squared difference is an FMA
// onto it (one instruction fewer than smc_weight(); -sw = +inf outside the disc -> weight 0)
template <int C, int NG>
__device__ __forceinline__ float sym_weight(const SmcCentre<C, NG> &c, const SmcRec &r, float2 nsw) {
    float a;
    if (NG >= 2) {
        float2 e = smc_add2(make_float2(r.c2.z, r.c2.w), c.g[0]);
        float2 acc = smc_fma2(e, e, nsw);
        if (NG >= 4) {
            e = smc_add2(make_float2(r.c3.x, r.c3.y), c.g[1]);
            acc = smc_fma2(e, e, acc);
        }
        if (NG >= 6) {
            e = smc_add2(make_float2(r.c3.z, r.c3.w), c.g[2]);
            acc = smc_fma2(e, e, acc);
        }
        a = __fadd_rn(acc.x, acc.y);
    } else {
        a = nsw.x;
    }
    if (NG & 1) {
        const float e = __fadd_rn(smc_rec_odd_g<NG>(r), c.go);
        a = __fmaf_rn(e, e, a);
    }
    return smc_ex2(-a);
}

// The membership test (is_not_discriminated, stat_denoiser.cu:81-88: discC + discI <= 2.f * meanC * meanI for every
// channel; same operands and roundings as smc_member<3, NG, 0>) applied to a weight: returns w if the pair is accepted, else 0.
// Written as one PTX block -- three chained compares and one select -- so that the weight is computed NEXT TO the test and the
// eight pair evaluations of a loop iteration stay one straight-line block for the scheduler.  (As `ok ? weight(...) : 0.f` the
// weight chain is predicated on the test and the pairs serialise two by two; as `ok ? w : 0.f` on a bool the select is
// distributed over the three compares: three selects.  Measured at 4K with 12 warps: 7.19 ms for this form, 7.36 ms with the
// weight predicated on the test.)
template <int C, int NG>
__device__ __forceinline__ float sym_gate(const SmcCentre<C, NG> &c, const SmcRec &r, float w, int *ok, float sz, float pz) {
    float g;
    if (C == 1) {  // scalar statistics: one channel
        const float s0 = __fadd_rn(c.d01.x, r.c0.z), p0 = __fmul_rn(c.t01.x, r.c0.x);
        int o;
        asm("{\n"
            ".reg .pred p;\n"
            "setp.le.f32 p, %2, %3;\n"
            "selp.f32 %0, %4, 0f00000000, p;\n"
            "selp.s32 %1, 1, 0, p;\n"
            "}\n"
            : "=f"(g), "=r"(o)
            : "f"(s0), "f"(p0), "f"(w));
        if (ok) *ok = o;
        return g;
    }
    const float2 sd = smc_add2(c.d01, make_float2(r.c0.z, r.c0.w));
    const float2 pm = smc_mul2(c.t01, make_float2(r.c0.x, r.c0.y));
    // (sz, pz) = (c.dz + r.d.z, c.tz * r.m.z): formed by the caller for the lane's two columns at once (ZPair)
    if (ok) {
        asm("{\n"
            ".reg .pred p;\n"
            "setp.le.f32 p, %2, %3;\n"
            "setp.le.and.f32 p, %4, %5, p;\n"
            "setp.le.and.f32 p, %6, %7, p;\n"
            "selp.f32 %0, %8, 0f00000000, p;\n"
            "selp.s32 %1, 1, 0, p;\n"
            "}\n"
            : "=f"(g), "=r"(*ok)
            : "f"(sd.x), "f"(pm.x), "f"(sd.y), "f"(pm.y), "f"(sz), "f"(pz), "f"(w));
    } else {
        asm("{\n"
            ".reg .pred p;\n"
            "setp.le.f32 p, %1, %2;\n"
            "setp.le.and.f32 p, %3, %4, p;\n"
            "setp.le.and.f32 p, %5, %6, p;\n"
            "selp.f32 %0, %7, 0f00000000, p;\n"
            "}\n"
            : "=f"(g)
            : "f"(sd.x), "f"(pm.x), "f"(sd.y), "f"(pm.y), "f"(sz), "f"(pz), "f"(w));
    }
    return g;
}

// One pair evaluation, booked both ways (forward to the centre, mirror to the record).  A rejected pair takes part with
// weight 0 (one select) instead of predicating the four accumulations.
#ifndef SMC_SYM_GATE
#define SMC_SYM_GATE 1  // 1: sym_gate() (weight next to the test); 0: `ok ? weight(...) : 0` (weight predicated on the test)
#endif
template <int C, int NG, bool COUNT>
__device__ __forceinline__ void pair_sym(SymCentre<C, NG> &s, const SmcRec &r, float2 nsw, Mir &m, float sz, float pz) {
#if SMC_SYM_GATE
    int ok = 0;
    const float w = sym_gate<C, NG>(s.c, r, sym_weight<C, NG>(s.c, r, nsw), COUNT ? &ok : nullptr, sz, pz);
#else
    const bool ok = smc_member<C, NG, 0>(s.c, r);
    const float w = ok ? sym_weight<C, NG>(s.c, r, nsw) : 0.f;
#endif
    if (C == 1) {  // scalar statistics: value in record slot 4; sums (num, -, -, den)
        s.n01.x = __fmaf_rn(w, r.c1.x, s.n01.x);
        s.n2d.y = __fadd_rn(s.n2d.y, w);
        m.m01.x = __fmaf_rn(w, s.v01.x, m.m01.x);
        m.m2d.y = __fadd_rn(m.m2d.y, w);
    } else {
        const float2 ww = make_float2(w, w);  // folded by ptxas into the scalar-broadcast operand form of FFMA2
        s.n01 = smc_fma2(ww, make_float2(r.c2.x, r.c2.y), s.n01);
        if (NG <= 6) {
            s.n2d = smc_fma2(ww, make_float2(r.c1.z, r.c1.w), s.n2d);  // record slot 7 == 1.0f: den += w * 1
            m.m2d = smc_fma2(ww, s.v2o, m.m2d);
        } else {
            s.n2d.x = __fmaf_rn(w, r.c1.z, s.n2d.x);
            s.n2d.y = __fadd_rn(s.n2d.y, w);
            m.m2d.x = __fmaf_rn(w, s.v2o.x, m.m2d.x);
            m.m2d.y = __fadd_rn(m.m2d.y, w);
        }
        m.m01 = smc_fma2(ww, s.v01, m.m01);
    }
    if (COUNT) {
        // a tap outside the disc has -sw = +inf -> w = 0: it adds nothing, but must not be counted
        const int one = (ok && nsw.x != INFINITY) ? 1 : 0;
        s.cnt += one;
        m.cnt += one;
    }
}

// The two forward offsets that are taps of the record's window only: booked to the record.
template <int C, int NG, bool COUNT>
__device__ __forceinline__ void pair_mirror_only(const SymCentre<C, NG> &s, const SmcRec &r, float2 nsw, Mir &m) {
    int ok = 0;
    const float w = sym_gate<C, NG>(s.c, r, sym_weight<C, NG>(s.c, r, nsw), COUNT ? &ok : nullptr,
                                    C == 3 ? __fadd_rn(s.c.dz, r.c1.y) : 0.f, C == 3 ? __fmul_rn(s.c.tz, r.c1.x) : 0.f);
    if (C == 1) {
        m.m01.x = __fmaf_rn(w, s.v01.x, m.m01.x);
        m.m2d.y = __fadd_rn(m.m2d.y, w);
    } else {
        const float2 ww = make_float2(w, w);
        m.m01 = smc_fma2(ww, s.v01, m.m01);
        if (NG <= 6) {
            m.m2d = smc_fma2(ww, s.v2o, m.m2d);
        } else {
            m.m2d.x = __fmaf_rn(w, s.v2o.x, m.m2d.x);
            m.m2d.y = __fadd_rn(m.m2d.y, w);
        }
    }
    if (COUNT) m.cnt += ok ? 1 : 0;
}

// ---- three scalar images per record (TRI) ----------------------------------------------------------------------------------
// The record has the RGB layout with image k in channel k's slots; the three membership tests gate three weights (instead of
// one test over three channels gating one weight) and every image keeps its own denominator.
template <int NG>
__device__ __forceinline__ void sym_gate_tri(const SmcCentre<3, NG> &c, const SmcRec &r, float w, float &w0, float &w1, float &w2) {
    const float2 sd = smc_add2(c.d01, make_float2(r.c0.z, r.c0.w));
    const float2 pm = smc_mul2(c.t01, make_float2(r.c0.x, r.c0.y));
    const float sz = __fadd_rn(c.dz, r.c1.y);
    const float pz = __fmul_rn(c.tz, r.c1.x);
    asm("{\n"
        ".reg .pred p;\n"
        "setp.le.f32 p, %3, %4;\n"
        "selp.f32 %0, %9, 0f00000000, p;\n"
        "setp.le.f32 p, %5, %6;\n"
        "selp.f32 %1, %9, 0f00000000, p;\n"
        "setp.le.f32 p, %7, %8;\n"
        "selp.f32 %2, %9, 0f00000000, p;\n"
        "}\n"
        : "=f"(w0), "=f"(w1), "=f"(w2)
        : "f"(sd.x), "f"(pm.x), "f"(sd.y), "f"(pm.y), "f"(sz), "f"(pz), "f"(w));
}

template <int NG>
__device__ __forceinline__ void pair_sym_tri(SymCentre<3, NG> &s, const SmcRec &r, float2 nsw, Mir &m) {
    float w0, w1, w2;
    sym_gate_tri<NG>(s.c, r, sym_weight<3, NG>(s.c, r, nsw), w0, w1, w2);
    const float2 w01 = make_float2(w0, w1), w22 = make_float2(w2, w2);
    s.n01 = smc_fma2(w01, make_float2(r.c2.x, r.c2.y), s.n01);
    s.d01 = smc_add2(s.d01, w01);
    m.m01 = smc_fma2(w01, s.v01, m.m01);
    m.m2d = smc_add2(m.m2d, w01);
    if (NG <= 6) {
        s.n2d = smc_fma2(w22, make_float2(r.c1.z, r.c1.w), s.n2d);  // record slot 7 == 1.0f
        m.e = smc_fma2(w22, s.v2o, m.e);
    } else {
        s.n2d.x = __fmaf_rn(w2, r.c1.z, s.n2d.x);
        s.n2d.y = __fadd_rn(s.n2d.y, w2);
        m.e.x = __fmaf_rn(w2, s.v2o.x, m.e.x);
        m.e.y = __fadd_rn(m.e.y, w2);
    }
}

template <int NG>
__device__ __forceinline__ void pair_mirror_only_tri(const SymCentre<3, NG> &s, const SmcRec &r, float2 nsw, Mir &m) {
    float w0, w1, w2;
    sym_gate_tri<NG>(s.c, r, sym_weight<3, NG>(s.c, r, nsw), w0, w1, w2);
    const float2 w01 = make_float2(w0, w1);
    m.m01 = smc_fma2(w01, s.v01, m.m01);
    m.m2d = smc_add2(m.m2d, w01);
    m.e.x = __fmaf_rn(w2, s.v2o.x, m.e.x);
    m.e.y = __fadd_rn(m.e.y, w2);
}

struct SymTile {
    int z, sx, uy, k, nt;       // image, strip, unit row, tile within the unit, tiles of the unit
    int x0, y0, i0, nrt;        // first centre column / row, first streamed row index, streamed rows
    const unsigned char *src0;  // record segment of streamed row i0
    uint32_t bytes;             // bytes per record segment
    size_t scr0;                // float4 index of the scratch row of the unit's first row (row y = yfirst)
    int yfirst;
};

__device__ __forceinline__ int sym_unit_t0(const SmcSymParams &g, int uy) {
    return uy < g.n_big ? uy * g.u_big : g.n_big * g.u_big + (uy - g.n_big) * g.u_small;
}
__device__ __forceinline__ int sym_unit_srow0(const SmcSymParams &g, int uy, int r) {
    return uy < g.n_big ? uy * (2 * g.u_big + r) : g.n_big * (2 * g.u_big + r) + (uy - g.n_big) * (2 * g.u_small + r);
}
__device__ __forceinline__ int sym_trow_unit(const SmcSymParams &g, int t) {
    const int tb = g.n_big * g.u_big;
    return t < tb ? t / g.u_big : g.n_big + (t - tb) / g.u_small;
}

__device__ __forceinline__ void sym_tile_place(SymTile &t, const SmcFilterParams &p, const SmcSymParams &g) {
    const int r = p.radius;
    const int t0 = sym_unit_t0(g, t.uy);
    t.x0 = g.xorg + t.sx * kTW;
    t.y0 = g.ystart + 2 * (t0 + t.k);
    t.i0 = max(0, p.row_begin - t.y0);
    t.nrt = r + 2 - t.i0;
    const int seg_start = (t.x0 + p.padX - r) & ~1;
    const int nrec = min(g.seg_rec, p.rec_pitch - seg_start);  // even
    t.src0 = p.rec + (size_t)t.z * p.rec_image_stride + (size_t)(t.y0 + t.i0 + r) * smc_rec_row_bytes(p.rec_pitch) +
             smc_rec_offset(seg_start);
    t.bytes = (uint32_t)(nrec / 2) * SMC_LINE_BYTES;
    t.yfirst = g.ystart + 2 * t0;
    t.scr0 = ((size_t)(t.z * g.n_strips + t.sx) * g.scratch_rows + sym_unit_srow0(g, t.uy, r)) * (size_t)g.seg_rec;
}

__device__ __forceinline__ void sym_unit_start(SymTile &t, int u, const SmcFilterParams &p, const SmcSymParams &g) {
    t.sx = u % g.n_strips;
    const int q = u / g.n_strips;
    t.uy = q % g.n_units_y;
    t.z = q / g.n_units_y;
    t.k = 0;
    const int t0 = sym_unit_t0(g, t.uy);
    t.nt = min(t.uy < g.n_big ? g.u_big : g.u_small, g.n_trows - t0);
    sym_tile_place(t, p, g);
}

// The z-channel operands of the membership test for the lane's two columns of one centre row, as register pairs: against one
// record both columns' tests need d.z + r.d.z and 2 m.z * r.m.z -- one packed add and one packed multiply for the two.
struct ZPair {
    float2 dz, tz;
};
// (Measured at 4K: the packed form costs more than it saves -- 7.39 against 7.19 ms: the extra register moves of the pairs run
// on the FMA pipe, which is the pipe the kernel is bound by -- so the default forms the operands with scalar instructions.)
#ifndef SMC_SYM_ZPACK
#define SMC_SYM_ZPACK 0
#endif
template <int C>
__device__ __forceinline__ void zpair_eval(const ZPair &z, const SmcRec &r, float2 &sz, float2 &pz) {
    if (C == 3 && SMC_SYM_ZPACK) {
        sz = smc_add2(z.dz, make_float2(r.c1.y, r.c1.y));
        pz = smc_mul2(z.tz, make_float2(r.c1.x, r.c1.x));
    } else if (C == 3) {
        sz = make_float2(__fadd_rn(z.dz.x, r.c1.y), __fadd_rn(z.dz.y, r.c1.y));
        pz = make_float2(__fmul_rn(z.tz.x, r.c1.x), __fmul_rn(z.tz.y, r.c1.x));
    } else {
        sz = pz = make_float2(0.f, 0.f);
    }
}

// One record row (already in the warp's ring slot; the mirror buffer holds zeros) against the warp's 2 x 2 centres per lane.
template <int C, int NG, bool COUNT, bool TRI>
__device__ __forceinline__ void sym_row(const SmcFilterParams &p, const SmcSymParams &g, SymCentre<C, NG> (&cen)[2][2],
                                        const ZPair (&zp)[2], const int2 *rowrange, const float2 *sw, const unsigned char *slot, float4 *macc,
                                        int *mcnt, int i, int base_idx) {
    const int r = p.radius;
    const int half = g.seg_rec >> 1;
    // table row of centre row ky: dy = i - ky  ->  row index i - ky + margin
    const int2 rr0 = rowrange[i + g.sw_my], rr1 = rowrange[i - 1 + g.sw_my];
    const int lo = min(rr0.x, rr1.x), hi = max(rr0.y, rr1.y);
    if (lo <= hi) {
        // `sw` holds the two table rows this streamed row needs: dy = i - 1 (centre row ky = 1), then dy = i (ky = 0)
        const float2 *swp1 = sw + (r + g.sw_mx) + lo;
        const float2 *swp0 = swp1 + g.sw_stride;
        float2 sw_prev0 = swp0[-1], sw_prev1 = swp1[-1];
        const int first = base_idx + lo;
        const unsigned char *rp = slot + smc_rec_offset(first);
        const int par = first & 1;
        const int d0 = par ? SMC_LINE_BYTES - SMC_REC_BYTES : SMC_REC_BYTES;
        float4 *mp0 = macc + par * half + (first >> 1);               // sums of record `first`, then first + 2, ...
        float4 *mp1 = macc + (par ^ 1) * half + ((first + 1) >> 1);   // sums of record first + 1, first + 3, ...
        int *cp0 = mcnt + par * half + (first >> 1), *cp1 = mcnt + (par ^ 1) * half + ((first + 1) >> 1);
        // TRI: the third image's (num, den) sit in a float2 array of the same indexing that takes the count array's place
        float2 *ep0 = (float2 *)mcnt + par * half + (first >> 1), *ep1 = (float2 *)mcnt + (par ^ 1) * half + ((first + 1) >> 1);
        SmcRec cur = lds_rec(rp);
        // (centre kx = 0 sees the record at dx = j, centre kx = 1 at dx = j - 1: the table value of the previous step)
        auto load_m = [&](const float4 *mp, const int *cp, const float2 *ep) {
            const float4 mv = *mp;
            Mir m;
            m.m01 = make_float2(mv.x, mv.y);
            m.m2d = make_float2(mv.z, mv.w);
            m.e = TRI ? *ep : make_float2(0.f, 0.f);
            m.cnt = COUNT ? *cp : 0;
            return m;
        };
        auto store_m = [&](float4 *mp, int *cp, float2 *ep, const Mir &m) {
            *mp = make_float4(m.m01.x, m.m01.y, m.m2d.x, m.m2d.y);
            if (TRI) *ep = m.e;
            if (COUNT) *cp = m.cnt;
        };
        auto pair = [&](SymCentre<C, NG> &sc, const SmcRec &rec, float2 nsw, Mir &m, float sz, float pz) {
            if constexpr (TRI) pair_sym_tri<NG>(sc, rec, nsw, m);
            else pair_sym<C, NG, COUNT>(sc, rec, nsw, m, sz, pz);
        };
        int j = lo;
        for (; j + 1 <= hi; j += 2) {
            // two records per iteration (an even and an odd one: their sums sit in different arrays), eight independent pair
            // evaluations for the scheduler to interleave
            const SmcRec nxt = lds_rec(rp + d0);
            const float2 a0 = swp0[0], a1 = swp1[0], b0 = swp0[1], b1 = swp1[1];
            sym_order();
            Mir ma = load_m(mp0, cp0, ep0), mb = load_m(mp1, cp1, ep1);
            float2 sa0, pa0, sa1, pa1, sb0, pb0, sb1, pb1;  // z-channel test operands: (column 0, column 1) per record and row
            zpair_eval<C>(zp[0], cur, sa0, pa0);
            zpair_eval<C>(zp[1], cur, sa1, pa1);
            zpair_eval<C>(zp[0], nxt, sb0, pb0);
            zpair_eval<C>(zp[1], nxt, sb1, pb1);
            pair(cen[0][0], cur, a0, ma, sa0.x, pa0.x);
            pair(cen[0][0], nxt, b0, mb, sb0.x, pb0.x);
            pair(cen[0][1], cur, sw_prev0, ma, sa0.y, pa0.y);
            pair(cen[0][1], nxt, a0, mb, sb0.y, pb0.y);
            pair(cen[1][0], cur, a1, ma, sa1.x, pa1.x);
            pair(cen[1][0], nxt, b1, mb, sb1.x, pb1.x);
            pair(cen[1][1], cur, sw_prev1, ma, sa1.y, pa1.y);
            pair(cen[1][1], nxt, a1, mb, sb1.y, pb1.y);
            store_m(mp0, cp0, ep0, ma);
            store_m(mp1, cp1, ep1, mb);
            sym_order();
            rp += SMC_LINE_BYTES;
            cur = lds_rec(rp);  // record j + 2 (one past the end stays inside the slot)
            sw_prev0 = b0;
            sw_prev1 = b1;
            swp0 += 2; swp1 += 2;
            mp0++; mp1++; cp0++; cp1++; ep0++; ep1++;
        }
        if (j <= hi) {
            sym_order();
            Mir ma = load_m(mp0, cp0, ep0);
            float2 sa0, pa0, sa1, pa1;
            zpair_eval<C>(zp[0], cur, sa0, pa0);
            zpair_eval<C>(zp[1], cur, sa1, pa1);
            pair(cen[0][0], cur, swp0[0], ma, sa0.x, pa0.x);
            pair(cen[0][1], cur, sw_prev0, ma, sa0.y, pa0.y);
            pair(cen[1][0], cur, swp1[0], ma, sa1.x, pa1.x);
            pair(cen[1][1], cur, sw_prev1, ma, sa1.y, pa1.y);
            store_m(mp0, cp0, ep0, ma);
            sym_order();
        }
    }
    // the two offsets booked to the record only: (0, r) in the centre's own row, (r, 0) r rows below
    auto special = [&](const SymCentre<C, NG> &s0, const SymCentre<C, NG> &s1, int dxs) {
        const float2 nsw = make_float2(-g.sw_special, 0.f);
#pragma unroll
        for (int kx = 0; kx < 2; kx++) {
            const int idx = base_idx + kx + dxs;
            const SmcRec rec = lds_rec(slot + smc_rec_offset(idx));
            float4 *mp = macc + (idx & 1) * half + (idx >> 1);
            int *cp = mcnt + (idx & 1) * half + (idx >> 1);
            float2 *ep = (float2 *)mcnt + (idx & 1) * half + (idx >> 1);
            __syncwarp();
            const float4 mv = *mp;
            Mir m;
            m.m01 = make_float2(mv.x, mv.y);
            m.m2d = make_float2(mv.z, mv.w);
            m.e = TRI ? *ep : make_float2(0.f, 0.f);
            m.cnt = COUNT ? *cp : 0;
            if constexpr (TRI) pair_mirror_only_tri<NG>(kx ? s1 : s0, rec, nsw, m);
            else pair_mirror_only<C, NG, COUNT>(kx ? s1 : s0, rec, nsw, m);
            *mp = make_float4(m.m01.x, m.m01.y, m.m2d.x, m.m2d.y);
            if (TRI) *ep = m.e;
            if (COUNT) *cp = m.cnt;
        }
    };
    if (i == 0) special(cen[0][0], cen[0][1], r);
    if (i == 1) special(cen[1][0], cen[1][1], r);
    if (i == r) special(cen[0][0], cen[0][1], 0);
    if (i == r + 1) special(cen[1][0], cen[1][1], 0);
}

template <int C, int NG, bool COUNT, bool TRI>
__global__ void __launch_bounds__(kSymThreads, 1) filter_sym_kernel(const SmcFilterParams p, const SmcSymParams g) {
    extern __shared__ __align__(128) unsigned char smem[];
    // layout: [per warp: 2 record slots | 2 x 2 spatial-table rows | mirror buffer (| count buffer)] ... [rowrange]
    //         [barriers: nwarps x 2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem + (size_t)warp * g.warp_bytes;
    unsigned char *ring = wbase;
    const uint32_t sw_bytes = 2u * (uint32_t)g.sw_stride * 8u;  // two table rows of (-sw, 0) pairs
    unsigned char *swb0 = wbase + 2 * (size_t)g.slot_bytes;
    float4 *macc = (float4 *)(swb0 + 2 * (size_t)sw_bytes);
    int *mcnt = (int *)((unsigned char *)macc + (size_t)g.macc_bytes);
    int2 *rowrange = (int2 *)(smem + (size_t)g.nwarps * g.warp_bytes);
    uint64_t *full = (uint64_t *)(rowrange + g.sw_rows) + 2 * warp;

    for (int i = threadIdx.x; i < g.sw_rows; i += blockDim.x) rowrange[i] = g.rowrange[i];
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[1])));
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
    const int rows_out = p.row_end - p.row_begin;
    // warps that run side by side start on neighbouring strips of the same unit row: the record rows they share come out of L2
    const int u_first = (int)blockIdx.x * g.nwarps + warp;
    if (u_first >= g.units_total) return;
    SymTile ti;
    sym_unit_start(ti, u_first, p, g);
    const uint32_t full0 = smem_u32(&full[0]);
    const uint32_t ring0 = smem_u32(ring), sws = smem_u32(swb0);

    // lane 0: queue the loads of stream position q = (tile t, streamed row i): the record segment and the two spatial-table
    // rows (dy = i - 1, i) into ring slot q & 1, completing on full[q & 1]
    auto issue = [&](const SymTile &t, int i, uint32_t q) {
        const uint32_t s = q & 1u;
        const uint32_t bar = full0 + 8u * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(t.bytes + sw_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ring0 + s * (uint32_t)g.slot_bytes),
                     "l"(t.src0 + (size_t)(i - t.i0) * row_bytes), "r"(t.bytes), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sws + s * sw_bytes),
                     "l"(g.sw + (size_t)(i - 1 + g.sw_my) * g.sw_stride), "r"(sw_bytes), "r"(bar)
                     : "memory");
    };
    // the tile after `t` in this warp's sequence (lane 0 knows the next unit): false when the queue is empty
    auto next_tile = [&](const SymTile &t, int u_next, SymTile &tn) {
        if (t.k + 1 < t.nt) {
            tn = t;
            tn.k = t.k + 1;
            sym_tile_place(tn, p, g);
            return true;
        }
        if (u_next >= g.units_total) return false;
        sym_unit_start(tn, u_next, p, g);
        return true;
    };
    if (lane == 0) {
        issue(ti, ti.i0, 0u);
        issue(ti, ti.i0 + 1, 1u);  // every tile streams at least two rows
    }
    // the mirror sums of the row being streamed: zero between rows
    for (int e = lane; e < g.seg_rec; e += 32) {
        macc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (COUNT) mcnt[e] = 0;
        if (TRI) ((float2 *)mcnt)[e] = make_float2(0.f, 0.f);
    }

    uint32_t pos = 0;  // rows consumed so far: ring slot = pos & 1, phase parity = (pos >> 1) & 1
    int nxt_raw = 0;   // lane 0: the unit after the current one
    for (;;) {
        // asked for when a unit starts (~1 us), first needed two rows before the unit's last tile ends
        if (ti.k == 0 && lane == 0) nxt_raw = total_warps + atomicAdd(g.unit_counter, 1);

        const unsigned char *img = p.rec + (size_t)ti.z * p.rec_image_stride;
        const int xf = ti.x0 + 2 * lane;  // first of this lane's two centre columns
        const int base_idx = xf + p.padX - ((ti.x0 + p.padX - r) & ~1);  // slot index of the record at dx = 0, column kx = 0

        SymCentre<C, NG> cen[2][2];
#pragma unroll
        for (int ky = 0; ky < 2; ky++)
#pragma unroll
            for (int kx = 0; kx < 2; kx++) {
                const int yc = ti.y0 + ky, xc = xf + kx;
                // virtual positions (replicated borders, rows of other bands) carry the padded array's record
                const int pcol = min(xc + p.padX, p.rec_pitch - 1);
                const SmcRec rc = ldg_rec(img + (size_t)(yc + r) * row_bytes, pcol);
                SymCentre<C, NG> &s = cen[ky][kx];
                smc_make_centre<C, NG, 0>(rc, s.c);
                // RGB: (V.x, V.y) and (V.z, 1) -- slot 7 of the record holds 1.0f when it is not the seventh G channel; taken
                // from the record so that the pair sits in adjacent registers as loaded (a literal 1.f would be rebuilt with
                // moves at every use).  Scalar statistics: the value is record slot 4.
                s.v01 = C == 3 ? make_float2(rc.c2.x, rc.c2.y) : make_float2(rc.c1.x, 0.f);
                s.v2o = C == 3 ? make_float2(rc.c1.z, NG <= 6 ? rc.c1.w : 1.f) : make_float2(0.f, 1.f);
                const bool real = yc >= p.row_begin && yc < p.row_end && xc >= 0 && xc < p.W;
                // the centre tap: weight 1 unconditionally (is_center, stat_denoiser.cu:78, :318-323)
                s.n01 = real ? s.v01 : make_float2(0.f, 0.f);
                s.n2d = real ? s.v2o : make_float2(0.f, 0.f);
                s.d01 = (TRI && real) ? make_float2(1.f, 1.f) : make_float2(0.f, 0.f);
                s.cnt = real ? 1 : 0;
            }

        ZPair zp[2];
#pragma unroll
        for (int ky = 0; ky < 2; ky++) {
            zp[ky].dz = make_float2(cen[ky][0].c.dz, cen[ky][1].c.dz);
            zp[ky].tz = make_float2(cen[ky][0].c.tz, cen[ky][1].c.tz);
        }

        for (int ii = 0; ii < ti.nrt; ii++, pos++) {
            const int i = ti.i0 + ii;
            const uint32_t s = pos & 1u;
            const bool first = ti.k == 0 || i >= r;  // rows r, r + 1 of a tile are new to the unit; everything is for its first tile
            float4 *sc = g.scratch + ti.scr0 + (size_t)(ti.y0 + i - ti.yfirst) * g.seg_rec;
            {
                const uint32_t bar = full0 + 8u * s, parity = (pos >> 1) & 1u;
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "SWAIT_LOOP:\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                    "@p bra SDONE;\n"
                    "bra SWAIT_LOOP;\n"
                    "SDONE:\n"
                    "}\n" ::"r"(bar),
                    "r"(parity)
                    : "memory");
            }
            __syncwarp();  // lanes leave the wait loop one by one: run the row converged (see sym_order())
            sym_row<C, NG, COUNT, TRI>(p, g, cen, zp, rowrange, (const float2 *)(swb0 + (size_t)s * sw_bytes), ring + (size_t)s * g.slot_bytes,
                                  macc, mcnt, i, base_idx);
            __syncwarp();  // every lane has read the slot and written its mirror sums
            // Flush the row's mirror sums into the unit's scratch: plain loads and stores by the lanes.  Entry e is always
            // handled by lane e % 32 (a strip's segments are aligned alike), so the partial sums a lane adds to are the ones it
            // stored itself one tile earlier: program order is all the ordering this needs.  The loads of up to four entries
            // per lane are issued together, and lane 0 queues the next TMA copies while they are in flight.
            int *scc = COUNT ? g.scratch_cnt + ti.scr0 + (size_t)(ti.y0 + i - ti.yfirst) * g.seg_rec : nullptr;
            float2 *sce = TRI ? g.scratch2 + ti.scr0 + (size_t)(ti.y0 + i - ti.yfirst) * g.seg_rec : nullptr;
            for (int e0 = 0; e0 < g.seg_rec; e0 += 128) {
                float4 o[4];
                float2 oe[4];
                int oc[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int e = e0 + 32 * k + lane;
                    o[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    oe[k] = make_float2(0.f, 0.f);
                    oc[k] = 0;
                    if (!first && e < g.seg_rec) {
                        o[k] = __ldcg(sc + e);
                        if (COUNT) oc[k] = __ldcg(scc + e);
                        if (TRI) oe[k] = __ldcg(sce + e);
                    }
                }
                if (e0 == 0 && lane == 0) {  // queue stream position pos + 2 into the slot every lane has just left
                    const int i2 = ii + 2;
                    if (i2 < ti.nrt) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(ti, ti.i0 + i2, pos + 2u);
                    } else {
                        SymTile tn;
                        if (next_tile(ti, nxt_raw, tn)) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            issue(tn, tn.i0 + (i2 - ti.nrt), pos + 2u);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int e = e0 + 32 * k + lane;
                    if (e < g.seg_rec) {
                        const float4 v = macc[e];
                        macc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                        // (first touch: o == 0, and 0 + v == v exactly, also for -0 sums, which cannot occur: weights are >= 0
                        // only in sign of the values; a -0 would at most become +0 in a sum nobody distinguishes)
                        __stcg(sc + e, first ? v : make_float4(__fadd_rn(o[k].x, v.x), __fadd_rn(o[k].y, v.y), __fadd_rn(o[k].z, v.z),
                                                               __fadd_rn(o[k].w, v.w)));
                        if (COUNT) {
                            const int c = mcnt[e];
                            mcnt[e] = 0;
                            __stcg(scc + e, c + oc[k]);
                        }
                        if (TRI) {
                            const float2 ve = ((float2 *)mcnt)[e];
                            ((float2 *)mcnt)[e] = make_float2(0.f, 0.f);
                            __stcg(sce + e, first ? ve : make_float2(__fadd_rn(oe[k].x, ve.x), __fadd_rn(oe[k].y, ve.y)));
                        }
                    }
                }
            }
            __syncwarp();  // the zeroed mirror buffer before the next row's sums
        }

        // forward sums of the tile's real centres (two adjacent pixels per lane and row: 32 contiguous bytes)
#pragma unroll
        for (int ky = 0; ky < 2; ky++)
#pragma unroll
            for (int kx = 0; kx < 2; kx++) {
                const int yc = ti.y0 + ky, xc = xf + kx;
                if (yc >= p.row_begin && yc < p.row_end && xc >= 0 && xc < p.W) {
                    const SymCentre<C, NG> &s = cen[ky][kx];
                    const size_t o = ((size_t)ti.z * rows_out + (yc - p.row_begin)) * p.W + xc;
                    if (TRI) {
                        g.fwd[o] = make_float4(s.n01.x, s.n01.y, s.d01.x, s.d01.y);
                        g.fwd2[o] = s.n2d;
                    } else {
                        g.fwd[o] = make_float4(s.n01.x, s.n01.y, s.n2d.x, s.n2d.y);
                    }
                    if (COUNT) g.fwd_cnt[o] = s.cnt;
                }
            }
        const int u_next = __shfl_sync(0xffffffffu, nxt_raw, 0);
        SymTile tn;
        if (!next_tile(ti, u_next, tn)) break;
        ti = tn;
    }
    }();
    // multi-GPU: when the last warp of the last CTA has read its last record, the neighbours may overwrite our halo rows
    __syncwarp();
    if (lane == 0 && atomicAdd(&warps_done, 1) == g.nwarps - 1) smc_halo_signal_last(p.halo, (int)gridDim.x);
}

// out(y, x) = (forward sums + every partial mirror sum that covers the pixel) / den   (stat_denoiser.cu:341-344)
template <int C, bool COUNT, bool TRI>
__global__ void __launch_bounds__(128) sym_gather_kernel(const SmcFilterParams p, const SmcSymParams g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = p.row_begin + blockIdx.y;
    const int z = blockIdx.z;
    if (x >= p.W) return;
    const int r = p.radius, rows_out = p.row_end - p.row_begin;
    const size_t o = ((size_t)z * rows_out + (y - p.row_begin)) * p.W + x;
    float4 s = g.fwd[o];
    float2 se = TRI ? g.fwd2[o] : make_float2(0.f, 0.f);
    int cnt = COUNT ? g.fwd_cnt[o] : 0;
    // tiles whose centre rows y - r .. y stream row y; strips whose centres reach column x
    const int tlo = (y - r - g.ystart) >> 1, thi = min(g.n_trows - 1, (y - g.ystart) >> 1);
    const int ulo = sym_trow_unit(g, tlo), uhi = sym_trow_unit(g, thi);
    const int a = x - r - (kTW - 1) - g.xorg;
    const int sxlo = a <= 0 ? 0 : (a + kTW - 1) / kTW, sxhi = min(g.n_strips - 1, (x + r - g.xorg) / kTW);
    for (int sx = sxlo; sx <= sxhi; sx++) {
        const int idx = x + p.padX - ((g.xorg + sx * kTW + p.padX - r) & ~1);
        const int e = (idx & 1) * (g.seg_rec >> 1) + (idx >> 1);
        for (int uy = ulo; uy <= uhi; uy++) {
            const int yrel = y - (g.ystart + 2 * sym_unit_t0(g, uy));
            const size_t q = (((size_t)(z * g.n_strips + sx) * g.scratch_rows + sym_unit_srow0(g, uy, r) + yrel) * g.seg_rec) + e;
            const float4 v = __ldg(g.scratch + q);
            s.x = __fadd_rn(s.x, v.x);
            s.y = __fadd_rn(s.y, v.y);
            s.z = __fadd_rn(s.z, v.z);
            s.w = __fadd_rn(s.w, v.w);
            if (COUNT) cnt += __ldg(g.scratch_cnt + q);
            if (TRI) {
                const float2 ve = __ldg(g.scratch2 + q);
                se.x = __fadd_rn(se.x, ve.x);
                se.y = __fadd_rn(se.y, ve.y);
            }
        }
    }
    if (TRI) {  // three scalar images: s = (num0, num1, den0, den1), se = (num2, den2)
        const float num[3] = {s.x, s.y, se.x}, den[3] = {s.z, s.w, se.y};
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (3 * z + k < g.images) {
                const SmcPtrStepSz ob = p.out_ptrs[3 * z + k];
                ((float *)(ob.data + (size_t)y * ob.step))[x] = __fdiv_rn(num[k], den[k]);
            }
        return;
    }
    if (C == 3) {
        const SmcPtrStepSz ob = (p.denoise_film && z == 0) ? p.film_filtered : p.out_ptrs[z];
        float *op = (float *)(ob.data + (size_t)y * ob.step) + x * 3;
        op[0] = __fdiv_rn(s.x, s.w);
        op[1] = __fdiv_rn(s.y, s.w);
        op[2] = __fdiv_rn(s.z, s.w);
    } else {  // scalar statistics (stat_denoiser.cu:263-273; the plan has no film to filter along: denoise_film == 0)
        const SmcPtrStepSz ob = p.out_ptrs[z];
        ((float *)(ob.data + (size_t)y * ob.step))[x] = __fdiv_rn(s.x, s.w);
    }
    if (COUNT && p.accepted && p.accepted[z].data) ((int *)(p.accepted[z].data + (size_t)y * p.accepted[z].step))[x] = cnt;
}

template <int C, int NG>
int launch_sym_ng(smc_context *ctx, const SmcFilterParams &p, const SmcSymParams &g, size_t smem, int grid) {
    if (C == 3 && g.tri) {
        auto k = filter_sym_kernel<3, NG, false, true>;
        SMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, g.nwarps * 32, smem, ctx->stream>>>(p, g);
    } else if (p.accepted != nullptr) {
        auto k = filter_sym_kernel<C, NG, true, false>;
        SMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, g.nwarps * 32, smem, ctx->stream>>>(p, g);
    } else {
        auto k = filter_sym_kernel<C, NG, false, false>;
        SMC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, g.nwarps * 32, smem, ctx->stream>>>(p, g);
    }
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}

template <int C>
int launch_sym_c(smc_context *ctx, const SmcFilterParams &p, const SmcSymParams &g, size_t smem, int grid) {
    switch (p.NG) {
        case 0: return launch_sym_ng<C, 0>(ctx, p, g, smem, grid);
        case 1: return launch_sym_ng<C, 1>(ctx, p, g, smem, grid);
        case 2: return launch_sym_ng<C, 2>(ctx, p, g, smem, grid);
        case 3: return launch_sym_ng<C, 3>(ctx, p, g, smem, grid);
        case 4: return launch_sym_ng<C, 4>(ctx, p, g, smem, grid);
        case 5: return launch_sym_ng<C, 5>(ctx, p, g, smem, grid);
        case 6: return launch_sym_ng<C, 6>(ctx, p, g, smem, grid);
        default: return launch_sym_ng<C, 7>(ctx, p, g, smem, grid);
    }
}

}  // namespace

// ---- host side: geometry -----------------------------------------------------------------------------------------------
bool smc_filter_sym_supported(const SmcFilterParams &p) {
    if (p.mode != SMC_MEMBER_WELCH) return false;  // the Moon test is not symmetric (stat_denoiser.cu:132-143)
    // scalar statistics: unless image 0 also filters the RGB film (five sums per tap: the one-sided per-warp kernel does that)
    if (!(p.C == 3 || (p.C == 1 && !p.denoise_film))) return false;
    if (p.tri && (p.C != 1 || p.accepted != nullptr)) return false;
    if (p.radius < 2 || p.radius > SMC_MAX_RADIUS) return false;
    if (p.NG < 0 || p.NG > 7 || p.NGX > 0) return false;
    if ((p.padX & 1) || (p.rec_pitch & 1)) return false;
    const int xorg = -(p.radius + (p.radius & 1));
    if (xorg + p.padX - p.radius < 0) return false;  // the first strip's record segment starts inside the padded array
    SmcSymParams g;
    size_t smem = 0;
    return smc_filter_sym_geometry(p, g, smem);
}

// fills everything of `g` that follows from the filter parameters (tables, scratch pointers and counters are the caller's)
bool smc_filter_sym_geometry(const SmcFilterParams &p, SmcSymParams &g, size_t &smem) {
    const int r = p.radius;
    g.tri = p.tri;
    g.images = p.images;
    g.xorg = -(r + (r & 1));
    g.n_strips = (p.W + r - g.xorg + kTW - 1) / kTW;
    g.ystart = p.row_begin - r;
    g.n_trows = (p.row_end - g.ystart + 1) / 2;
    g.seg_rec = ((kTW + 2 * r + 4 + 3) / 4) * 4;  // multiple of 4: the count rows (4 B per record) move as 16-byte bulk copies
    g.slot_bytes = (((g.seg_rec / 2) * SMC_LINE_BYTES + 127) / 128) * 128;
    g.macc_bytes = g.seg_rec * 16;
    const bool count = p.accepted != nullptr;
    g.sw_my = 2;
    g.sw_mx = 2;
    g.sw_rows = r + 1 + 2 * g.sw_my;
    g.sw_stride = 2 * r + 2 + 2 * g.sw_mx;  // even: two table rows of 8-byte entries move as one 16-byte-aligned bulk copy
    g.warp_bytes = 2 * g.slot_bytes + 2 * (2 * g.sw_stride * 8) + g.macc_bytes + (g.tri ? g.seg_rec * 8 : count ? g.seg_rec * 4 : 0);
    g.warp_bytes = ((g.warp_bytes + 127) / 128) * 128;
    const size_t tables = (size_t)g.sw_rows * 8;
    const size_t avail = 227 * 1024 - 1024;
    if (tables + 256 >= avail) return false;
    int nw = (int)((avail - tables - kSymMaxWarps * 16) / (size_t)g.warp_bytes);
    if (const char *e = getenv("SMC_SYM_NWARPS")) nw = std::min(nw, std::max(1, atoi(e)));  // tuning knob
    g.nwarps = std::min(nw, kSymMaxWarps);
    if (g.nwarps < 1) return false;
    smem = (size_t)g.nwarps * g.warp_bytes + tables + (size_t)g.nwarps * 16;
    // Units: runs of u_big tiles for the upper part of the rows, single tiles for the rest (the tail of the work queue).
    // Longer units leave fewer partial sums for the gather kernel (a row of a strip is written once per unit that streams
    // it); single tiles balance best.  With s tiles per resident warp, every warp takes floor(s / u_big) long units off the
    // queue and the remainder goes out tile by tile, so the makespan stays ceil(s) tiles whatever u_big <= s / 2 is: take the
    // longest run up to 4 tiles.  (4K on one GPU: s = 37, 4-tile units and 4 % single tiles; a 270-row band of an 8-GPU run:
    // s = 5, 2-tile units and 20 % single tiles.)
    const long long tiles = (long long)g.n_trows * g.n_strips * p.ptr_count;
    const long long workers = (long long)std::max(p.sm_count, 1) * g.nwarps;
    const double share = (double)tiles / (double)workers;
    int u_big = std::max(1, std::min(4, (int)(share / 2.0)));
    const long long big_tiles = (long long)(share / u_big) * u_big * workers;  // what the long units may cover
    int small_pct = (int)std::min<long long>(100, std::max<long long>(3, 100 - big_tiles * 100 / std::max<long long>(tiles, 1) + 1));
    int u_small = 1;
    if (const char *e = getenv("SMC_SYM_UNIT")) sscanf(e, "%d,%d,%d", &u_big, &u_small, &small_pct);  // tuning knob
    u_big = std::max(1, u_big);
    u_small = std::max(1, std::min(u_small, u_big));
    g.u_big = u_big;
    g.u_small = u_small;
    const int small_rows = (int)((long long)g.n_trows * small_pct / 100);
    g.n_big = (g.n_trows - small_rows) / u_big;
    const int rest = g.n_trows - g.n_big * u_big;
    const int n_small = (rest + u_small - 1) / u_small;
    g.n_units_y = g.n_big + n_small;
    g.scratch_rows = g.n_big * (2 * u_big + r) + n_small * (2 * u_small + r);
    const long long total = (long long)g.n_units_y * g.n_strips * p.ptr_count;
    if (total <= 0 || total > 0x3fffffffLL) return false;
    g.units_total = (int)total;
    return true;
}

size_t smc_filter_sym_scratch_elems(const SmcFilterParams &p, const SmcSymParams &g) {
    return (size_t)p.ptr_count * g.n_strips * g.scratch_rows * g.seg_rec;
}

int smc_launch_filter_sym(smc_context *ctx, const SmcFilterParams &p, const SmcSymParams &g, size_t smem, const char **name) {
    const int rows = p.row_end - p.row_begin;
    if (rows <= 0) return SMC_OK;
    static thread_local char nm[80];
    snprintf(nm, sizeof(nm), "sym-warp<%sNG=%d,PY=2,welch,W=%d,U=%d/%d>", g.tri ? "C=1x3," : p.C == 1 ? "C=1," : "", p.NG, g.nwarps,
             g.u_big, g.u_small);
    if (name) *name = nm;
    const int grid = (int)std::min<long long>((g.units_total + g.nwarps - 1) / g.nwarps, (long long)ctx->sm_count);
    SMC_CUDA(cudaMemsetAsync(g.unit_counter, 0, sizeof(int), ctx->stream));
    const int rc = (p.C == 3 || g.tri) ? launch_sym_c<3>(ctx, p, g, smem, grid) : launch_sym_c<1>(ctx, p, g, smem, grid);
    if (rc) return rc;
    const dim3 gb(128), gg((p.W + 127) / 128, rows, p.ptr_count);
    if (g.tri) {
        sym_gather_kernel<3, false, true><<<gg, gb, 0, ctx->stream>>>(p, g);
    } else if (p.C == 3) {
        if (p.accepted != nullptr) sym_gather_kernel<3, true, false><<<gg, gb, 0, ctx->stream>>>(p, g);
        else sym_gather_kernel<3, false, false><<<gg, gb, 0, ctx->stream>>>(p, g);
    } else {
        if (p.accepted != nullptr) sym_gather_kernel<1, true, false><<<gg, gb, 0, ctx->stream>>>(p, g);
        else sym_gather_kernel<1, false, false><<<gg, gb, 0, ctx->stream>>>(p, g);
    }
    SMC_CHECK_LAUNCH(ctx);
    return SMC_OK;
}
