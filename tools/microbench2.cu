// microbench2.cu -- the per-pair instruction mix of the streaming filter kernel with operands in registers
// (no shared/global memory): what the SM can sustain for exactly this mix, as cycles per pair-warp per SMSP.
// Variants isolate which pipe binds: full mix, without MUFU, without FSETP, scalar-only (no packed ops).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra = *(unsigned long long *)&a, rb = *(unsigned long long *)&b, rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *(float2 *)&rd;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long ra = *(unsigned long long *)&a, rb = *(unsigned long long *)&b, rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *(float2 *)&rd;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *(unsigned long long *)&a, rb = *(unsigned long long *)&b, rc = *(unsigned long long *)&c, rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *(float2 *)&rd;
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct Cen { float2 t01, d01, g0, g1, g2; float tz, dz, nc; };
struct Rec { float2 m01, d01, g0, g1, g2, v01; float mz, dz, vz; };
struct Acc { float n0, n1, n2, den; };

template <int MODE>
__device__ __forceinline__ void pair(const Cen &c, const Rec &r, float sw, Acc &a) {
    bool ok = true;
    if (MODE == 5 || MODE == 6 || MODE == 7) {  // FMA-form membership: fma(-tC, mI, dI) <= -dC
        ok = (fmaf(c.t01.x, r.m01.x, r.d01.x) <= c.d01.x) & (fmaf(c.t01.y, r.m01.y, r.d01.y) <= c.d01.y) &
             (fmaf(c.tz, r.mz, r.dz) <= c.dz);
    } else if (MODE != 2) {   // membership
        if (MODE == 3) {  // scalar only
            ok = (c.d01.x + r.d01.x <= c.t01.x * r.m01.x) & (c.d01.y + r.d01.y <= c.t01.y * r.m01.y) & (c.dz + r.dz <= c.tz * r.mz);
        } else {
            float2 s = add2(c.d01, r.d01), p = mul2(c.t01, r.m01);
            ok = (s.x <= p.x) & (s.y <= p.y) & (__fadd_rn(c.dz, r.dz) <= __fmul_rn(c.tz, r.mz));
        }
    }
    float acc;
    if (MODE == 3) {
        float e0 = r.g0.x + c.g0.x, e1 = r.g0.y + c.g0.y, e2 = r.g1.x + c.g1.x, e3 = r.g1.y + c.g1.y, e4 = r.g2.x + c.g2.x, e5 = r.g2.y + c.g2.y;
        float l = e0 * e0, h = e1 * e1;
        l = fmaf(e2, e2, l); h = fmaf(e3, e3, h); l = fmaf(e4, e4, l); h = fmaf(e5, e5, h);
        acc = l + h;
    } else if (MODE == 4 || MODE == 6 || MODE == 7) {
        // dot form: |gI|^2 (record) - 2 gI.gC + |gC|^2: acc = sum gI * (-2 gC) [+ nI + nC folded into sw]
        float2 q = mul2(r.g0, c.g0);
        q = fma2(r.g1, c.g1, q);
        q = fma2(r.g2, c.g2, q);
        acc = __fadd_rn(__fadd_rn(q.x, q.y), r.vz * 0.f + r.dz);  // + nI (a record slot)
    } else {
        float2 e = add2(r.g0, c.g0);
        float2 q = mul2(e, e);
        e = add2(r.g1, c.g1); q = fma2(e, e, q);
        e = add2(r.g2, c.g2); q = fma2(e, e, q);
        acc = __fadd_rn(q.x, q.y);
    }
    float arg = __fsub_rn(sw, acc);
    float w = (MODE == 1) ? arg * 0.5f : ex2(arg);  // MODE 1: no MUFU
    if (MODE == 7) {
        const float wm = ok ? w : 0.f;
        float2 n01 = fma2(make_float2(wm, wm), r.v01, make_float2(a.n0, a.n1));
        float2 n2d = fma2(make_float2(wm, wm), make_float2(r.vz, 1.f), make_float2(a.n2, a.den));
        a.n0 = n01.x; a.n1 = n01.y; a.n2 = n2d.x; a.den = n2d.y;
    } else if (ok) {
        a.n0 = fmaf(w, r.v01.x, a.n0); a.n1 = fmaf(w, r.v01.y, a.n1); a.n2 = fmaf(w, r.vz, a.n2); a.den += w;
    }
}

template <int MODE, int NC>
__global__ void __launch_bounds__(128) k(float *out, const float *in, int iters, long long *cyc) {
    Cen c[NC]; Acc a[NC];
    const float s = in[threadIdx.x & 31];
    for (int i = 0; i < NC; i++) {
        c[i].t01 = make_float2(s + i, s * 2 + i); c[i].d01 = make_float2(-s, -s - i); c[i].tz = s + 3; c[i].dz = -s * 3;
        c[i].g0 = make_float2(s + 0.01f * i, -s); c[i].g1 = make_float2(s * .5f, s * .25f - 0.02f * i); c[i].g2 = make_float2(-s * .5f + 0.03f * i, s * .125f); c[i].nc = s * i;
        a[i] = {0, 0, 0, 0};
    }
    // records live in shared memory and are re-read every iteration at a varying index (4 x LDS.128 per record, as
    // in the real kernel), so nothing of the pair evaluation is loop-invariant
    __shared__ float4 recs[256 * 4];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x)
        recs[i] = make_float4(s + 0.001f * i, s * 0.1f + 0.002f * i, 0.1f + 0.0003f * i, 0.2f + 0.0001f * i);
    __syncthreads();
    float sw = -s;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const float4 *q = recs + (((it * 5 + (threadIdx.x >> 5)) & 255) << 2);  // warp-uniform address: broadcast, no bank conflicts
        const float4 c0 = q[0], c1 = q[1], c2 = q[2], c3 = q[3];
        Rec r;
        r.m01 = make_float2(c0.x, c0.y); r.d01 = make_float2(c0.z, c0.w); r.mz = c1.x; r.dz = c1.y; r.vz = c1.z;
        r.v01 = make_float2(c2.x, c2.y); r.g0 = make_float2(c2.z, c2.w); r.g1 = make_float2(c3.x, c3.y);
        r.g2 = make_float2(c3.z, c3.w);
        sw -= 1e-4f;
#pragma unroll
        for (int i = 0; i < NC; i++) pair<MODE>(c[i], r, sw, a[i]);
    }
    long long t1 = clock64();
    float o = 0;
    for (int i = 0; i < NC; i++) o += a[i].n0 + a[i].n1 + a[i].n2 + a[i].den;
    out[blockIdx.x * blockDim.x + threadIdx.x] = o;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int NC>
void run(const char *name, int ctas_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * ctas_per_sm, iters = 20000;
    float *out, *in; long long *cyc;
    cudaMalloc(&out, 4 * blocks * 128); cudaMalloc(&in, 4 * 32); cudaMalloc(&cyc, 8 * blocks);
    float h[32]; for (int i = 0; i < 32; i++) h[i] = 1.f + i * 0.01f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    k<MODE, NC><<<blocks, 128>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    k<MODE, NC><<<blocks, 128>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    static long long hc[8192]; cudaMemcpy(hc, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += hc[i]; avg /= blocks;
    // each SMSP hosts ctas_per_sm warps (4 warps per CTA spread over 4 SMSPs); pairs per warp = iters * NC
    double cyc_per_pair_warp = avg / ((double)iters * NC * ctas_per_sm);
    printf("%-34s NC=%d warps/SMSP=%d  cycles per pair-warp per SMSP = %6.2f\n", name, NC, ctas_per_sm, cyc_per_pair_warp);
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
    for (int w = 1; w <= 4; w++) {
        run<0, 8>("full mix (packed)", w);
        run<1, 8>("no MUFU", w);
        run<2, 8>("no membership (no FSETP)", w);
        run<3, 8>("scalar only (no f32x2)", w);
        run<4, 8>("dot-form weight", w);
        run<5, 8>("fma membership", w);
        run<6, 8>("dot-form + fma membership", w);
        run<7, 8>("dot + fma memb + packed accum", w);
        run<0, 4>("full mix (packed)", w);
        run<6, 4>("dot-form + fma membership", w);
    }
    return 0;
}
