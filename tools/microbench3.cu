// microbench3.cu -- issue rate of scalar vs packed fp32 ops on sm_100a: warp-instructions per cycle per SMSP for
// FFMA, FFMA2, FADD2, FMUL2 and mixes, with independent dependency chains (no memory traffic).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ex2(float a) { float d; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }
__device__ __forceinline__ bool setp(float a, float b) { int p; asm volatile("{ .reg .pred q; setp.le.f32 q, %1, %2; selp.s32 %0, 1, 0, q; }" : "=r"(p) : "f"(a), "f"(b)); return p; }

// MODE: 0 = 16 FFMA, 1 = 16 FFMA2, 2 = 8 FFMA + 8 FFMA2, 3 = 16 FADD2, 4 = 8 FFMA2 + 8 FADD (scalar), 5 = 16 FADD,
//       6 = 8 FFMA2 + 8 FFMA + 2 MUFU, 7 = 12 FFMA + 4 MUFU
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, int iters, long long *cyc) {
    float s[16]; unsigned long long p[16];
    for (int i = 0; i < 16; i++) { s[i] = threadIdx.x * 0.001f + i; p[i] = ((unsigned long long)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] * 0.5f); }
    const float a = 1.0001f, b = 0.0001f;
    const unsigned long long pa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a), pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) s[i] = fma1(s[i], a, b);
            if (MODE == 1) p[i] = fma2(p[i], pa, pb);
            if (MODE == 2) { if (i & 1) s[i] = fma1(s[i], a, b); else p[i] = fma2(p[i], pa, pb); }
            if (MODE == 3) p[i] = add2(p[i], pb);
            if (MODE == 4) { if (i & 1) s[i] = add1(s[i], b); else p[i] = fma2(p[i], pa, pb); }
            if (MODE == 5) s[i] = add1(s[i], b);
            if (MODE == 6) { if (i & 1) s[i] = fma1(s[i], a, b); else p[i] = fma2(p[i], pa, pb); if (i == 3 || i == 11) s[i] = ex2(s[i]); }
            if (MODE == 7) { if ((i & 3) == 3) s[i] = ex2(s[i]); else s[i] = fma1(s[i], a, b); }
        }
    }
    long long t1 = clock64();
    float o = 0; for (int i = 0; i < 16; i++) o += s[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = o;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ninstr, int ctas_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * ctas_per_sm, iters = 20000;
    float *out; long long *cyc; cudaMalloc(&out, 4 * blocks * 128); cudaMalloc(&cyc, 8 * blocks);
    for (int rep = 0; rep < 2; rep++) { k<MODE><<<blocks, 128>>>(out, iters, cyc); cudaDeviceSynchronize(); }
    static long long hc[8192]; cudaMemcpy(hc, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += hc[i]; avg /= blocks;
    printf("%-34s warps/SMSP=%d  warp-instr per cycle per SMSP = %5.3f\n", name, ctas_per_sm, (double)iters * ninstr * ctas_per_sm / avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w = 1; w <= 4; w += 3) {
        run<0>("16 FFMA", 16, w);
        run<1>("16 FFMA2", 16, w);
        run<2>("8 FFMA + 8 FFMA2", 16, w);
        run<3>("16 FADD2", 16, w);
        run<4>("8 FFMA2 + 8 FADD", 16, w);
        run<5>("16 FADD", 16, w);
        run<6>("8 FFMA2 + 8 FFMA + 2 MUFU", 18, w);
        run<7>("12 FFMA + 4 MUFU", 16, w);
    }
    return 0;
}
