// microbench.cu -- issue/pipe throughput of the instructions the streaming filter kernel is built from, on the
// device it runs on: FFMA, FADD, FFMA2/FADD2/FMUL2 (packed fp32x2), FSETP, MUFU.EX2, LDS.128.
// Prints lane-ops (or instructions) per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITER 4096
#define CHAINS 8

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

template <int MODE>
__global__ void k(float *out, float seed, long long *cyc) {
    float a[CHAINS];
    unsigned long long p[CHAINS];
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + i + threadIdx.x; p[i] = pk(a[i], a[i] + 0.5f); }
    const unsigned long long c2 = pk(seed, seed * 0.5f), m2 = pk(1.0001f, 0.9999f);
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    int cnt = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) a[i] = fmaf(a[i], 1.0001f, seed);                      // FFMA (imm form possible)
            if (MODE == 1) a[i] = fmaf(a[i], a[(i + 1) % CHAINS], seed);          // FFMA 3-reg
            if (MODE == 2) a[i] = a[i] + seed;                                    // FADD
            if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
            if (MODE == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
            if (MODE == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(m2));
            if (MODE == 6) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i])); a[i] = y; }
            if (MODE == 7) { cnt += (a[i] <= seed + it) ? 1 : 0; }                 // FSETP + integer add
            if (MODE == 8) { float4 v = sm[(threadIdx.x * 9 + i * 37 + it) & 2047]; a[i] += v.x; }  // LDS.128 (+FADD)
            if (MODE == 9) {   // mix: 1 FFMA2 + 1 FFMA + 1 FSETP-ish per chain step (does packed issue overlap scalar?)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
                a[i] = fmaf(a[i], 1.0001f, seed);
            }
            if (MODE == 10) {  // FADD2 + FADD
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
                a[i] = a[i] + seed;
            }
        }
    }
    long long t1 = clock64();
    float s = cnt;
    for (int i = 0; i < CHAINS; i++) { s += a[i]; s += __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double lanes_per_instr, int threads, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm;
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<MODE><<<blocks, threads, 32768>>>(out, 1.0f, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads, 32768>>>(out, 1.0f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096]; cudaMemcpy(h, cyc, sizeof(long long) * (blocks < 4096 ? blocks : 4096), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks && i < 4096; i++) avg += h[i]; avg /= (blocks < 4096 ? blocks : 4096);
    double instr_per_sm = (double)ITER * CHAINS * threads * blocks_per_sm;   // thread-instructions per SM
    printf("%-28s threads/SM=%4d  thread-instr/clk/SM=%7.1f  lane-ops/clk/SM=%7.1f  (%.3f ms, %.0f clk)\n", name,
           threads * blocks_per_sm, instr_per_sm / avg, instr_per_sm * lanes_per_instr / avg, ms, avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int bps = 1; bps <= 2; bps++) {
        int th = 512;
        run<0>("FFMA imm", 1, th, bps);
        run<1>("FFMA 3-reg", 1, th, bps);
        run<2>("FADD", 1, th, bps);
        run<3>("FFMA2 (f32x2)", 2, th, bps);
        run<4>("FADD2 (f32x2)", 2, th, bps);
        run<5>("FMUL2 (f32x2)", 2, th, bps);
        run<6>("MUFU.EX2", 1, th, bps);
        run<7>("FSETP+IADD", 1, th, bps);
        run<8>("LDS.128+FADD", 1, th, bps);
        run<9>("FFMA2 + FFMA pair", 3, th, bps);
        run<10>("FADD2 + FADD pair", 3, th, bps);
    }
    return 0;
}
