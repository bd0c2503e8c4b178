/*
 * ref_harness.cu -- C-ABI shim around the REFERENCE's own CUDA denoiser, compiled unmodified.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/statmc_oracle.c header).  This file is ours; the code it
 * calls is /root/reference/src/ext/opencv_contrib/modules/cudaimgproc/src/cuda/stat_denoiser.cu, compiled
 * where it lies by oracle/Makefile into oracle/_ref/ (git-ignored).  No reference source is copied.
 *
 * It does what Estimator::AllocateBuffers (src/statistics/estimator.cpp:35-84, 271-288) and the sample
 * (src/ext/opencv_contrib/samples/stat_denoiser/main.cpp:30-44) do with cv::cuda::GpuMat: build device-resident
 * 1xN tables of PtrStepSzb descriptors {data, step, cols, rows}, a 1xG uchar channel-count array and a Gx1 float
 * factor array, then call cv::cuda::device::imgproc::stat_denoiser::filter<T> (stat_denoiser.cu:397-475).
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "opencv2/core/cuda_types.hpp"

using cv::cuda::PtrStepSzb;

// Declarations of the reference's device-layer entry points (defined in stat_denoiser.cu:352-483).
namespace cv { namespace cuda { namespace device { namespace imgproc { namespace stat_denoiser {
void setup();
void synchronize(cudaStream_t stream);
template <typename T>
void calculate_mean_vars(const unsigned short ptrCount, const unsigned short width, const unsigned short height,
                         const PtrStepSzb &nPtrs, const PtrStepSzb &m2Ptrs, PtrStepSzb meanVarPtrs,
                         cudaStream_t stream);
template <typename T>
void filter(const unsigned short ptrCount, const unsigned short width, const unsigned short height,
            const float dSFactor, const unsigned char radius, const bool denoiseFilm, const PtrStepSzb &nPtrs,
            const PtrStepSzb &meanPtrs, const PtrStepSzb &m2Ptrs, const PtrStepSzb &m3Ptrs,
            const PtrStepSzb &filmPtrs, const PtrStepSzb &film, const PtrStepSzb &gBufPtrs,
            const PtrStepSzb &gBufChannelCounts, const PtrStepSzb &gBufDRFactors, const unsigned char nGBufs,
            PtrStepSzb meanCorrPtrs, PtrStepSzb discriminatorPtrs, PtrStepSzb filmFilteredPtrs,
            PtrStepSzb filmFiltered, cudaStream_t stream);
}}}}}  // namespace cv::cuda::device::imgproc::stat_denoiser

// The only non-CUDA-runtime symbol stat_denoiser.cu needs (cudaSafeCall -> cv::error, common.hpp:66-76).
namespace cv {
void error(int code, const std::string &err, const char *func, const char *file, int line) {
    std::fprintf(stderr, "[statmc ref] cv::error %d: %s in %s (%s:%d)\n", code, err.c_str(), func, file, line);
    std::abort();
}
}  // namespace cv

namespace ref = cv::cuda::device::imgproc::stat_denoiser;

extern "C" {

typedef struct smr_plane {
    void *dev;     // device pointer of row 0
    size_t step;   // bytes between rows
} smr_plane;

struct smr_filter {
    int channels, ptr_count, W, H, radius, denoise_film, n_gbufs;
    float ds_factor;
    void *d_tables[9] = {};  // n, mean, m2, m3, filmPtrs, gbufs, meanCorr, disc, filmFiltered
    unsigned char *d_gch = nullptr;
    float *d_gf = nullptr;
    PtrStepSzb film, film_filtered;
};

static void *upload_table(const smr_plane *planes, int count, int W, int H) {
    std::vector<PtrStepSzb> h(count > 0 ? count : 1);
    for (int i = 0; i < count; i++) h[i] = PtrStepSzb(H, W, (unsigned char *)planes[i].dev, planes[i].step);
    void *d = nullptr;
    cudaMalloc(&d, h.size() * sizeof(PtrStepSzb));
    cudaMemcpy(d, h.data(), h.size() * sizeof(PtrStepSzb), cudaMemcpyHostToDevice);
    return d;
}

static PtrStepSzb table_desc(void *d, int count) {
    // a 1xN GpuMat of CV_8UC(sizeof(PtrStepSzb)) converted to PtrStepSzb: rows=1, cols=N (estimator.cpp:41-44)
    return PtrStepSzb(1, count, (unsigned char *)d, (size_t)count * sizeof(PtrStepSzb));
}

void smr_setup(void) { ref::setup(); }

// channels: 1 -> filter<float>, 3 -> filter<float3>.  Plane arrays have ptr_count entries (gbufs: n_gbufs).
// film / film_filtered may have dev == NULL when denoise_film == 0.
smr_filter *smr_filter_create(int channels, int ptr_count, int W, int H, float ds_factor, int radius,
                              int denoise_film, const smr_plane *n, const smr_plane *mean, const smr_plane *m2,
                              const smr_plane *m3, const smr_plane *film_ptrs, smr_plane film,
                              const smr_plane *gbufs, const unsigned char *gbuf_channels,
                              const float *gbuf_dr_factors, int n_gbufs, const smr_plane *mean_corr,
                              const smr_plane *disc, const smr_plane *film_filtered_ptrs,
                              smr_plane film_filtered) {
    smr_filter *f = new smr_filter;
    f->channels = channels; f->ptr_count = ptr_count; f->W = W; f->H = H; f->radius = radius;
    f->denoise_film = denoise_film; f->n_gbufs = n_gbufs; f->ds_factor = ds_factor;
    const smr_plane *src[9] = {n, mean, m2, m3, film_ptrs, gbufs, mean_corr, disc, film_filtered_ptrs};
    for (int i = 0; i < 9; i++) f->d_tables[i] = upload_table(src[i], i == 5 ? n_gbufs : ptr_count, W, H);
    cudaMalloc(&f->d_gch, n_gbufs > 0 ? n_gbufs : 1);
    cudaMalloc(&f->d_gf, sizeof(float) * (n_gbufs > 0 ? n_gbufs : 1));
    if (n_gbufs > 0) {
        cudaMemcpy(f->d_gch, gbuf_channels, n_gbufs, cudaMemcpyHostToDevice);
        cudaMemcpy(f->d_gf, gbuf_dr_factors, sizeof(float) * n_gbufs, cudaMemcpyHostToDevice);
    }
    f->film = PtrStepSzb(H, W, (unsigned char *)film.dev, film.step);
    f->film_filtered = PtrStepSzb(H, W, (unsigned char *)film_filtered.dev, film_filtered.step);
    return f;
}

int smr_filter_run(smr_filter *f, cudaStream_t stream) {
    const PtrStepSzb gch(1, f->n_gbufs, f->d_gch, (size_t)f->n_gbufs);
    const PtrStepSzb gf(f->n_gbufs, 1, (unsigned char *)f->d_gf, sizeof(float));
    const PtrStepSzb t[9] = {table_desc(f->d_tables[0], f->ptr_count), table_desc(f->d_tables[1], f->ptr_count),
                             table_desc(f->d_tables[2], f->ptr_count), table_desc(f->d_tables[3], f->ptr_count),
                             table_desc(f->d_tables[4], f->ptr_count), table_desc(f->d_tables[5], f->n_gbufs),
                             table_desc(f->d_tables[6], f->ptr_count), table_desc(f->d_tables[7], f->ptr_count),
                             table_desc(f->d_tables[8], f->ptr_count)};
    if (f->channels == 3)
        ref::filter<float3>(f->ptr_count, f->W, f->H, f->ds_factor, f->radius, f->denoise_film != 0, t[0], t[1], t[2],
                            t[3], t[4], f->film, t[5], gch, gf, f->n_gbufs, t[6], t[7], t[8], f->film_filtered,
                            stream);
    else
        ref::filter<float>(f->ptr_count, f->W, f->H, f->ds_factor, f->radius, f->denoise_film != 0, t[0], t[1], t[2],
                           t[3], t[4], f->film, t[5], gch, gf, f->n_gbufs, t[6], t[7], t[8], f->film_filtered,
                           stream);
    return (int)cudaGetLastError();
}

void smr_filter_destroy(smr_filter *f) {
    if (!f) return;
    for (void *d : f->d_tables) cudaFree(d);
    cudaFree(f->d_gch);
    cudaFree(f->d_gf);
    delete f;
}

// calculate_mean_vars<T> (stat_denoiser.cu:361-388) on single planes
int smr_calculate_mean_vars(int channels, int W, int H, smr_plane n, smr_plane m2, smr_plane out,
                            cudaStream_t stream) {
    void *tn = upload_table(&n, 1, W, H), *tm = upload_table(&m2, 1, W, H), *to = upload_table(&out, 1, W, H);
    if (channels == 3)
        ref::calculate_mean_vars<float3>(1, W, H, table_desc(tn, 1), table_desc(tm, 1), table_desc(to, 1), stream);
    else
        ref::calculate_mean_vars<float>(1, W, H, table_desc(tn, 1), table_desc(tm, 1), table_desc(to, 1), stream);
    cudaStreamSynchronize(stream);
    int e = (int)cudaGetLastError();
    cudaFree(tn); cudaFree(tm); cudaFree(to);
    return e;
}

// ---------------------------------------------------------------------------------------------------------------
// Self-contained timing of the reference path for bench.py --impl reference: allocation (cudaMallocPitch, what
// GpuMat(rows, cols, type) does, gpu_mat.cu:112-123), copies (cudaMemcpy2DAsync, gpu_mat.cu:224-234) and the three
// reference kernels, all with plain CUDA runtime calls on one stream -- nothing of libstatmc_b200 is involved.
//   host planes: n (int32, W), mean/m2/m3/film/normal/albedo (float32, 3W), tightly packed rows
//   kernel_ms:   CUDA events around `steps` x filter<float3> (johnson + discriminator + filter kernels), device-resident
//   e2e_ms[0]:   Upload(); Denoise(); Download(); Synchronize() per step as Estimator does (statpath.cpp:406-418),
//                from the host memory it was given (pageable cv::Mat memory in the reference)
//   e2e_ms[1]:   the same with the host planes page-locked first (cudaHostRegister): separates the copy policy from the
//                kernels' speed
// Returns 0, or the first CUDA error.
// ---------------------------------------------------------------------------------------------------------------
int smr_bench_rgb(int W, int H, int radius, float ds_factor, const float *gbuf_dr_factors, const int *h_n,
                  const float *h_mean, const float *h_m2, const float *h_m3, const float *h_film, const float *h_normal,
                  const float *h_albedo, float *h_out, int steps, int warmup, double *kernel_ms, double *e2e_ms) {
#define SMR_CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { std::fprintf(stderr, "[statmc ref] %s: %s\n", #x, cudaGetErrorString(e__)); return (int)e__; } } while (0)
    ref::setup();
    cudaStream_t s;
    SMR_CK(cudaStreamCreate(&s));
    const void *host[7] = {h_n, h_mean, h_m2, h_m3, h_film, h_normal, h_albedo};
    const size_t px[7] = {4, 12, 12, 12, 12, 12, 12};
    smr_plane d[7], d_mc, d_disc, d_dummy, d_out;
    for (int i = 0; i < 7; i++) SMR_CK(cudaMallocPitch(&d[i].dev, &d[i].step, (size_t)W * px[i], H));
    SMR_CK(cudaMallocPitch(&d_mc.dev, &d_mc.step, (size_t)W * 12, H));
    SMR_CK(cudaMallocPitch(&d_disc.dev, &d_disc.step, (size_t)W * 12, H));
    SMR_CK(cudaMallocPitch(&d_dummy.dev, &d_dummy.step, (size_t)W * 12, H));
    SMR_CK(cudaMallocPitch(&d_out.dev, &d_out.step, (size_t)W * 12, H));
    const smr_plane g[2] = {d[5], d[6]};
    const unsigned char gch[2] = {3, 3};
    smr_filter *f = smr_filter_create(3, 1, W, H, ds_factor, radius, 1, &d[0], &d[1], &d[2], &d[3], &d[4], d[4], g, gch,
                                      gbuf_dr_factors, 2, &d_mc, &d_disc, &d_dummy, d_out);
    auto upload = [&]() -> cudaError_t {
        for (int i = 0; i < 7; i++) {
            cudaError_t e = cudaMemcpy2DAsync(d[i].dev, d[i].step, host[i], (size_t)W * px[i], (size_t)W * px[i], H,
                                              cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    auto download = [&]() { return cudaMemcpy2DAsync(h_out, (size_t)W * 12, d_out.dev, d_out.step, (size_t)W * 12, H, cudaMemcpyDeviceToHost, s); };
    SMR_CK(upload());
    SMR_CK(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1;
    SMR_CK(cudaEventCreate(&e0));
    SMR_CK(cudaEventCreate(&e1));
    float ms = 0.f;
    for (int i = 0; i < warmup; i++) SMR_CK((cudaError_t)smr_filter_run(f, s));
    SMR_CK(cudaStreamSynchronize(s));
    SMR_CK(cudaEventRecord(e0, s));
    for (int i = 0; i < steps; i++) SMR_CK((cudaError_t)smr_filter_run(f, s));
    SMR_CK(cudaEventRecord(e1, s));
    SMR_CK(cudaStreamSynchronize(s));
    SMR_CK(cudaEventElapsedTime(&ms, e0, e1));
    *kernel_ms = (double)ms / steps;
    for (int pinned = 0; pinned < 2; pinned++) {
        if (pinned) {
            for (int i = 0; i < 7; i++) SMR_CK(cudaHostRegister((void *)host[i], (size_t)W * px[i] * H, cudaHostRegisterDefault));
            SMR_CK(cudaHostRegister(h_out, (size_t)W * 12 * H, cudaHostRegisterDefault));
        }
        auto e2e = [&]() -> cudaError_t {
            cudaError_t e = upload();
            if (e != cudaSuccess) return e;
            if ((e = (cudaError_t)smr_filter_run(f, s)) != cudaSuccess) return e;
            if ((e = download()) != cudaSuccess) return e;
            return cudaStreamSynchronize(s);
        };
        for (int i = 0; i < (warmup < 2 ? warmup : 2); i++) SMR_CK(e2e());
        SMR_CK(cudaEventRecord(e0, s));
        for (int i = 0; i < steps; i++) SMR_CK(e2e());
        SMR_CK(cudaEventRecord(e1, s));
        SMR_CK(cudaStreamSynchronize(s));
        SMR_CK(cudaEventElapsedTime(&ms, e0, e1));
        e2e_ms[pinned] = (double)ms / steps;
        if (pinned) {
            for (int i = 0; i < 7; i++) cudaHostUnregister((void *)host[i]);
            cudaHostUnregister(h_out);
        }
    }
    smr_filter_destroy(f);
    for (int i = 0; i < 7; i++) cudaFree(d[i].dev);
    cudaFree(d_mc.dev); cudaFree(d_disc.dev); cudaFree(d_dummy.dev); cudaFree(d_out.dev);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    return 0;
#undef SMR_CK
}

int smr_synchronize(cudaStream_t stream) {
    ref::synchronize(stream);
    return (int)cudaGetLastError();
}

}  // extern "C"
