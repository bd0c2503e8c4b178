// statmc_b200.hpp -- header-only C++ host layer over the C ABI (statmc_b200.h): the B200 mirror of the part of
// StatMC's src/statistics/ that sits on the data-parallel hot path.  Same names, argument meaning and error behaviour as
// the reference, so that the integrator-side code (StatPathIntegrator::Render / ::Denoise, statpath.cpp:118-550) reads
// the same; no OpenCV, no CUDA headers, plain C++17 + libstatmc_b200.so.
//
//   reference (file:line)                                             here
//   cv::cuda::PtrStepSzb                 cuda_types.hpp:103-135       statmc::PtrStepSzb        (same 24-byte layout)
//   cv::cuda::Stream + setup()           estimator.h:280,326          statmc::Stream            (owns an smc_context)
//   cv::Mat (rows, cols, CV_32FC3 ...)   estimator.cpp:121-146        statmc::Mat               (pinned host memory)
//   cv::cuda::GpuMat                     buffer.h:24-63               statmc::GpuMat            (ref-counted smc_buffer)
//   pbrt::Buffer                         buffer.h:19-71               statmc::Buffer            (name, mat, matPtr, gpuMat, upload, download)
//   cv::cuda::stat_denoiser::*           cudaimgproc.hpp:736-799      statmc::stat_denoiser::*  (setup, synchronize, calculateMeanVars<T>, filter<T>)
//   pbrt::StatTypeConfig(s)              estimator.h:71-101           statmc::StatTypeConfig(s) (same fields)
//   pbrt::Estimator                      estimator.h:241-379          statmc::Estimator         (RegisterGBuffer, AllocateBuffers, Upload,
//                                                                     Denoise, Download, CalculateMeanVars, Synchronize, public buffer vectors)
//   StatTile<T>::Add*Sample* + Estimator::Merge*Tiles  estimator.h:162-232, estimator.cpp:341-407
//                                                                     Estimator::AddSamples (batch of samples -> planes resident in HBM)
// Errors: the reference turns every CUDA failure into a thrown cv::Exception (common.hpp:66-76); here every non-zero status
// of the C ABI is thrown as statmc::Exception carrying smc_last_error().
#ifndef STATMC_B200_HPP
#define STATMC_B200_HPP

#include <algorithm>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "statmc_b200.h"

namespace statmc {

class Exception : public std::runtime_error {
  public:
    Exception(int code, const std::string &what) : std::runtime_error(what), code(code) {}
    int code;
};

inline void check(int rc) {
    if (rc != SMC_OK) throw Exception(rc, smc_last_error());
}

// The host side names float3 through a local struct (estimator.cpp:8-10, samples/stat_denoiser/main.cpp:14-16).
struct float3 {
    float x, y, z;
};

// cv::cuda::PtrStepSzb (cuda_types.hpp:103-135): what the device-resident descriptor tables hold.
struct PtrStepSzb {
    unsigned char *data = nullptr;
    size_t step = 0;
    int cols = 0;
    int rows = 0;
};
static_assert(sizeof(PtrStepSzb) == 24, "must match cv::cuda::PtrStepSzb");

enum Depth { S32 = SMC_I32, F32 = SMC_F32 };  // CV_32S / CV_32F: the only depths the hot path uses

// cv::cuda::Stream + cv::cuda::stat_denoiser::setup(): one context (device, CUDA stream, t-quantile table).
class Stream {
  public:
    explicit Stream(int device = 0) { check(smc_context_create(device, &ctx_)); }
    Stream(int device, void *cudaStream) { check(smc_context_create_on_stream(device, cudaStream, &ctx_)); }
    ~Stream() { smc_context_destroy(ctx_); }
    Stream(const Stream &) = delete;
    Stream &operator=(const Stream &) = delete;
    void waitForCompletion() { check(smc_synchronize(ctx_)); }
    smc_context *ctx() const { return ctx_; }
    void *cudaStream() const { return smc_context_stream(ctx_); }

  private:
    smc_context *ctx_ = nullptr;
};

// Host matrix, the part of cv::Mat the hot path needs: rows x cols x channels of int32 / float32, row-major interleaved,
// shared ownership.  Memory is pinned (cudaMallocHost) so that uploads and downloads are asynchronous.
class Mat {
  public:
    Mat() = default;
    Mat(int rows, int cols, int channels, Depth depth = F32) : rows(rows), cols(cols), channels_(channels), depth_(depth) {
        void *p = nullptr;
        check(smc_host_alloc(bytes(), &p));
        std::memset(p, 0, bytes());
        mem_ = std::shared_ptr<void>(p, [](void *q) { smc_host_free(q); });
    }
    // borrowed memory (e.g. an existing cv::Mat's data pointer); the caller keeps it alive
    Mat(int rows, int cols, int channels, Depth depth, void *data, size_t step)
        : rows(rows), cols(cols), channels_(channels), depth_(depth), step_(step), mem_(data, [](void *) {}) {}
    int channels() const { return channels_; }
    Depth depth() const { return depth_; }
    size_t step() const { return step_ ? step_ : (size_t)cols * channels_ * 4; }
    size_t bytes() const { return step() * (size_t)rows; }
    bool empty() const { return !mem_; }
    unsigned char *ptr() const { return (unsigned char *)mem_.get(); }
    template <typename T>
    T *ptr(int y = 0) const { return (T *)(ptr() + (size_t)y * step()); }
    int rows = 0, cols = 0;

  private:
    int channels_ = 1;
    Depth depth_ = F32;
    size_t step_ = 0;
    std::shared_ptr<void> mem_;
};

// Device matrix: replaces cv::cuda::GpuMat(rows, cols, type) for the planes of the hot path.
class GpuMat {
  public:
    GpuMat() = default;
    GpuMat(Stream &s, int rows, int cols, int channels, Depth depth = F32) : rows(rows), cols(cols), channels_(channels), depth_(depth) {
        smc_buffer *b = nullptr;
        check(smc_buffer_create(s.ctx(), rows, cols, channels, depth, &b));
        buf_ = std::shared_ptr<smc_buffer>(b, [](smc_buffer *q) { smc_buffer_destroy(q); });
    }
    void upload(const Mat &m, Stream &) { check(smc_buffer_upload(buf_.get(), m.ptr(), m.step())); }
    void download(Mat &m, Stream &) const { check(smc_buffer_download(buf_.get(), m.ptr(), m.step())); }
    void upload(const void *host, size_t hostStep) { check(smc_buffer_upload(buf_.get(), host, hostStep)); }
    void download(void *host, size_t hostStep) const { check(smc_buffer_download(buf_.get(), host, hostStep)); }
    void setToZero() { check(smc_buffer_fill_zero(buf_.get())); }
    int channels() const { return channels_; }
    Depth depth() const { return depth_; }
    bool empty() const { return !buf_; }
    smc_plane plane() const { return buf_ ? smc_buffer_plane(buf_.get()) : smc_plane{nullptr, 0}; }
    operator PtrStepSzb() const {  // "*ptrsCPUPtr = mats[i] is an implicit conversion of GPUMat to PtrStepSzb" (estimator.cpp:71)
        PtrStepSzb p;
        p.data = (unsigned char *)smc_buffer_dev(buf_.get());
        p.step = smc_buffer_step(buf_.get());
        p.cols = cols;
        p.rows = rows;
        return p;
    }
    int rows = 0, cols = 0;

  private:
    int channels_ = 1;
    Depth depth_ = F32;
    std::shared_ptr<smc_buffer> buf_;
};

// pbrt::Buffer (buffer.h:19-71) without the output (PFM / display) half, which stays in the reference.
class Buffer {
  public:
    Buffer() {}
    Buffer(Stream &s, const std::string &name, Mat mat) : Buffer(name, mat, GpuMat(s, mat.rows, mat.cols, mat.channels(), mat.depth())) {}
    Buffer(const std::string &name, Mat mat, GpuMat gpuMat) : name(name), mat(mat), matPtr(mat.ptr()), gpuMat(gpuMat) {
        const int nChannels = mat.channels();
        if (nChannels == 1) channelNames = {name};
        else if (nChannels == 3) channelNames = {name + ".R", name + ".G", name + ".B"};
        else
            for (int i = 1; i <= nChannels; i++) channelNames.push_back(name + "." + std::to_string(i));
    }
    inline void upload(Stream &stream) { gpuMat.upload(mat, stream); }
    inline void download(Stream &stream) { gpuMat.download(mat, stream); }

    std::string name;
    std::vector<std::string> channelNames;
    Mat mat;
    const unsigned char *matPtr = nullptr;
    GpuMat gpuMat;
};

// cv::cuda::stat_denoiser (cudaimgproc.hpp:736-799): the kernel-level API with the reference's own argument lists.  Every
// *Ptrs argument describes a DEVICE-resident 1 x N array of PtrStepSzb (the reference wraps it in a 1 x N GpuMat of
// CV_8UC(24)); film / filmFiltered / gBufferChannelCounts / gBufferDRFactors are direct device planes.
namespace stat_denoiser {
inline void setup() {}  // heap limit / cache config of the reference (stat_denoiser.cu:352-355) have no counterpart: nothing to set
inline void synchronize(Stream &stream) { stream.waitForCompletion(); }

template <typename T>
struct Channels;
template <>
struct Channels<float> {
    static constexpr int value = 1;
};
template <>
struct Channels<float3> {
    static constexpr int value = 3;
};

template <typename T>
inline void calculateMeanVars(const unsigned short ptrCount, const unsigned short width, const unsigned short height,
                              const PtrStepSzb &nPtrs, const PtrStepSzb &m2Ptrs, PtrStepSzb meanVarPtrs, Stream &stream) {
    check(smc_calculate_mean_vars_device_tables(stream.ctx(), Channels<T>::value, ptrCount, width, height, nPtrs.data,
                                                m2Ptrs.data, meanVarPtrs.data, stream.cudaStream()));
}

template <typename T>
inline void filter(const unsigned short ptrCount, const unsigned short width, const unsigned short height, const float dSFactor,
                   const unsigned char radius, const bool denoiseFilm, const PtrStepSzb &nPtrs, const PtrStepSzb &meanPtrs,
                   const PtrStepSzb &m2Ptrs, const PtrStepSzb &m3Ptrs, const PtrStepSzb &filmPtrs, const PtrStepSzb &film,
                   const PtrStepSzb &gBufferPtrs, const PtrStepSzb &gBufferChannelCounts, const PtrStepSzb &gBufferDRFactors,
                   const unsigned char nGBufs, PtrStepSzb meanCorrPtrs, PtrStepSzb discriminatorPtrs,
                   PtrStepSzb filmFilteredPtrs, PtrStepSzb filmFiltered, Stream &stream) {
    check(smc_filter_device_tables(stream.ctx(), Channels<T>::value, ptrCount, width, height, dSFactor, radius, denoiseFilm,
                                   nPtrs.data, meanPtrs.data, m2Ptrs.data, m3Ptrs.data, filmPtrs.data, film.data, film.step,
                                   gBufferPtrs.data, gBufferChannelCounts.data, gBufferDRFactors.data, nGBufs,
                                   meanCorrPtrs.data, discriminatorPtrs.data, filmFilteredPtrs.data, filmFiltered.data,
                                   filmFiltered.step, stream.cudaStream()));
}
// overload without film / filmFiltered (cudaimgproc.hpp:779-798)
template <typename T>
inline void filter(const unsigned short ptrCount, const unsigned short width, const unsigned short height, const float dSFactor,
                   const unsigned char radius, const PtrStepSzb &nPtrs, const PtrStepSzb &meanPtrs, const PtrStepSzb &m2Ptrs,
                   const PtrStepSzb &m3Ptrs, const PtrStepSzb &filmPtrs, const PtrStepSzb &gBufferPtrs,
                   const PtrStepSzb &gBufferChannelCounts, const PtrStepSzb &gBufferDRFactors, const unsigned char nGBufs,
                   PtrStepSzb meanCorrPtrs, PtrStepSzb discriminatorPtrs, PtrStepSzb filmFilteredPtrs, Stream &stream) {
    filter<T>(ptrCount, width, height, dSFactor, radius, false, nPtrs, meanPtrs, m2Ptrs, m3Ptrs, filmPtrs, PtrStepSzb(),
              gBufferPtrs, gBufferChannelCounts, gBufferDRFactors, nGBufs, meanCorrPtrs, discriminatorPtrs, filmFilteredPtrs,
              PtrStepSzb(), stream);
}
}  // namespace stat_denoiser

// ---- estimator.h:60-101 ------------------------------------------------------------------------------------------------
enum CUDAGroupIndex { DenoiseGroup = 0, CalculateMeanVarianceGroup = 1 };
static constexpr unsigned char nCUDAGroupIndices = 2;
enum StatTypeIndex {  // statpath.h:27-36
    Radiance = 0, MISBSDFWinRate = 1, MISLightWinRate = 2, StatMaterialID = 3, StatDepth = 4, StatNormal = 5, StatAlbedo = 6,
    ItRadiance = 7
};

struct StatTypeConfig {
    unsigned char type = 0;
    unsigned char index = 0;
    bool enable = false;
    unsigned char nBounces = 0;
    unsigned char bounceStart = 0;
    unsigned char bounceEnd = 0;
    unsigned char nChannels = 1;
    bool transform = false;
    unsigned char maxMoment = 1;
    bool gBuffer = false;
    bool enableForFilter = false;
    float filterSD = 0.f;
    std::vector<unsigned char> cudaGroups = {};
};

struct StatTypeConfigs {
    StatTypeConfig &operator[](size_t i) { return configs[i]; }
    const StatTypeConfig &operator[](size_t i) const { return configs[i]; }
    unsigned char nEnabled = 0;
    std::vector<StatTypeConfig> configs;
};

// pbrt::Estimator (estimator.h:241-379, estimator.cpp): owns every statistic plane (host Mat + device GpuMat), the
// G-buffer list and range factors, and runs Upload -> Denoise -> Download -> Synchronize on one stream.  Differences, all
// additions: AddSamples() accumulates batches of samples straight into the device planes (the reference accumulates in
// per-tile CPU state and MergeTile copies into the host planes); DenoiseHost() is Upload + Denoise + Download as one
// pipelined call; CalculateMeanVars() runs on the GPU per pixel (the reference's CPU loop reads n once per row).
class Estimator {
  public:
    Estimator(Stream &stream, const Buffer &filmBuffer, const StatTypeConfigs &statTypeConfigs, const float filterSD,
              const unsigned char filterRadius, const bool denoiseFilm, const bool acrrEnabled = false,
              const bool smisEnabled = false)
        : width(filmBuffer.mat.cols), height(filmBuffer.mat.rows), filterDSFactor(-.5f / (filterSD * filterSD)),
          filterRadius(filterRadius), denoiseFilm(denoiseFilm), acrrEnabled(acrrEnabled), smisEnabled(smisEnabled),
          stream(stream), filmBuffer(filmBuffer), filmFilteredBuffer(stream, "film-f", Mat(filmBuffer.mat.rows, filmBuffer.mat.cols, 3)) {
        floatBufferCounts = std::vector<unsigned char>(nCUDAGroupIndices, 0);
        rgbBufferCounts = std::vector<unsigned char>(nCUDAGroupIndices, 0);
        // this statTypeConfigs only holds configs of enabled buffers (estimator.h:268-269)
        std::copy_if(statTypeConfigs.configs.begin(), statTypeConfigs.configs.end(), std::back_inserter(this->statTypeConfigs.configs),
                     [](const StatTypeConfig &cfg) { return cfg.enable; });
        this->statTypeConfigs.nEnabled = (unsigned char)this->statTypeConfigs.configs.size();
        if (denoiseFilm) {
            uploadBuffers.insert(&this->filmBuffer);
            downloadBuffers.insert(&this->filmFilteredBuffer);
        }
        stat_denoiser::setup();
    }
    ~Estimator() {
        for (smc_denoiser *d : plans)
            if (d) smc_denoiser_destroy(d);
    }
    Estimator(const Estimator &) = delete;
    Estimator &operator=(const Estimator &) = delete;

    void RegisterGBuffer(Buffer &b, const float filterSD) {  // estimator.cpp:14-17
        gBuffers.push_back(b);
        gBufferDRFactors.emplace_back(-.5f / (filterSD * filterSD));
    }

    // estimator.cpp:86-289: per (type i, bounce j) the planes n, mean, m2, m3, mean-corr, discriminator, film-mean, film-m2,
    // film-mean-var, film-mean-f with the reference's names "t<i>-b<j>-<suffix>", aliasing (mean == film-mean and
    // m2 == film-m2 when !transform; film-mean-f of the radiance image shares film-f's host matrix), G-buffer registration and
    // upload / download sets.  The device pointer tables of the reference become denoise plans (smc_denoiser_create).
    void AllocateBuffers() {
        auto &cfgs = statTypeConfigs;
        for (auto *v : {&nBuffers, &meanBuffers, &m2Buffers, &m3Buffers, &meanCorrBuffers, &discriminatorBuffers, &filmBuffers,
                        &filmFilteredBuffers, &filmM2Buffers, &filmVarBuffers})
            v->resize(cfgs.nEnabled);
        for (unsigned char i = 0; i < cfgs.nEnabled; i++) {
            auto &cfg = cfgs.configs[i];
            cfg.index = i;
            for (auto *v : {&nBuffers, &meanBuffers, &m2Buffers, &m3Buffers, &meanCorrBuffers, &discriminatorBuffers, &filmBuffers,
                            &filmFilteredBuffers, &filmM2Buffers, &filmVarBuffers})
                (*v)[i].reserve(cfg.nBounces);  // pointers into these vectors are kept in the upload / download sets
            const int C = cfg.nChannels == 3 ? 3 : 1;
            for (unsigned char j = cfg.bounceStart; j < cfg.bounceEnd; j++) {
                const std::string pre = "t" + std::to_string(i) + "-b" + std::to_string(j);
                auto alloc = [&](std::vector<std::vector<Buffer>> &v, const char *suffix, Mat m) -> Buffer & {
                    return v[i].emplace_back(stream, pre + suffix, m);
                };
                alloc(nBuffers, "-n", Mat(height, width, 1, S32));
                if (cfg.transform) {
                    alloc(meanBuffers, "-mean", Mat(height, width, C));
                    alloc(m2Buffers, "-m2", Mat(height, width, C));
                    alloc(filmBuffers, "-film-mean", Mat(height, width, C));
                    alloc(filmM2Buffers, "-film-m2", Mat(height, width, C));
                } else {  // m2 and mean point to their film counterparts in case of no transformation (estimator.cpp:128-136)
                    Mat mean(height, width, C), m2(height, width, C);
                    GpuMat meanGPU(stream, height, width, C), m2GPU(stream, height, width, C);
                    meanBuffers[i].emplace_back(pre + "-mean", mean, meanGPU);
                    m2Buffers[i].emplace_back(pre + "-m2", m2, m2GPU);
                    filmBuffers[i].emplace_back(pre + "-film-mean", mean, meanGPU);
                    filmM2Buffers[i].emplace_back(pre + "-film-m2", m2, m2GPU);
                }
                alloc(m3Buffers, "-m3", Mat(height, width, C));
                alloc(meanCorrBuffers, "-mean-corr", Mat(height, width, C));
                alloc(discriminatorBuffers, "-discriminator", Mat(height, width, C));
                alloc(filmVarBuffers, "-film-mean-var", Mat(height, width, C));
                const bool isFilm = C == 3 && denoiseFilm && cfg.type == Radiance && j == 0;
                if (isFilm) filmFilteredBuffers[i].emplace_back(stream, pre + "-film-mean-f", filmFilteredBuffer.mat);
                else alloc(filmFilteredBuffers, "-film-mean-f", Mat(height, width, C));
                const size_t jj = filmBuffers[i].size() - 1;

                if (cfg.gBuffer && cfg.enableForFilter) {
                    RegisterGBuffer(filmBuffers[i][jj], cfg.filterSD);
                    uploadBuffers.insert(&filmBuffers[i][jj]);
                }
                for (unsigned char k : cfg.cudaGroups) {
                    auto &counts = C == 3 ? rgbBufferCounts : floatBufferCounts;
                    if (k != CalculateMeanVarianceGroup) {
                        counts[k]++;
                        runCUDA = true;
                    } else if (j == 0) {
                        counts[k]++;
                    }
                }
                const bool inDenoise = std::find(cfg.cudaGroups.begin(), cfg.cudaGroups.end(), DenoiseGroup) != cfg.cudaGroups.end();
                const bool inVar = std::find(cfg.cudaGroups.begin(), cfg.cudaGroups.end(), CalculateMeanVarianceGroup) != cfg.cudaGroups.end();
                if (inDenoise) {
                    uploadBuffers.insert(&nBuffers[i][jj]);
                    uploadBuffers.insert(&meanBuffers[i][jj]);
                    uploadBuffers.insert(&m2Buffers[i][jj]);
                    uploadBuffers.insert(&m3Buffers[i][jj]);
                    // RGB: skip the radiance image at bounce 0, covered by the film buffer; scalar: only with ACRR / SMIS
                    const bool moveValue = C == 3 ? !isFilm : (acrrEnabled || smisEnabled);
                    if (moveValue) {
                        if (cfg.transform) uploadBuffers.insert(&filmBuffers[i][jj]);
                        downloadBuffers.insert(&filmFilteredBuffers[i][jj]);
                    }
                    (C == 3 ? rgbImages : floatImages).push_back({i, jj});
                }
                if (inVar && j == 0) {
                    uploadBuffers.insert(&nBuffers[i][jj]);
                    uploadBuffers.insert(&filmM2Buffers[i][jj]);
                    downloadBuffers.insert(&filmVarBuffers[i][jj]);
                    varImages.push_back({i, jj});
                }
            }
        }
        BuildPlans();
    }

    // A batch of `nSamples` samples per pixel for statistic type `statTypeIndex` (index among ENABLED types), bounce j:
    // StatTile<T>::Add[Transform]SampleM<maxMoment> for every sample in order, then Merge[Transform]Tile -- the planes on
    // the device are the running state.  samples: [nSamples][height][width][nChannels] float32, tightly packed, in DEVICE
    // memory when onDevice, else in host memory (staged through a device buffer owned by the estimator).
    void AddSamples(const unsigned char statTypeIndex, const unsigned char bounceIndex, const float *samples, const int nSamples,
                    const bool onDevice = false) {
        const auto &cfg = statTypeConfigs[statTypeIndex];
        smc_moments m;
        m.width = width; m.height = height; m.channels = cfg.nChannels == 3 ? 3 : 1;
        m.n = nBuffers[statTypeIndex][bounceIndex].gpuMat.plane();
        m.mean = meanBuffers[statTypeIndex][bounceIndex].gpuMat.plane();
        m.m2 = m2Buffers[statTypeIndex][bounceIndex].gpuMat.plane();
        m.m3 = m3Buffers[statTypeIndex][bounceIndex].gpuMat.plane();
        m.film_mean = filmBuffers[statTypeIndex][bounceIndex].gpuMat.plane();
        m.film_m2 = filmM2Buffers[statTypeIndex][bounceIndex].gpuMat.plane();
        const float *dev = samples;
        if (!onDevice) {
            const size_t floats = (size_t)nSamples * height * width * m.channels;
            if (floats > 0x7fffffffu) throw Exception(SMC_ERR_UNSUPPORTED, "AddSamples: batch too large for one staging plane");
            if (staging.empty() || (size_t)staging.cols < floats) staging = GpuMat(stream, 1, (int)floats, 1);  // 1 row: contiguous
            check(smc_memcpy_device(stream.ctx(), staging.plane().dev, samples, floats * 4));
            dev = (const float *)staging.plane().dev;
        }
        check(smc_accumulate(stream.ctx(), &m, dev, nSamples, cfg.transform, cfg.maxMoment, 0, 0));
        deviceResident.insert(&nBuffers[statTypeIndex][bounceIndex]);
        deviceResident.insert(&meanBuffers[statTypeIndex][bounceIndex]);
        deviceResident.insert(&m2Buffers[statTypeIndex][bounceIndex]);
        deviceResident.insert(&m3Buffers[statTypeIndex][bounceIndex]);
        deviceResident.insert(&filmBuffers[statTypeIndex][bounceIndex]);
        deviceResident.insert(&filmM2Buffers[statTypeIndex][bounceIndex]);
    }

    // estimator.cpp:409-416; planes that AddSamples keeps on the device are already there and are skipped
    void Upload() {
        for (Buffer *b : uploadBuffers)
            if (!deviceResident.count(b)) b->upload(stream);
    }
    void Download() {  // estimator.cpp:418-425
        for (Buffer *b : downloadBuffers) b->download(stream);
    }
    // estimator.cpp:427-489: the float group first, then the RGB group, both with the same film / film-f
    void Denoise() {
        for (smc_denoiser *d : plans)
            if (d) check(smc_denoiser_run(d));
    }
    // Upload + Denoise + Download of the RGB group as one pipelined call (smc_denoiser_run_host); the scalar group, if
    // any, runs through the plain sequence first.
    void DenoiseHost() {
        if (plans[0]) {
            Upload();
            check(smc_denoiser_run(plans[0]));
        }
        if (!plans[1]) return;
        std::vector<smc_plane> hn, hmean, hm2, hm3, hfilm, hout, hg;
        auto host = [&](Buffer &b, bool wanted) {
            return wanted ? smc_plane{(void *)b.mat.ptr(), b.mat.step()} : smc_plane{nullptr, 0};
        };
        for (auto &ij : rgbImages) {
            const size_t i = ij.first, j = ij.second;
            hn.push_back(host(nBuffers[i][j], uploadBuffers.count(&nBuffers[i][j]) && !deviceResident.count(&nBuffers[i][j])));
            hmean.push_back(host(meanBuffers[i][j], uploadBuffers.count(&meanBuffers[i][j]) && !deviceResident.count(&meanBuffers[i][j])));
            hm2.push_back(host(m2Buffers[i][j], uploadBuffers.count(&m2Buffers[i][j]) && !deviceResident.count(&m2Buffers[i][j])));
            hm3.push_back(host(m3Buffers[i][j], uploadBuffers.count(&m3Buffers[i][j]) && !deviceResident.count(&m3Buffers[i][j])));
            hfilm.push_back(host(filmBuffers[i][j], uploadBuffers.count(&filmBuffers[i][j]) && !deviceResident.count(&filmBuffers[i][j])));
            hout.push_back(host(filmFilteredBuffers[i][j], downloadBuffers.count(&filmFilteredBuffers[i][j]) > 0));
        }
        for (Buffer &g : gBuffers) {
            bool resident = false;
            for (Buffer *r : deviceResident) resident |= (r->matPtr == g.matPtr);
            hg.push_back(host(g, !resident));
        }
        smc_host_io io{};
        io.n = hn.data(); io.mean = hmean.data(); io.m2 = hm2.data(); io.m3 = hm3.data(); io.film_ptrs = hfilm.data();
        io.film = denoiseFilm ? host(filmBuffer, true) : smc_plane{nullptr, 0};
        io.gbufs = hg.data();
        io.film_filtered_ptrs = hout.data();
        io.film_filtered = denoiseFilm ? host(filmFilteredBuffer, true) : smc_plane{nullptr, 0};
        check(smc_denoiser_run_host(plans[1], &io, 0));
    }
    // estimator.cpp:491-569, on the GPU and per pixel (stat_denoiser.cu:148-159)
    void CalculateMeanVars() {
        for (auto &ij : varImages) {
            const size_t i = ij.first, j = ij.second;
            check(smc_calculate_mean_vars(stream.ctx(), width, height, statTypeConfigs[i].nChannels == 3 ? 3 : 1,
                                          nBuffers[i][j].gpuMat.plane(), filmM2Buffers[i][j].gpuMat.plane(),
                                          filmVarBuffers[i][j].gpuMat.plane()));
        }
    }
    void Synchronize() { stat_denoiser::synchronize(stream); }  // estimator.cpp:571-573

    const unsigned short width;
    const unsigned short height;
    const float filterDSFactor;
    const unsigned char filterRadius;
    const bool denoiseFilm;
    const bool acrrEnabled;
    const bool smisEnabled;

    std::vector<unsigned char> floatBufferCounts;
    std::vector<unsigned char> rgbBufferCounts;
    bool runCUDA = false;

    Stream &stream;
    Buffer filmBuffer;
    Buffer filmFilteredBuffer;
    StatTypeConfigs statTypeConfigs;
    std::unordered_set<Buffer *> uploadBuffers;
    std::unordered_set<Buffer *> downloadBuffers;
    std::unordered_set<Buffer *> deviceResident;  // planes whose current contents live on the device (AddSamples)

    std::vector<std::vector<Buffer>> nBuffers, meanBuffers, m2Buffers, m3Buffers;
    std::vector<std::vector<Buffer>> filmBuffers, filmM2Buffers, filmFilteredBuffers, filmVarBuffers;
    std::vector<std::vector<Buffer>> meanCorrBuffers, discriminatorBuffers;
    std::vector<Buffer> gBuffers;
    std::vector<float> gBufferDRFactors;

  private:
    // one plan per CUDA denoise group: [0] float images, [1] RGB images (the reference's two filter<T> calls)
    void BuildPlans() {
        plans[0] = BuildPlan(floatImages, 1);
        plans[1] = BuildPlan(rgbImages, 3);
    }
    smc_denoiser *BuildPlan(const std::vector<std::pair<size_t, size_t>> &images, int C) {
        if (images.empty()) return nullptr;
        std::vector<smc_plane> n, mean, m2, m3, film, mc, disc, out, g;
        for (auto &ij : images) {
            const size_t i = ij.first, j = ij.second;
            n.push_back(nBuffers[i][j].gpuMat.plane());
            mean.push_back(meanBuffers[i][j].gpuMat.plane());
            m2.push_back(m2Buffers[i][j].gpuMat.plane());
            m3.push_back(m3Buffers[i][j].gpuMat.plane());
            film.push_back(filmBuffers[i][j].gpuMat.plane());
            mc.push_back(meanCorrBuffers[i][j].gpuMat.plane());
            disc.push_back(discriminatorBuffers[i][j].gpuMat.plane());
            out.push_back(filmFilteredBuffers[i][j].gpuMat.plane());
        }
        std::vector<uint8_t> gch;
        for (Buffer &b : gBuffers) {
            g.push_back(b.gpuMat.plane());
            gch.push_back((uint8_t)b.gpuMat.channels());
        }
        smc_filter_desc d{};
        d.channels = C; d.ptr_count = (int)images.size(); d.width = width; d.height = height;
        d.ds_factor = filterDSFactor; d.radius = filterRadius; d.denoise_film = denoiseFilm; d.membership = SMC_MEMBER_WELCH;
        d.n = n.data(); d.mean = mean.data(); d.m2 = m2.data(); d.m3 = m3.data(); d.film_ptrs = film.data();
        d.film = filmBuffer.gpuMat.plane();
        d.n_gbufs = (int)g.size(); d.gbufs = g.data(); d.gbuf_channels = gch.data(); d.gbuf_dr_factors = gBufferDRFactors.data();
        d.mean_corr = mc.data(); d.disc = disc.data(); d.film_filtered_ptrs = out.data();
        d.film_filtered = filmFilteredBuffer.gpuMat.plane();
        smc_denoiser *plan = nullptr;
        check(smc_denoiser_create(stream.ctx(), &d, &plan));
        return plan;
    }

    std::vector<std::pair<size_t, size_t>> floatImages, rgbImages, varImages;  // (type, bounce slot) per CUDA group
    smc_denoiser *plans[2] = {nullptr, nullptr};
    GpuMat staging;
};

}  // namespace statmc
#endif  // STATMC_B200_HPP
