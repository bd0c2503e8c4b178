/*
 * ref_null_device.cpp -- CPU stand-in for the DEVICE half of the OpenCV surface src/statistics/ uses, so that the
 * reference's renderer (compiled unmodified, oracle/Makefile target `pbrt`) runs in a container WITHOUT a GPU and
 * produces REAL statistic dumps (BASELINE.json configs[0]: veach-mis, StatPathIntegrator, 16 spp).
 *
 * TEST / FIXTURE INFRASTRUCTURE ONLY -- this is the oracle's side of the fence: "device" memory is host memory,
 * uploads are memcpy, and cv::cuda::stat_denoiser::filter<T> runs the CPU restatement (statmc_oracle.c: smo_prepass +
 * smo_filter_f32) with the reference's dispatch (stat_denoiser.cu:397-475) and routing (:251-273, :319-344).  The
 * product never links this file; the host half (cv::Mat & co.) comes from integration/opencv_link_shim.cpp compiled
 * with -DSMC_SHIM_HOST_ONLY.  The Student-t table is read from the float32 file named by $STATMC_T_LUT (1024 entries,
 * written from tests/golden/t_quantiles.json by tools/make_golden_render.py).
 */
#include <opencv2/core.hpp>
#include <opencv2/core/cuda.hpp>
#include <opencv2/cudaimgproc.hpp>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

struct float3 {
    float x, y, z;
};

// ---- oracle/statmc_oracle.c ------------------------------------------------------------------------------------------
extern "C" {
typedef struct smo_plane {
    void *data;
    size_t step;
} smo_plane;
typedef struct smo_filter_args {
    int W, H;
    int C;
    int value_channels;
    int radius;
    float ds_factor;
    int n_gbufs;
    const smo_plane *gbufs;
    const uint8_t *gbuf_channels;
    const float *gbuf_dr_factors;
    const smo_plane *mean_corr, *disc;
    const smo_plane *n, *mean, *m2;
    const float *lut;
    const smo_plane *value;
    const smo_plane *out;
    const smo_plane *accepted;
    int mode;
} smo_filter_args;
void smo_prepass(int W, int H, int C, const float *lut, const smo_plane *n, const smo_plane *mean, const smo_plane *m2,
                 const smo_plane *m3, const smo_plane *mean_corr, const smo_plane *disc);
void smo_filter_f32(const smo_filter_args *a);
}

namespace {

[[noreturn]] void fail(const std::string &what) { throw std::runtime_error("null device: " + what); }

const float *t_lut() {
    static std::vector<float> lut;
    if (lut.empty()) {
        const char *path = std::getenv("STATMC_T_LUT");
        if (!path) fail("set STATMC_T_LUT to the 1024-entry float32 Student-t table (tools/make_golden_render.py)");
        FILE *f = std::fopen(path, "rb");
        if (!f) fail(std::string("cannot open ") + path);
        lut.resize(1024);
        const size_t got = std::fread(lut.data(), 4, 1024, f);
        std::fclose(f);
        if (got != 1024) fail("short Student-t table");
    }
    return lut.data();
}

class HostAllocator : public cv::cuda::GpuMat::Allocator {
public:
    bool allocate(cv::cuda::GpuMat *mat, int rows, int cols, size_t elemSize) override {
        const size_t row_bytes = (size_t)cols * elemSize;
        // pitched like cudaMallocPitch when 2-D (gpu_mat.cu:112-123), so that every consumer sees step != cols * elemSize
        const size_t step = (rows > 1 && cols > 1) ? ((row_bytes + 255) / 256) * 256 : row_bytes;
        void *p = nullptr;
        if (posix_memalign(&p, 256, step * (size_t)rows + 256)) return false;
        std::memset(p, 0, step * (size_t)rows);
        mat->data = static_cast<uchar *>(p);
        mat->step = step;
        mat->refcount = static_cast<int *>(std::malloc(sizeof(int)));
        return true;
    }
    void free(cv::cuda::GpuMat *mat) override {
        std::free(mat->datastart);
        std::free(mat->refcount);
    }
};

smo_plane plane(const cv::cuda::PtrStepSzb &p) { return smo_plane{p.data, p.step}; }

template <typename T>
struct ChannelsOf;
template <>
struct ChannelsOf<float> {
    static constexpr int value = 1;
};
template <>
struct ChannelsOf< ::float3> {
    static constexpr int value = 3;
};

}  // namespace

namespace cv {
namespace cuda {

GpuMat::Allocator *GpuMat::defaultAllocator() {
    static HostAllocator *a = new HostAllocator;
    return a;
}

void GpuMat::create(int _rows, int _cols, int _type) {
    _type &= Mat::TYPE_MASK;
    if (rows == _rows && cols == _cols && type() == _type && data) return;
    if (data) release();
    if (_rows > 0 && _cols > 0) {
        flags = Mat::MAGIC_VAL + _type;
        rows = _rows;
        cols = _cols;
        const size_t esz = elemSize();
        if (!allocator) allocator = defaultAllocator();
        if (!allocator->allocate(this, rows, cols, esz)) fail("allocation failed");
        if (esz * cols == step) flags |= Mat::CONTINUOUS_FLAG;
        datastart = data;
        dataend = data + step * (rows - 1) + cols * esz;
        if (refcount) *refcount = 1;
    }
}

void GpuMat::release() {
    if (refcount && CV_XADD(refcount, -1) == 1) allocator->free(this);
    dataend = data = datastart = 0;
    step = rows = cols = 0;
    refcount = 0;
}

void GpuMat::upload(InputArray arr, Stream &) {
    const Mat &m = *static_cast<const Mat *>(arr.getObj());
    if (m.rows <= 0 || m.cols <= 0 || !m.data) return;
    create(m.rows, m.cols, m.type());
    const size_t row_bytes = (size_t)cols * elemSize();
    for (int y = 0; y < rows; y++) std::memcpy(data + step * y, m.data + m.step.p[0] * y, row_bytes);
}

void GpuMat::download(OutputArray _dst, Stream &) const {
    Mat &dst = *static_cast<Mat *>(_dst.getObj());
    if (dst.rows != rows || dst.cols != cols || dst.type() != type() || !dst.data) dst = Mat(rows, cols, type());
    const size_t row_bytes = (size_t)cols * elemSize();
    for (int y = 0; y < rows; y++) std::memcpy(dst.data + dst.step.p[0] * y, data + step * y, row_bytes);
}

class Stream::Impl {};
Stream::Stream() : impl_(new Impl) {}
Stream &Stream::Null() {
    static Stream *s = new Stream;
    return *s;
}

namespace stat_denoiser {

void setup() {}
void synchronize(Stream &) {}

// stat_denoiser.cu:397-475 on the CPU: per image z the two prepass kernels, then the filter with the reference's routing
template <typename T>
void filter(const unsigned short ptrCount, const unsigned short width, const unsigned short height, const float dSFactor,
            const unsigned char radius, const bool denoiseFilm, const PtrStepSzb &nPtrs, const PtrStepSzb &meanPtrs,
            const PtrStepSzb &m2Ptrs, const PtrStepSzb &m3Ptrs, const PtrStepSzb &filmPtrs, const PtrStepSzb &film,
            const PtrStepSzb &gBufferPtrs, const PtrStepSzb &gBufferChannelCounts, const PtrStepSzb &gBufferDRFactors,
            const unsigned char nGBufs, PtrStepSzb meanCorrPtrs, PtrStepSzb discriminatorPtrs, PtrStepSzb filmFilteredPtrs,
            PtrStepSzb filmFiltered, Stream &) {
    constexpr int C = ChannelsOf<T>::value;
    const auto tab = [](const PtrStepSzb &t) { return reinterpret_cast<const PtrStepSzb *>(t.data); };
    std::vector<smo_plane> gb(nGBufs);
    for (int g = 0; g < nGBufs; g++) gb[g] = plane(tab(gBufferPtrs)[g]);
    const smo_plane filmP = plane(film), filmFP = plane(filmFiltered);
    for (int z = 0; z < ptrCount; z++) {
        const smo_plane n = plane(tab(nPtrs)[z]), mean = plane(tab(meanPtrs)[z]), m2 = plane(tab(m2Ptrs)[z]),
                        m3 = plane(tab(m3Ptrs)[z]), mc = plane(tab(meanCorrPtrs)[z]), dc = plane(tab(discriminatorPtrs)[z]),
                        val = plane(tab(filmPtrs)[z]), out = plane(tab(filmFilteredPtrs)[z]);
        smo_prepass(width, height, C, t_lut(), &n, &mean, &m2, &m3, &mc, &dc);
        smo_filter_args a;
        std::memset(&a, 0, sizeof(a));
        a.W = width;
        a.H = height;
        a.C = C;
        a.radius = radius;
        a.ds_factor = dSFactor;
        a.n_gbufs = nGBufs;
        a.gbufs = gb.data();
        a.gbuf_channels = gBufferChannelCounts.data;
        a.gbuf_dr_factors = reinterpret_cast<const float *>(gBufferDRFactors.data);
        a.mean_corr = &mc;
        a.disc = &dc;
        a.mode = 0;
        const bool filmImage = denoiseFilm && z == 0;
        if (C == 3) {  // :319-344: image 0 filters the film instead of its own film-mean
            a.value_channels = 3;
            a.value = filmImage ? &filmP : &val;
            a.out = filmImage ? &filmFP : &out;
            smo_filter_f32(&a);
        } else {  // :251-273: the scalar kernel filters its film-mean and, for image 0, the RGB film as well
            a.value_channels = 1;
            a.value = &val;
            a.out = &out;
            smo_filter_f32(&a);
            if (filmImage) {
                a.value_channels = 3;
                a.value = &filmP;
                a.out = &filmFP;
                smo_filter_f32(&a);
            }
        }
    }
}

template void filter<float>(const unsigned short, const unsigned short, const unsigned short, const float, const unsigned char,
                            const bool, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &,
                            const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &,
                            const unsigned char, PtrStepSzb, PtrStepSzb, PtrStepSzb, PtrStepSzb, Stream &);
template void filter< ::float3>(const unsigned short, const unsigned short, const unsigned short, const float,
                               const unsigned char, const bool, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &,
                               const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &,
                               const PtrStepSzb &, const PtrStepSzb &, const unsigned char, PtrStepSzb, PtrStepSzb, PtrStepSzb,
                               PtrStepSzb, Stream &);

}  // namespace stat_denoiser
}  // namespace cuda
}  // namespace cv
