// opencv_link_shim.cpp -- link-level drop-in: the handful of OpenCV symbols StatMC's src/statistics/ actually uses,
// implemented on top of libstatmc_b200.so's C ABI (include/statmc_b200.h).
//
// With this file, the reference's src/statistics/estimator.cpp and buffer.cpp compile UNMODIFIED against the OpenCV
// *headers* of the StatMC checkout and link against libstatmc_b200.so instead of a CUDA-enabled OpenCV build:
// `nm -u` of those two objects lists exactly the cv:: symbols defined below.  What each
// one replaces (paths relative to the StatMC checkout; OCV = src/ext/opencv/modules, CIP = src/ext/opencv_contrib/
// modules/cudaimgproc):
//   cv::Mat ctor/dtor/copy/move/convertTo        OCV/core/src/matrix.cpp, convert.dispatch.cpp   (host planes; >= 64 KiB pinned)
//   cv::cuda::GpuMat create/release/upload/download, defaultAllocator   OCV/core/src/cuda/gpu_mat.cu:112-234
//                                                 -> smc_buffer_create / smc_buffer_destroy / smc_buffer_upload / _download
//   cv::cuda::Stream                              OCV/core/src/cuda_stream.cpp -> the smc_context's stream
//   cv::cuda::stat_denoiser::{setup, synchronize, calculateMeanVars<T>, filter<T>}   CIP/src/stat_denoiser.cpp:90-137,
//                                                 CIP/src/cuda/stat_denoiser.cu:352-483 -> smc_filter_device_tables & co.
//   cv::cvtColor (RGB<->BGR only), cv::merge, cv::imwrite / cv::imread (.pfm only), cv::glob   what buffer.cpp:40-53 and
//                                                 statpath.cpp:449-453, 479-481 need to dump / reload statistic planes
// Everything else of OpenCV that those headers declare is deliberately absent: an accidental new dependency shows up as a
// link error, not as silently different behaviour.
//
// This file is ours (no OpenCV or StatMC code in it); it needs the OpenCV headers only for the class layouts, so it is
// compiled where a StatMC checkout is present (oracle/Makefile builds it for the parity harness).  One process-wide
// smc_context on device $STATMC_B200_DEVICE (default 0) backs every GpuMat and Stream, like OpenCV's current device.
// Errors throw std::runtime_error with smc_last_error() (the reference throws cv::Exception and dies, common.hpp:66-76).
// -DSMC_SHIM_HOST_ONLY compiles the host half only (cv::Mat, convertTo, cvtColor, merge, imwrite/imread, glob): the parity
// suite pairs it with a CPU stand-in for the device half (oracle/ref_null_device.cpp) to render fixtures without a GPU.
#include <opencv2/core.hpp>
#include <opencv2/core/cuda.hpp>
#include <opencv2/cudaimgproc.hpp>
#include <opencv2/imgcodecs.hpp>
#include <opencv2/imgproc.hpp>

#include <glob.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "statmc_b200.h"

// the element type the reference's host code names the RGB instantiation with (estimator.cpp:8-10,
// samples/stat_denoiser/main.cpp:14-16); same mangled name as CUDA's ::float3
#ifndef __VECTOR_TYPES_H__
struct float3 {
    float x, y, z;
};
#endif

namespace {

[[noreturn]] void fail(const std::string &what) {
    const char *e = smc_last_error();
    throw std::runtime_error("statmc_b200 OpenCV link shim: " + what + (e && *e ? std::string(": ") + e : std::string()));
}
void check(int rc, const char *what) {
    if (rc != SMC_OK) fail(what);
}

#ifndef SMC_SHIM_HOST_ONLY
// A GpuMat::upload of a page-locked plane that has not been issued yet.  Estimator::Upload (EST.cpp:409-432) uploads every
// plane and Estimator::Denoise runs the filter right after it (statpath.cpp:406-418); holding the copies back until the filter
// call lets the library move them inside its row-chunked pipeline, PCIe overlapping the kernels (smc_filter_device_tables_host).
// Anything else that could observe the device plane or release the host memory issues the held copies first.
// STATMC_B200_DEFER_UPLOADS=0 copies at once, as GpuMat::upload does.
struct Pending {
    smc_buffer *buf;
    void *dev;
    size_t dev_step;
    const void *host;
    size_t host_step, row_bytes;
    int rows;
    const void *host_owner;  // the cv::UMatData the rows belong to
};
struct Shim {
    smc_context *ctx = nullptr;
    std::mutex mu;
    std::unordered_map<const void *, smc_buffer *> owner;  // GpuMat::datastart -> the smc_buffer that owns the memory
    std::vector<Pending> pending;
    bool defer = true;
    Shim() {
        const char *d = std::getenv("STATMC_B200_DEVICE");
        check(smc_context_create(d ? std::atoi(d) : 0, &ctx), "smc_context_create");
        const char *f = std::getenv("STATMC_B200_DEFER_UPLOADS");
        defer = !(f && std::atoi(f) == 0);
    }
};
Shim &shim() {
    static Shim *s = new Shim;  // never destroyed: GpuMats with static storage may outlive any destructor order
    return *s;
}
smc_buffer *owner_of(const void *dev) {
    Shim &s = shim();
    std::lock_guard<std::mutex> g(s.mu);
    auto it = s.owner.find(dev);
    return it == s.owner.end() ? nullptr : it->second;
}
std::vector<Pending> pending_take() {
    Shim &s = shim();
    std::lock_guard<std::mutex> g(s.mu);
    std::vector<Pending> v;
    v.swap(s.pending);
    return v;
}
void pending_flush() {  // issue every held copy on the context stream
    for (const Pending &p : pending_take()) check(smc_buffer_upload(p.buf, p.host, p.host_step), "GpuMat::upload (deferred)");
}
void pending_drop(const void *dev) {  // the device plane goes away or is about to be overwritten by a newer upload
    Shim &s = shim();
    std::lock_guard<std::mutex> g(s.mu);
    for (size_t i = 0; i < s.pending.size();)
        if (s.pending[i].dev == dev) s.pending.erase(s.pending.begin() + i);
        else i++;
}
bool pending_reads(const void *host_owner) {
    Shim &s = shim();
    std::lock_guard<std::mutex> g(s.mu);
    for (const Pending &p : s.pending)
        if (p.host_owner == host_owner) return true;
    return false;
}

#endif  // SMC_SHIM_HOST_ONLY

// ---- host memory of cv::Mat ------------------------------------------------------------------------------------------
constexpr size_t kPinnedFrom = 64 << 10;  // planes are pinned so that Buffer::upload/download overlap; tiny tables are not

cv::UMatData *host_new(size_t bytes) {
    cv::UMatData *u = static_cast<cv::UMatData *>(std::calloc(1, sizeof(cv::UMatData)));
    if (!u) throw std::bad_alloc();
    void *p = nullptr;
    bool pinned = false;
    if (bytes >= kPinnedFrom && smc_host_alloc(bytes, &p) == SMC_OK && p) {
        pinned = true;
        std::memset(p, 0, bytes);
    } else {
        p = nullptr;
        if (posix_memalign(&p, 64, bytes ? bytes : 64)) {
            std::free(u);
            throw std::bad_alloc();
        }
    }
    u->data = u->origdata = static_cast<uchar *>(p);
    u->size = bytes;
    u->refcount = 1;
    u->userdata = pinned ? u : nullptr;
    return u;
}
void host_delete(cv::UMatData *u) {
#ifndef SMC_SHIM_HOST_ONLY
    if (u->userdata && pending_reads(u)) {  // a held upload still reads these rows
        pending_flush();
        smc_synchronize(shim().ctx);
    }
#endif
    if (u->userdata)
        smc_host_free(u->origdata);
    else
        std::free(u->origdata);
    std::free(u);
}

void mat_reset(cv::Mat &m) {  // fields of an empty header (what Mat() leaves)
    m.flags = cv::Mat::MAGIC_VAL;
    m.dims = m.rows = m.cols = 0;
    m.data = nullptr;
    m.datastart = m.dataend = m.datalimit = nullptr;
    m.allocator = nullptr;
    m.u = nullptr;
    m.step.buf[0] = m.step.buf[1] = 0;
}
void mat_take_header(cv::Mat &d, const cv::Mat &s) {  // header copy without touching reference counts
    if (s.dims > 2) fail("cv::Mat with more than 2 dimensions");
    d.flags = s.flags;
    d.dims = s.dims;
    d.rows = s.rows;
    d.cols = s.cols;
    d.data = s.data;
    d.datastart = s.datastart;
    d.dataend = s.dataend;
    d.datalimit = s.datalimit;
    d.allocator = s.allocator;
    d.u = s.u;
    d.step.buf[0] = s.step.p[0];
    d.step.buf[1] = s.dims >= 2 ? s.step.p[1] : 0;
}
void mat_create2d(cv::Mat &m, int rows, int cols, int type) {  // `m` must hold no data
    if (rows < 0 || cols < 0) fail("negative cv::Mat size");
    type &= cv::Mat::TYPE_MASK;
    const size_t esz = CV_ELEM_SIZE(type);
    m.flags = cv::Mat::MAGIC_VAL | type | cv::Mat::CONTINUOUS_FLAG;
    m.dims = 2;
    m.rows = rows;
    m.cols = cols;
    m.step.buf[0] = esz * (size_t)cols;
    m.step.buf[1] = esz;
    const size_t bytes = esz * (size_t)rows * (size_t)cols;
    if (bytes) {
        m.u = host_new(bytes);
        m.data = m.u->data;
        m.datastart = m.data;
        m.dataend = m.datalimit = m.data + bytes;
    }
}
// destination of an OutputArray that wraps a cv::Mat, (re)allocated like _OutputArray::create
cv::Mat &out_mat(const cv::_OutputArray &a, int rows, int cols, int type) {
    if ((a.getFlags() & cv::_InputArray::KIND_MASK) != cv::_InputArray::MAT) fail("only cv::Mat output arrays are supported");
    cv::Mat &m = *static_cast<cv::Mat *>(a.getObj());
    if (m.dims != 2 || m.rows != rows || m.cols != cols || m.type() != (int)(type & cv::Mat::TYPE_MASK) ||
        (!m.data && rows > 0 && cols > 0))
        m = cv::Mat(rows, cols, type);
    return m;
}
const cv::Mat &in_mat(const cv::_InputArray &a) {
    if ((a.getFlags() & cv::_InputArray::KIND_MASK) != cv::_InputArray::MAT) fail("only cv::Mat input arrays are supported");
    return *static_cast<const cv::Mat *>(a.getObj());
}

#ifndef SMC_SHIM_HOST_ONLY
// ---- device memory of cv::cuda::GpuMat ---------------------------------------------------------------------------------
class ShimAllocator : public cv::cuda::GpuMat::Allocator {
public:
    bool allocate(cv::cuda::GpuMat *mat, int rows, int cols, size_t elemSize) override {
        // smc_buffer_create speaks in 4-byte elements; descriptor tables (24 B) and planes (4 / 12 B) map directly, byte
        // arrays (the G-buffer channel counts, CV_8UC1) are rounded up to whole words.  Pitch rule = gpu_mat.cu:112-123.
        smc_buffer *b = nullptr;
        const size_t row_bytes = (size_t)cols * elemSize;
        int rc;
        if (elemSize % 4 == 0 && elemSize / 4 <= 512)
            rc = smc_buffer_create(shim().ctx, rows, cols, (int)(elemSize / 4), SMC_F32, &b);
        else
            rc = smc_buffer_create(shim().ctx, rows, (int)((row_bytes + 3) / 4), 1, SMC_F32, &b);
        if (rc != SMC_OK) return false;
        mat->data = static_cast<uchar *>(smc_buffer_dev(b));
        mat->step = (rows > 1 && cols > 1) ? smc_buffer_step(b) : row_bytes;
        mat->refcount = static_cast<int *>(std::malloc(sizeof(int)));
        Shim &s = shim();
        std::lock_guard<std::mutex> g(s.mu);
        s.owner[mat->data] = b;
        return true;
    }
    void free(cv::cuda::GpuMat *mat) override {
        smc_buffer *b = nullptr;
        pending_drop(mat->datastart);
        {
            Shim &s = shim();
            std::lock_guard<std::mutex> g(s.mu);
            auto it = s.owner.find(mat->datastart);
            if (it != s.owner.end()) {
                b = it->second;
                s.owner.erase(it);
            }
        }
        // kernels of the context stream may still read the plane: GpuMat::release is synchronous in OpenCV too (cudaFree)
        if (b) {
            smc_synchronize(shim().ctx);
            smc_buffer_destroy(b);
        }
        std::free(mat->refcount);
    }
};

struct Plane2D {
    smc_buffer *buf;
    size_t row_bytes;
};
Plane2D plane_of(const cv::cuda::GpuMat &g) {
    if (g.data != g.datastart) fail("GpuMat sub-views (ROI) are not supported by the shim");
    smc_buffer *b = owner_of(g.datastart);
    if (!b) fail("GpuMat memory was not allocated through the shim");
    return {b, (size_t)g.cols * g.elemSize()};
}

#endif  // SMC_SHIM_HOST_ONLY

}  // namespace

// ========================================================================================================================
namespace cv {

Mat::Mat() CV_NOEXCEPT : flags(MAGIC_VAL), dims(0), rows(0), cols(0), data(0), datastart(0), dataend(0), datalimit(0),
                         allocator(0), u(0), size(&rows), step(0) {}

Mat::Mat(int _rows, int _cols, int _type) : flags(MAGIC_VAL), dims(0), rows(0), cols(0), data(0), datastart(0), dataend(0),
                                            datalimit(0), allocator(0), u(0), size(&rows), step(0) {
    mat_create2d(*this, _rows, _cols, _type);
}

Mat::Mat(const Mat &m) : flags(MAGIC_VAL), dims(0), rows(0), cols(0), data(0), datastart(0), dataend(0), datalimit(0),
                         allocator(0), u(0), size(&rows), step(0) {
    mat_take_header(*this, m);
    if (u) CV_XADD(&u->refcount, 1);
}

Mat::Mat(Mat &&m) : flags(MAGIC_VAL), dims(0), rows(0), cols(0), data(0), datastart(0), dataend(0), datalimit(0),
                    allocator(0), u(0), size(&rows), step(0) {
    mat_take_header(*this, m);
    mat_reset(m);
}

Mat::~Mat() { release(); }

Mat &Mat::operator=(const Mat &m) {
    if (this != &m) {
        if (m.u) CV_XADD(&m.u->refcount, 1);
        release();
        mat_take_header(*this, m);
    }
    return *this;
}

Mat &Mat::operator=(Mat &&m) {
    if (this != &m) {
        release();
        mat_take_header(*this, m);
        mat_reset(m);
    }
    return *this;
}

void Mat::release() {  // matrix.cpp Mat::release
    if (u && CV_XADD(&u->refcount, -1) == 1) deallocate();
    u = nullptr;
    datastart = dataend = datalimit = data = nullptr;
    for (int i = 0; i < dims; i++) size.p[i] = 0;
}

void Mat::deallocate() {  // the last reference went
    if (u) {
        UMatData *u_ = u;
        u = nullptr;
        host_delete(u_);
    }
}

// convert.dispatch.cpp Mat::convertTo, for the depths the statistics planes use (CV_32S `n`, CV_32F everything else):
// dst = saturate_cast<D>(src * alpha + beta); float -> int rounds half to even (cvRound).
void Mat::convertTo(OutputArray _dst, int rtype, double alpha, double beta) const {
    const int sdepth = depth(), cn = channels();
    const int ddepth = rtype < 0 ? sdepth : CV_MAT_DEPTH(rtype);
    if ((sdepth != CV_32F && sdepth != CV_32S) || (ddepth != CV_32F && ddepth != CV_32S))
        fail("Mat::convertTo: only CV_32F / CV_32S are supported");
    if (dims != 2 && data) fail("Mat::convertTo: only 2-D matrices");
    const Mat src = *this;  // dst may be *this (ReadFile converts in place)
    Mat &dst = out_mat(_dst, src.rows, src.cols, CV_MAKETYPE(ddepth, cn));
    const bool plain = alpha == 1.0 && beta == 0.0;
    const size_t n = (size_t)src.cols * cn;
    for (int y = 0; y < src.rows; y++) {
        const uchar *s = src.data + (size_t)y * src.step.p[0];
        uchar *d = dst.data + (size_t)y * dst.step.p[0];
        if (sdepth == ddepth && plain) {
            if (s != d) std::memmove(d, s, n * 4);
        } else if (sdepth == CV_32S && ddepth == CV_32F) {
            for (size_t i = 0; i < n; i++)
                reinterpret_cast<float *>(d)[i] = plain ? (float)reinterpret_cast<const int *>(s)[i]
                                                        : (float)(reinterpret_cast<const int *>(s)[i] * alpha + beta);
        } else if (sdepth == CV_32F && ddepth == CV_32S) {
            for (size_t i = 0; i < n; i++)
                reinterpret_cast<int *>(d)[i] = plain ? saturate_cast<int>(reinterpret_cast<const float *>(s)[i])
                                                      : saturate_cast<int>(reinterpret_cast<const float *>(s)[i] * alpha + beta);
        } else if (sdepth == CV_32F) {
            for (size_t i = 0; i < n; i++)
                reinterpret_cast<float *>(d)[i] = (float)(reinterpret_cast<const float *>(s)[i] * alpha + beta);
        } else {
            for (size_t i = 0; i < n; i++)
                reinterpret_cast<int *>(d)[i] = saturate_cast<int>(reinterpret_cast<const int *>(s)[i] * alpha + beta);
        }
    }
}

// imgproc color.cpp, the two codes buffer.cpp:48 / statpath.cpp:452 use (COLOR_RGB2BGR == COLOR_BGR2RGB): swap channels 0 and 2
void cvtColor(InputArray _src, OutputArray _dst, int code, int) {
    if (code != COLOR_RGB2BGR) fail("cvtColor: only COLOR_RGB2BGR / COLOR_BGR2RGB are supported");
    const Mat src = in_mat(_src);
    if (src.type() != CV_32FC3) fail("cvtColor: only CV_32FC3 is supported");
    Mat &dst = out_mat(_dst, src.rows, src.cols, src.type());
    for (int y = 0; y < src.rows; y++) {
        const float *s = src.ptr<float>(y);
        float *d = dst.ptr<float>(y);
        for (int x = 0; x < src.cols; x++) {
            const float a = s[3 * x], b = s[3 * x + 1], c = s[3 * x + 2];
            d[3 * x] = c;
            d[3 * x + 1] = b;
            d[3 * x + 2] = a;
        }
    }
}

// core merge.dispatch.cpp for a std::vector<Mat> of CV_32F matrices of equal size (buffer.cpp:62, the display path)
void merge(InputArrayOfArrays _mv, OutputArray _dst) {
    if ((_mv.getFlags() & _InputArray::KIND_MASK) != _InputArray::STD_VECTOR_MAT) fail("merge: expects std::vector<cv::Mat>");
    const std::vector<Mat> &mv = *static_cast<const std::vector<Mat> *>(_mv.getObj());
    if (mv.empty()) fail("merge: empty input");
    int cn = 0;
    for (const Mat &m : mv) {
        if (m.depth() != CV_32F || m.rows != mv[0].rows || m.cols != mv[0].cols) fail("merge: CV_32F matrices of one size expected");
        cn += m.channels();
    }
    if (cn > CV_CN_MAX) fail("merge: too many channels");
    Mat &dst = out_mat(_dst, mv[0].rows, mv[0].cols, CV_MAKETYPE(CV_32F, cn));
    int c0 = 0;
    for (const Mat &m : mv) {
        const int mc = m.channels();
        for (int y = 0; y < m.rows; y++) {
            const float *s = m.ptr<float>(y);
            float *d = dst.ptr<float>(y) + c0;
            for (int x = 0; x < m.cols; x++)
                for (int c = 0; c < mc; c++) d[(size_t)x * cn + c] = s[(size_t)x * mc + c];
        }
        c0 += mc;
    }
}

// imgcodecs grfmt_pfm.cpp:77-258, the only format StatMC dumps statistics in.  A 3-channel cv::Mat is BGR by OpenCV's
// convention and the file holds RGB, bottom row first, little-endian, scale -1 (encoder :228-257, decoder :127-156).
bool imwrite(const String &filename, InputArray _img, const std::vector<int> &) {
    const size_t dot = filename.find_last_of('.');
    if (dot == String::npos || filename.substr(dot) != ".pfm") fail("imwrite: only .pfm is supported (" + filename + ")");
    const Mat img = in_mat(_img);
    if (img.depth() != CV_32F || (img.channels() != 1 && img.channels() != 3)) fail("imwrite: PFM needs CV_32FC1 or CV_32FC3");
    FILE *f = std::fopen(filename.c_str(), "wb");
    if (!f) return false;
    const int cn = img.channels();
    std::fprintf(f, "P%c\n%d %d\n-1.000000\n", cn == 3 ? 'F' : 'f', img.cols, img.rows);
    std::vector<float> row((size_t)img.cols * cn);
    bool ok = true;
    for (int y = img.rows - 1; y >= 0 && ok; --y) {
        const float *s = img.ptr<float>(y);
        if (cn == 3)
            for (int x = 0; x < img.cols; x++) {
                row[3 * x] = s[3 * x + 2];
                row[3 * x + 1] = s[3 * x + 1];
                row[3 * x + 2] = s[3 * x];
            }
        ok = std::fwrite(cn == 3 ? row.data() : s, 4, row.size(), f) == row.size();
    }
    return (std::fclose(f) == 0) && ok;
}

Mat imread(const String &filename, int) {
    Mat out;
    FILE *f = std::fopen(filename.c_str(), "rb");
    if (!f) return out;  // OpenCV returns an empty matrix when the file cannot be read
    char t = 0;
    int cols = 0, rows = 0;
    double scale = 0;
    if (std::fscanf(f, "P%c %d %d %lf", &t, &cols, &rows, &scale) != 4 || (t != 'F' && t != 'f') || cols <= 0 || rows <= 0 ||
        scale == 0 || std::fgetc(f) == EOF) {
        std::fclose(f);
        return out;
    }
    const int cn = t == 'F' ? 3 : 1;
    out = Mat(rows, cols, CV_MAKETYPE(CV_32F, cn));
    const bool swap = scale > 0;  // positive scale = big-endian file on this little-endian host
    const float inv = (float)(1.0 / (scale < 0 ? -scale : scale));
    bool ok = true;
    const size_t n = (size_t)cols * cn;
    for (int y = rows - 1; y >= 0 && ok; --y) {
        float *d = out.ptr<float>(y);
        ok = std::fread(d, 4, n, f) == n;
        if (swap)
            for (size_t i = 0; i < n; i++) {
                uint32_t u;
                std::memcpy(&u, d + i, 4);
                u = (u << 24) | ((u & 0xff00u) << 8) | ((u >> 8) & 0xff00u) | (u >> 24);
                std::memcpy(d + i, &u, 4);
            }
        if (inv != 1.f)
            for (size_t i = 0; i < n; i++) d[i] *= inv;
        if (cn == 3)
            for (int x = 0; x < cols; x++) std::swap(d[3 * x], d[3 * x + 2]);
    }
    std::fclose(f);
    if (!ok) out = Mat();
    return out;
}

// core glob.cpp, as StatPathIntegrator::Denoise uses it (statpath.cpp:479-481): the files matching a shell pattern, sorted
void glob(String pattern, std::vector<String> &result, bool recursive) {
    if (recursive) fail("glob: recursive search is not supported");
    result.clear();
    glob_t g;
    std::memset(&g, 0, sizeof(g));
    if (::glob(pattern.c_str(), 0, nullptr, &g) == 0)
        for (size_t i = 0; i < g.gl_pathc; i++) result.emplace_back(g.gl_pathv[i]);
    ::globfree(&g);
    std::sort(result.begin(), result.end());
}

#ifndef SMC_SHIM_HOST_ONLY
// ========================================================================================================================
namespace cuda {

GpuMat::Allocator *GpuMat::defaultAllocator() {
    static ShimAllocator *a = new ShimAllocator;
    return a;
}

void GpuMat::create(int _rows, int _cols, int _type) {  // gpu_mat.cu / cuda_gpu_mat.cpp GpuMat::create
    if (_rows < 0 || _cols < 0) fail("negative GpuMat size");
    _type &= Mat::TYPE_MASK;
    if (rows == _rows && cols == _cols && type() == _type && data) return;
    if (data) release();
    if (_rows > 0 && _cols > 0) {
        flags = Mat::MAGIC_VAL + _type;
        rows = _rows;
        cols = _cols;
        const size_t esz = elemSize();
        if (!allocator) allocator = defaultAllocator();
        if (!allocator->allocate(this, rows, cols, esz)) fail("GpuMat::create: device allocation failed");
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

// gpu_mat.cu:224-234: create(size, type) + cudaMemcpy2DAsync on the stream.  An empty source (the 1 x 0 descriptor tables
// AllocateBuffers builds for unused groups, estimator.cpp:43-46) is a no-op.
void GpuMat::upload(InputArray arr, Stream &) {
    const Mat &m = in_mat(arr);
    if (m.rows <= 0 || m.cols <= 0 || !m.data) return;
    create(m.rows, m.cols, m.type());
    const Plane2D p = plane_of(*this);
    const size_t dev_row = ((p.row_bytes + 3) / 4) * 4;  // byte matrices live in whole words on the device (ShimAllocator)
    pending_drop(datastart);
    Shim &s = shim();
    if (s.defer && dev_row == p.row_bytes && m.u && m.u->userdata && p.row_bytes * (size_t)m.rows >= kPinnedFrom) {
        std::lock_guard<std::mutex> g(s.mu);
        s.pending.push_back(Pending{p.buf, data, step, m.data, m.step.p[0], p.row_bytes, m.rows, m.u});
    } else if (dev_row == p.row_bytes) {
        check(smc_buffer_upload(p.buf, m.data, m.step.p[0]), "GpuMat::upload");
    } else {  // pad the rows on the way (pageable source: the copy is staged before the call returns)
        std::vector<uchar> tmp(dev_row * (size_t)m.rows, 0);
        for (int y = 0; y < m.rows; y++) std::memcpy(tmp.data() + dev_row * y, m.data + m.step.p[0] * y, p.row_bytes);
        check(smc_buffer_upload(p.buf, tmp.data(), dev_row), "GpuMat::upload");
    }
}
void GpuMat::upload(InputArray arr) {
    upload(arr, Stream::Null());
    pending_flush();
    check(smc_synchronize(shim().ctx), "GpuMat::upload");
}

void GpuMat::download(OutputArray _dst, Stream &) const {
    if (!data) fail("GpuMat::download of an empty matrix");
    pending_flush();
    Mat &dst = out_mat(_dst, rows, cols, type());
    const Plane2D p = plane_of(*this);
    const size_t dev_row = ((p.row_bytes + 3) / 4) * 4;
    if (dev_row == p.row_bytes) {
        check(smc_buffer_download(p.buf, dst.data, dst.step.p[0]), "GpuMat::download");
    } else {
        std::vector<uchar> tmp(dev_row * (size_t)rows);
        check(smc_buffer_download(p.buf, tmp.data(), dev_row), "GpuMat::download");
        check(smc_synchronize(shim().ctx), "GpuMat::download");
        for (int y = 0; y < rows; y++) std::memcpy(dst.data + dst.step.p[0] * y, tmp.data() + dev_row * y, p.row_bytes);
    }
}
void GpuMat::download(OutputArray _dst) const {
    download(_dst, Stream::Null());
    check(smc_synchronize(shim().ctx), "GpuMat::download");
}

// ---- Stream: every Stream is the context's stream (the reference uses exactly one, estimator.h:326) ------------------------
class Stream::Impl {
public:
    smc_context *ctx = shim().ctx;
};
Stream::Stream() : impl_(new Impl) {}
Stream::Stream(const Ptr<GpuMat::Allocator> &) : impl_(new Impl) {}
Stream::Stream(const size_t) : impl_(new Impl) {}
Stream &Stream::Null() {
    static Stream *s = new Stream;
    return *s;
}
void Stream::waitForCompletion() {
    pending_flush();
    check(smc_synchronize(impl_->ctx), "Stream::waitForCompletion");
}
void *Stream::cudaPtr() const { return smc_context_stream(impl_->ctx); }

// ---- the denoiser -----------------------------------------------------------------------------------------------------
#ifdef SMC_SHIM_REF_KERNELS
}  // namespace cuda
}  // namespace cv
// Measurement variant (oracle/Makefile target _ref/pbrt_ref_refkernels): the Estimator still allocates and copies through
// libstatmc_b200, but stat_denoiser::filter<T> runs the REFERENCE'S OWN kernels (stat_denoiser.cu compiled unmodified,
// oracle/_ref/stat_denoiser.o), so that the reference's "CUDA time [ns]" line times its kernels and ours on the same data,
// through the same Upload / Denoise / Download / Synchronize sequence.
struct CUstream_st;
namespace cv { namespace cuda { namespace device { namespace imgproc { namespace stat_denoiser {
template <typename T>
void filter(const unsigned short ptrCount, const unsigned short width, const unsigned short height, const float dSFactor,
            const unsigned char radius, const bool denoiseFilm, const PtrStepSzb &nPtrs, const PtrStepSzb &meanPtrs,
            const PtrStepSzb &m2Ptrs, const PtrStepSzb &m3Ptrs, const PtrStepSzb &filmPtrs, const PtrStepSzb &film,
            const PtrStepSzb &gBufPtrs, const PtrStepSzb &gBufChannelCounts, const PtrStepSzb &gBufDRFactors,
            const unsigned char nGBufs, PtrStepSzb meanCorrPtrs, PtrStepSzb discriminatorPtrs, PtrStepSzb filmFilteredPtrs,
            PtrStepSzb filmFiltered, CUstream_st *stream);
}}}}}
namespace cv {
void error(int code, const std::string &err, const char *func, const char *file, int line) {
    std::fprintf(stderr, "[statmc ref kernels] cv::error %d: %s in %s (%s:%d)\n", code, err.c_str(), func, file, line);
    std::abort();
}
namespace cuda {
#endif

namespace stat_denoiser {

namespace {
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

void setup() { (void)shim(); }  // SD.cu:352-355 sets a malloc-heap limit and a cache preference; nothing of that is needed

void synchronize(Stream &) {
    pending_flush();
    check(smc_synchronize(shim().ctx), "stat_denoiser::synchronize");
}

template <typename T>
void calculateMeanVars(const unsigned short ptrCount, const unsigned short width, const unsigned short height,
                       const PtrStepSzb &nPtrs, const PtrStepSzb &m2Ptrs, PtrStepSzb meanVarPtrs, Stream &) {
    Shim &s = shim();
    pending_flush();
    check(smc_calculate_mean_vars_device_tables(s.ctx, ChannelsOf<T>::value, ptrCount, width, height, nPtrs.data, m2Ptrs.data,
                                                meanVarPtrs.data, smc_context_stream(s.ctx)),
          "stat_denoiser::calculateMeanVars");
}

template <typename T>
void filter(const unsigned short ptrCount, const unsigned short width, const unsigned short height, const float dSFactor,
            const unsigned char radius, const bool denoiseFilm, const PtrStepSzb &nPtrs, const PtrStepSzb &meanPtrs,
            const PtrStepSzb &m2Ptrs, const PtrStepSzb &m3Ptrs, const PtrStepSzb &filmPtrs, const PtrStepSzb &film,
            const PtrStepSzb &gBufferPtrs, const PtrStepSzb &gBufferChannelCounts, const PtrStepSzb &gBufferDRFactors,
            const unsigned char nGBufs, PtrStepSzb meanCorrPtrs, PtrStepSzb discriminatorPtrs, PtrStepSzb filmFilteredPtrs,
            PtrStepSzb filmFiltered, Stream &) {
    Shim &s = shim();
#ifdef SMC_SHIM_REF_KERNELS
    pending_flush();
    device::imgproc::stat_denoiser::filter<T>(ptrCount, width, height, dSFactor, radius, denoiseFilm, nPtrs, meanPtrs, m2Ptrs,
                                              m3Ptrs, filmPtrs, film, gBufferPtrs, gBufferChannelCounts, gBufferDRFactors, nGBufs,
                                              meanCorrPtrs, discriminatorPtrs, filmFilteredPtrs, filmFiltered,
                                              (CUstream_st *)smc_context_stream(s.ctx));
    return;
#endif
    // uploads held back since Estimator::Upload travel inside the filter's row-chunked pipeline
    std::vector<smc_host_rows> ups;
    for (const Pending &p : pending_take()) ups.push_back(smc_host_rows{p.dev, p.dev_step, p.host, p.host_step, p.row_bytes, p.rows});
    check(smc_filter_device_tables_host(s.ctx, ChannelsOf<T>::value, ptrCount, width, height, dSFactor, radius, denoiseFilm,
                                        nPtrs.data, meanPtrs.data, m2Ptrs.data, m3Ptrs.data, filmPtrs.data, film.data, film.step,
                                        gBufferPtrs.data, gBufferChannelCounts.data, gBufferDRFactors.data, nGBufs,
                                        meanCorrPtrs.data, discriminatorPtrs.data, filmFilteredPtrs.data, filmFiltered.data,
                                        filmFiltered.step, smc_context_stream(s.ctx), ups.data(), (int)ups.size()),
          "stat_denoiser::filter");
}

template <typename T>
void filter(const unsigned short ptrCount, const unsigned short width, const unsigned short height, const float dSFactor,
            const unsigned char radius, const PtrStepSzb &nPtrs, const PtrStepSzb &meanPtrs, const PtrStepSzb &m2Ptrs,
            const PtrStepSzb &m3Ptrs, const PtrStepSzb &filmPtrs, const PtrStepSzb &gBufferPtrs,
            const PtrStepSzb &gBufferChannelCounts, const PtrStepSzb &gBufferDRFactors, const unsigned char nGBufs,
            PtrStepSzb meanCorrPtrs, PtrStepSzb discriminatorPtrs, PtrStepSzb filmFilteredPtrs, Stream &stream) {
    filter<T>(ptrCount, width, height, dSFactor, radius, false, nPtrs, meanPtrs, m2Ptrs, m3Ptrs, filmPtrs, PtrStepSzb(),
              gBufferPtrs, gBufferChannelCounts, gBufferDRFactors, nGBufs, meanCorrPtrs, discriminatorPtrs, filmFilteredPtrs,
              PtrStepSzb(), stream);
}

#define SMC_SHIM_INSTANTIATE(T)                                                                                                \
    template void calculateMeanVars<T>(const unsigned short, const unsigned short, const unsigned short, const PtrStepSzb &,  \
                                       const PtrStepSzb &, PtrStepSzb, Stream &);                                             \
    template void filter<T>(const unsigned short, const unsigned short, const unsigned short, const float, const unsigned char, \
                            const bool, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &,       \
                            const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, \
                            const unsigned char, PtrStepSzb, PtrStepSzb, PtrStepSzb, PtrStepSzb, Stream &);                   \
    template void filter<T>(const unsigned short, const unsigned short, const unsigned short, const float, const unsigned char, \
                            const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, \
                            const PtrStepSzb &, const PtrStepSzb &, const PtrStepSzb &, const unsigned char, PtrStepSzb,      \
                            PtrStepSzb, PtrStepSzb, Stream &);
SMC_SHIM_INSTANTIATE(float)
SMC_SHIM_INSTANTIATE(::float3)
#undef SMC_SHIM_INSTANTIATE

}  // namespace stat_denoiser
}  // namespace cuda
#endif  // SMC_SHIM_HOST_ONLY
}  // namespace cv
