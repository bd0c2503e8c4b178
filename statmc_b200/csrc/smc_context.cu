// smc_context.cu -- context, buffers, pinned memory, Student-t tables.
// Replaces cv::cuda::stat_denoiser::setup/synchronize (stat_denoiser.cu:352-359), the Estimator's cv::cuda::Stream
// (estimator.h:326) and Buffer's GpuMat alloc/upload/download (buffer.h:24-63) for the hot path.
#include <cmath>
#include <cstring>
#include <new>

#include "smc_internal.h"
#include "t_quantile_tables.inc"

static thread_local char g_err[512] = "";

void smc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *smc_last_error(void) { return g_err; }
extern "C" int smc_version(void) { return SMC_VERSION; }

// ---------------------------------------------------------------------------------------------------------
// Student-t distribution in double precision (host).  Used to build quantile tables for arbitrary alpha;
// the nine levels the reference ships are served from the generated tables (bit-identical to its text).
// ---------------------------------------------------------------------------------------------------------
namespace {

// continued fraction for the regularised incomplete beta function (modified Lentz)
double betacf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 100000; m++) {
        const int m2 = 2 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < eps) break;
    }
    return h;
}

double betai(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    const double lbt = std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b) + a * std::log(x) + b * std::log1p(-x);
    const double bt = std::exp(lbt);
    if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
    return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

// upper tail P(T > t), t >= 0
double t_tail(double t, double df) {
    const double x = df / (df + t * t);
    return 0.5 * betai(0.5 * df, 0.5, x);
}

double t_pdf(double t, double df) {
    const double l = std::lgamma(0.5 * (df + 1.0)) - std::lgamma(0.5 * df) - 0.5 * std::log(df * M_PI) -
                     0.5 * (df + 1.0) * std::log1p(t * t / df);
    return std::exp(l);
}

}  // namespace

extern "C" double smc_t_cdf(double t, double df) {
    if (!(df > 0.0)) return NAN;
    const double tail = t_tail(std::fabs(t), df);
    return t >= 0 ? 1.0 - tail : tail;
}

extern "C" double smc_t_quantile(double p, double df) {
    if (!(p > 0.0 && p < 1.0) || !(df > 0.0)) return NAN;
    if (p == 0.5) return 0.0;
    const bool neg = p < 0.5;
    const double q = neg ? p : 1.0 - p;  // upper-tail mass of |t|
    // bracket: tail is decreasing in t
    double lo = 0.0, hi = 1.0;
    while (t_tail(hi, df) > q) {
        lo = hi;
        hi *= 2.0;
        if (hi > 1e300) break;
    }
    double t = 0.5 * (lo + hi);
    for (int it = 0; it < 200; it++) {
        const double f = t_tail(t, df) - q;
        if (f > 0) lo = t; else hi = t;
        const double pdf = t_pdf(t, df);
        double tn = pdf > 0 ? t + f / pdf : 0.5 * (lo + hi);  // d tail / dt = -pdf
        if (!(tn > lo && tn < hi)) tn = 0.5 * (lo + hi);
        if (std::fabs(tn - t) <= 1e-16 * std::fabs(tn)) {
            t = tn;
            break;
        }
        t = tn;
    }
    return neg ? -t : t;
}

static void build_table(double alpha, float *out) {
    for (int k = 0; k < SMC_T_NUM_TABLES; k++)
        if (alpha == smc_t_table_alphas[k]) {
            std::memcpy(out, smc_t_tables[k], sizeof(float) * SMC_T_LUT_SIZE);
            return;
        }
    for (int i = 0; i < SMC_T_LUT_SIZE; i++) out[i] = (float)smc_t_quantile(1.0 - 0.5 * alpha, (double)(i + 1));
}

// ---------------------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------------------
static int context_create(int device, cudaStream_t stream, bool own, smc_context **out) {
    if (!out) SMC_FAIL(SMC_ERR_INVALID, "smc_context_create: out == NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        SMC_FAIL(SMC_ERR_CUDA, "no CUDA device available (%s); libstatmc_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) SMC_FAIL(SMC_ERR_INVALID, "device %d out of range [0, %d)", device, count);
    SMC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SMC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        SMC_FAIL(SMC_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                 prop.minor);
    smc_context *ctx = new (std::nothrow) smc_context;
    if (!ctx) SMC_FAIL(SMC_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->own_stream = own;
    ctx->stream = stream;
    if (own) {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            SMC_FAIL(SMC_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
        }
    }
    e = cudaMalloc(&ctx->d_lut, sizeof(float) * SMC_T_LUT_ENTRIES);
    if (e != cudaSuccess) {
        if (own) cudaStreamDestroy(ctx->stream);
        delete ctx;
        SMC_FAIL(SMC_ERR_NOMEM, "cudaMalloc(lut) failed: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return smc_set_alpha(ctx, 0.005);  // t_005_quantiles is the table the reference compiles in (stat_denoiser.cu:56,67)
}

extern "C" int smc_context_create(int device, smc_context **out) { return context_create(device, nullptr, true, out); }

extern "C" int smc_context_create_on_stream(int device, void *cuda_stream, smc_context **out) {
    return context_create(device, (cudaStream_t)cuda_stream, false, out);
}

extern "C" void smc_context_destroy(smc_context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (smc_denoiser *c : ctx->cached)
        if (c) smc_denoiser_destroy(c);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_accum_fallback);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int smc_synchronize(smc_context *ctx) {
    if (!ctx) SMC_FAIL(SMC_ERR_INVALID, "ctx == NULL");
    SMC_CUDA(cudaStreamSynchronize(ctx->stream));
    SMC_CUDA(cudaGetLastError());
    return SMC_OK;
}

extern "C" void *smc_context_stream(smc_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int smc_context_device(smc_context *ctx) { return ctx ? ctx->device : -1; }
extern "C" uint64_t smc_context_launch_count(smc_context *ctx) { return ctx ? ctx->launches : 0; }

extern "C" uint64_t smc_accumulate_fallback_samples(smc_context *ctx) {
    if (!ctx || !ctx->d_accum_fallback) return 0;
    unsigned long long v = 0;
    cudaSetDevice(ctx->device);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
    if (cudaMemcpy(&v, ctx->d_accum_fallback, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    cudaMemset(ctx->d_accum_fallback, 0, sizeof(v));
    return v;
}

extern "C" int smc_set_alpha(smc_context *ctx, double alpha) {
    if (!ctx) SMC_FAIL(SMC_ERR_INVALID, "ctx == NULL");
    if (!(alpha > 0.0 && alpha < 1.0)) SMC_FAIL(SMC_ERR_INVALID, "alpha must be in (0,1), got %g", alpha);
    build_table(alpha, ctx->h_lut);
    ctx->alpha = alpha;
    SMC_CUDA(cudaSetDevice(ctx->device));
    // stream-ordered so that a table change never races with kernels already queued
    SMC_CUDA(cudaMemcpyAsync(ctx->d_lut, ctx->h_lut, sizeof(float) * SMC_T_LUT_ENTRIES, cudaMemcpyHostToDevice,
                             ctx->stream));
    SMC_CUDA(cudaStreamSynchronize(ctx->stream));  // h_lut is pageable; keep it simple and safe
    return SMC_OK;
}

extern "C" double smc_get_alpha(smc_context *ctx) { return ctx ? ctx->alpha : NAN; }

extern "C" int smc_get_t_table(smc_context *ctx, float *out) {
    if (!ctx || !out) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    std::memcpy(out, ctx->h_lut, sizeof(float) * SMC_T_LUT_ENTRIES);
    return SMC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Buffers
// ---------------------------------------------------------------------------------------------------------
extern "C" int smc_buffer_create(smc_context *ctx, int rows, int cols, int channels, int dtype, smc_buffer **out) {
    if (!ctx || !out) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (rows <= 0 || cols <= 0 || channels <= 0 || channels > 512)
        SMC_FAIL(SMC_ERR_INVALID, "bad plane shape %d x %d x %d", rows, cols, channels);
    if (dtype != SMC_F32 && dtype != SMC_I32) SMC_FAIL(SMC_ERR_INVALID, "bad dtype %d", dtype);
    SMC_CUDA(cudaSetDevice(ctx->device));
    smc_buffer *b = new (std::nothrow) smc_buffer;
    if (!b) SMC_FAIL(SMC_ERR_NOMEM, "out of host memory");
    b->ctx = ctx;
    b->rows = rows;
    b->cols = cols;
    b->channels = channels;
    b->dtype = dtype;
    b->elem_bytes = (size_t)channels * 4;
    const size_t row_bytes = (size_t)cols * b->elem_bytes;
    // GpuMat: pitched when rows > 1 && cols > 1 (gpu_mat.cu:112-123), contiguous otherwise
    b->step = (rows > 1 && cols > 1) ? ((row_bytes + 255) / 256) * 256 : row_bytes;
    cudaError_t e = cudaMalloc(&b->dev, b->step * (size_t)rows);
    if (e != cudaSuccess) {
        delete b;
        SMC_FAIL(SMC_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", b->step * (size_t)rows, cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(b->dev, 0, b->step * (size_t)rows, ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(b->dev);
        delete b;
        SMC_FAIL(SMC_ERR_CUDA, "cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    }
    *out = b;
    return SMC_OK;
}

extern "C" void smc_buffer_destroy(smc_buffer *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaFree(b->dev);
    delete b;
}

extern "C" void *smc_buffer_dev(const smc_buffer *b) { return b ? b->dev : nullptr; }
extern "C" size_t smc_buffer_step(const smc_buffer *b) { return b ? b->step : 0; }
extern "C" smc_plane smc_buffer_plane(const smc_buffer *b) {
    smc_plane p = {b ? b->dev : nullptr, b ? b->step : 0};
    return p;
}

static int copy_rows(const smc_buffer *b, int row0, int nrows, void *host, size_t host_step, bool up) {
    if (!b || !host) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    if (row0 < 0 || nrows < 0 || row0 + nrows > b->rows)
        SMC_FAIL(SMC_ERR_INVALID, "row range [%d, %d) outside plane of %d rows", row0, row0 + nrows, b->rows);
    if (nrows == 0) return SMC_OK;
    const size_t row_bytes = (size_t)b->cols * b->elem_bytes;
    if (host_step == 0) host_step = row_bytes;
    if (host_step < row_bytes) SMC_FAIL(SMC_ERR_INVALID, "host pitch %zu < row bytes %zu", host_step, row_bytes);
    SMC_CUDA(cudaSetDevice(b->ctx->device));
    char *d = (char *)b->dev + (size_t)row0 * b->step;
    if (up)
        SMC_CUDA(cudaMemcpy2DAsync(d, b->step, host, host_step, row_bytes, nrows, cudaMemcpyHostToDevice,
                                   b->ctx->stream));
    else
        SMC_CUDA(cudaMemcpy2DAsync(host, host_step, d, b->step, row_bytes, nrows, cudaMemcpyDeviceToHost,
                                   b->ctx->stream));
    return SMC_OK;
}

extern "C" int smc_buffer_upload(smc_buffer *b, const void *host, size_t host_step) {
    return copy_rows(b, 0, b ? b->rows : 0, (void *)host, host_step, true);
}
extern "C" int smc_buffer_download(const smc_buffer *b, void *host, size_t host_step) {
    return copy_rows(b, 0, b ? b->rows : 0, host, host_step, false);
}
extern "C" int smc_buffer_upload_rows(smc_buffer *b, int row0, int nrows, const void *host, size_t host_step) {
    return copy_rows(b, row0, nrows, (void *)host, host_step, true);
}
extern "C" int smc_buffer_download_rows(const smc_buffer *b, int row0, int nrows, void *host, size_t host_step) {
    return copy_rows(b, row0, nrows, host, host_step, false);
}

extern "C" int smc_buffer_fill_zero(smc_buffer *b) {
    if (!b) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    SMC_CUDA(cudaSetDevice(b->ctx->device));
    SMC_CUDA(cudaMemsetAsync(b->dev, 0, b->step * (size_t)b->rows, b->ctx->stream));
    return SMC_OK;
}

extern "C" int smc_memcpy_device(smc_context *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx || !dst || !src) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    SMC_CUDA(cudaSetDevice(ctx->device));
    SMC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return SMC_OK;
}

extern "C" int smc_host_alloc(size_t bytes, void **out) {
    if (!out) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) SMC_FAIL(SMC_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return SMC_OK;
}
extern "C" void smc_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
extern "C" int smc_host_register(void *p, size_t bytes) {
    if (!p) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    SMC_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    return SMC_OK;
}
extern "C" int smc_host_unregister(void *p) {
    if (!p) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    SMC_CUDA(cudaHostUnregister(p));
    return SMC_OK;
}
