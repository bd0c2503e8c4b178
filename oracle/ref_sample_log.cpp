/* ref_sample_log.cpp -- sink of the sample-logging build of the reference's renderer (see
 * ref_stubs_samplelog/samplelog_hook.h).  TEST INFRASTRUCTURE ONLY.
 * Environment: STATMC_SAMPLE_LOG=<file>, STATMC_SAMPLE_LOG_W / _H / _S = film width, height and samples per pixel in
 * total.  The file written at exit holds float32 [S][H][W][3]: sample s of a pixel is the one added when its tile pixel
 * had n == s (image tiles are disjoint, so threads never write the same element). */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
struct Log {
    int W = 0, H = 0, S = 0;
    std::vector<float> data;
    const char *path = nullptr;
    long long dropped = 0;
    Log() {
        path = std::getenv("STATMC_SAMPLE_LOG");
        const char *w = std::getenv("STATMC_SAMPLE_LOG_W"), *h = std::getenv("STATMC_SAMPLE_LOG_H"),
                   *s = std::getenv("STATMC_SAMPLE_LOG_S");
        if (path && w && h && s) {
            W = std::atoi(w);
            H = std::atoi(h);
            S = std::atoi(s);
            data.assign((size_t)W * H * S * 3, 0.f);
        }
    }
    ~Log() {
        if (!path || data.empty()) return;
        FILE *f = std::fopen(path, "wb");
        if (!f) return;
        std::fwrite(data.data(), 4, data.size(), f);
        std::fclose(f);
        if (dropped) std::fprintf(stderr, "ref_sample_log: %lld samples outside the declared W x H x S\n", dropped);
    }
};
Log &log() {
    static Log l;
    return l;
}
struct Init {  // construct before the render threads start
    Init() { (void)log(); }
} init;
}  // namespace

extern "C" void smr_log_sample(int x, int y, unsigned long long n_before, const float *rgb) {
    Log &l = log();
    if (l.data.empty()) return;
    if (x < 0 || y < 0 || x >= l.W || y >= l.H || n_before >= (unsigned long long)l.S) {
        __atomic_fetch_add(&l.dropped, 1, __ATOMIC_RELAXED);
        return;
    }
    float *d = l.data.data() + ((((size_t)n_before * l.H + y) * l.W) + x) * 3;
    d[0] = rgb[0];
    d[1] = rgb[1];
    d[2] = rgb[2];
}
