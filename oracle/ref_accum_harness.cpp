/*
 * ref_accum_harness.cpp -- C-ABI shim around the REFERENCE's own moment accumulation, compiled unmodified.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/statmc_oracle.c header).  This file is ours; the code it calls is
 * pbrt::StatTile<T>::Add[Transform]SampleM{1,2,3} in /root/reference/src/statistics/estimator.h:162-232, included
 * where it lies by oracle/Makefile (-I $(REFERENCE)/src) and built into oracle/_ref/ (git-ignored).  No reference
 * source is copied.  estimator.h pulls in pbrt.h / core/film.h / OpenCV headers; only inline and template code of
 * those is used (cv::Vec element-wise operators, Bounds2i/Point2i), so nothing of pbrt or OpenCV is linked.  The
 * one generated header that does not exist in the checkout (<glog/logging.h>, made by glog's CMake) is replaced by
 * ref_stubs_accum/glog/logging.h, which turns the CHECK/LOG macros into no-ops.
 *
 * What it does is what StatPathIntegrator::Render does per tile (src/statistics/statpath.cpp:166-190, 357-371):
 * hold a StatTile over the pixel block, feed it the samples of every pixel in order through the member function the
 * configuration selects (radiance: transform + M3, statpath.cpp:1042-1046; features: M1, no transform, :1117-1118),
 * and read the running totals back (Estimator::Merge[Transform]Tile, src/statistics/estimator.cpp:341-388).
 *
 * Sample layout [S][npix][C], state planes [npix][C] -- the same as smo_accumulate() in statmc_oracle.c.
 */
#include <cstdint>
#include <cstring>

#include "statistics/estimator.h"

namespace {

template <typename T>
struct Ch;
template <>
struct Ch<pbrt::Float> {
    static constexpr int C = 1;
    static pbrt::Float load(const float *p) { return p[0]; }
    static void store(float *p, const pbrt::Float &v) { p[0] = v; }
};
template <>
struct Ch<pbrt::Vec3> {
    static constexpr int C = 3;
    static pbrt::Vec3 load(const float *p) { return pbrt::Vec3(p[0], p[1], p[2]); }
    static void store(float *p, const pbrt::Vec3 &v) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; }
};

template <typename T>
int run(int64_t npix, int S, const float *samples, int transform, int max_moment, int64_t *n, float *mean, float *m2,
        float *m3, float *film_mean, float *film_m2) {
    using namespace pbrt;
    constexpr int C = Ch<T>::C;
    typedef void (StatTile<T>::*Add)(const Point2i, const T);
    Add add = nullptr;
    if (transform)
        add = max_moment >= 3 ? &StatTile<T>::AddTransformSampleM3
            : max_moment == 2 ? &StatTile<T>::AddTransformSampleM2 : &StatTile<T>::AddTransformSampleM1;
    else
        add = max_moment >= 3 ? &StatTile<T>::AddSampleM3
            : max_moment == 2 ? &StatTile<T>::AddSampleM2 : &StatTile<T>::AddSampleM1;
    const int64_t block = 1 << 16;  // one "tile" = a 1-row strip of up to 65536 pixels
    for (int64_t p0 = 0; p0 < npix; p0 += block) {
        const int w = (int)((npix - p0 < block) ? npix - p0 : block);
        StatTile<T> tile(Bounds2i(Point2i(0, 0), Point2i(w, 1)));
        for (int x = 0; x < w; x++) {  // the tile persists across iterations in the reference: restore its state
            StatTilePixel<T> &px = tile.GetPixel(Point2i(x, 0));
            const size_t i = (size_t)(p0 + x) * C;
            px.n = (uint64_t)n[p0 + x];
            px.mean = Ch<T>::load(mean + i);
            px.m2 = Ch<T>::load(m2 + i);
            px.m3 = Ch<T>::load(m3 + i);
            px.filmMean = Ch<T>::load(film_mean + i);
            px.filmM2 = Ch<T>::load(film_m2 + i);
        }
        for (int s = 0; s < S; s++)
            for (int x = 0; x < w; x++)
                (tile.*add)(Point2i(x, 0), Ch<T>::load(samples + ((size_t)s * (size_t)npix + (size_t)(p0 + x)) * C));
        for (int x = 0; x < w; x++) {
            const StatTilePixel<T> &px = tile.GetPixel(Point2i(x, 0));
            const size_t i = (size_t)(p0 + x) * C;
            n[p0 + x] = (int64_t)px.n;
            Ch<T>::store(mean + i, px.mean);
            Ch<T>::store(m2 + i, px.m2);
            Ch<T>::store(m3 + i, px.m3);
            Ch<T>::store(film_mean + i, px.filmMean);
            Ch<T>::store(film_m2 + i, px.filmM2);
        }
    }
    return 0;
}

}  // namespace

extern "C" int smr_accumulate(int64_t npix, int C, int nsamples, const float *samples, int transform, int max_moment,
                              int64_t *n, float *mean, float *m2, float *m3, float *film_mean, float *film_m2) {
    if (C == 3) return run<pbrt::Vec3>(npix, nsamples, samples, transform, max_moment, n, mean, m2, m3, film_mean, film_m2);
    if (C == 1) return run<pbrt::Float>(npix, nsamples, samples, transform, max_moment, n, mean, m2, m3, film_mean, film_m2);
    return -1;
}

/* sizeof/alignof of the reference's accumulator record (estimator.h:104-124), for the layout test */
extern "C" int smr_tile_pixel_layout(int C, int *size, int *align) {
    if (C == 3) { *size = (int)sizeof(pbrt::StatTilePixel<pbrt::Vec3>); *align = (int)alignof(pbrt::StatTilePixel<pbrt::Vec3>); return 0; }
    if (C == 1) { *size = (int)sizeof(pbrt::StatTilePixel<pbrt::Float>); *align = (int)alignof(pbrt::StatTilePixel<pbrt::Float>); return 0; }
    return -1;
}

/* boxCox(val, lambda), estimator.h:135-137 */
extern "C" float smr_box_cox(float v, float lambda) { return pbrt::boxCox(v, lambda); }
