/*
 * ref_estimator_harness.cpp -- drives the REFERENCE's own host code (pbrt::Estimator / Buffer / BufferRegistry /
 * OutputBufferSelection, src/statistics/estimator.{h,cpp} and buffer.{h,cpp}, compiled UNMODIFIED from where they lie
 * under /root/reference by oracle/Makefile) on top of libstatmc_b200.so through integration/opencv_link_shim.cpp.
 *
 * TEST INFRASTRUCTURE ONLY.  It proves the link-level drop-in of INTEGRATION.md section 2(0): the reference's
 * AllocateBuffers / Upload / Denoise / Download / Synchronize run as they are, and every GpuMat, every upload and the
 * cv::cuda::stat_denoiser::filter<T> call they make land in our C ABI.  This file is ours; it does what
 * StatPathIntegrator does around the estimator, citing it:
 *   configs          CreateStatPathIntegrator, src/statistics/statpath.cpp:1026-1160
 *   construction     StatPathIntegrator::StatPathIntegrator, statpath.cpp:57-83
 *   filling planes   Estimator::Merge[Transform]Tile, estimator.cpp:341-388 / StatPathIntegrator::ReadFile, statpath.cpp:449-454
 *   the CUDA section statpath.cpp:406-418 ("CUDA time [ns]")
 *   the dump         OutputBufferSelection::PrepareOutput + Write, statpath.cpp:421-424, buffer.cpp:34-53
 * Outputs go to oracle/_ref/ (git-ignored).  No reference source is copied.
 */
#include <chrono>
#include <cstdint>
#include <cstring>
#include <regex>
#include <string>

#include "statistics/estimator.h"
#include "statistics/statpath.h"

// buffer.cpp:56-70 references the pbrt-v4 display server client (src/display/); the display path is out of scope
namespace pbrtv4 {
void DisplayStatic(std::string, unsigned short, unsigned short, const float *, std::vector<std::string>) {}
}  // namespace pbrtv4

namespace {

thread_local std::string g_err;

void put(pbrt::Buffer &b, const void *src, size_t bytes) {  // what MergeTile does through matPtr, for a whole plane
    if ((size_t)(b.mat.dataend - b.mat.datastart) != bytes) throw std::runtime_error("plane size mismatch for " + b.name);
    std::memcpy(b.mat.data, src, bytes);
}
void get(const cv::Mat &m, void *dst, size_t bytes) {
    if ((size_t)(m.dataend - m.datastart) != bytes) throw std::runtime_error("output plane size mismatch");
    std::memcpy(dst, m.data, bytes);
}

}  // namespace

extern "C" const char *smr_estimator_last_error() { return g_err.c_str(); }

/*
 * One denoise of `nb` radiance bounce images (nb = 1 unless ACRR tracks bounces) with two RGB G-buffers, through the
 * reference's Estimator.  C = 3 (multichannelstats=true, filter<float3>) or 1 (filter<float>).
 *   n        [nb][H][W] int32;   mean, m2, m3, film_mean: [nb][H][W][C] float32 (film_mean = "t0-b<j>-film-mean")
 *   film     [H][W][3] (pbrt film), normal / albedo [H][W][3]
 *   film_f   [H][W][3] out ("film-f"); film_mean_f [nb][H][W][C] out ("t0-b<j>-film-mean-f"); mean_corr, disc: [nb][H][W][C] out
 *   dump_stem  if non-NULL, OutputBufferSelection(reg, regex(dump_regex), dump_stem + ".pfm").Write(dump_suffix)
 *   reps       how many times the Upload/Denoise/Download/Synchronize section runs; *cuda_time_ns = the last one's time
 *   nbm > 0    statistical MIS (enableSMIS, statpath.cpp:1056-1080): BSDF and light win-rate statistics, scalar, untransformed,
 *              nbm tracked bounces each: mis_n [2][nbm][H][W] int32, mis_mean / mis_m2 / mis_m3 [2][nbm][H][W];
 *              mis_f [2][nbm][H][W] out ("t1-b<j>-film-mean-f", "t2-b<j>-film-mean-f").  With C = 3 this populates BOTH CUDA
 *              groups, so Estimator::Denoise launches filter<float> and then filter<float3> (estimator.cpp:434-488).
 */
extern "C" int smr_estimator_denoise(int W, int H, int C, int nb, float filterSD, int radius, int denoiseFilm, int acrr,
                                     const int32_t *n, const float *mean, const float *m2, const float *m3,
                                     const float *film_mean, const float *film, const float *normal, float normalSD,
                                     const float *albedo, float albedoSD, float *film_f, float *film_mean_f, float *mean_corr,
                                     float *disc, const char *dump_stem, const char *dump_regex, const char *dump_suffix,
                                     int reps, double *cuda_time_ns, int *n_registered, int nbm, const int32_t *mis_n,
                                     const float *mis_mean, const float *mis_m2, const float *mis_m3, float *mis_f) {
    using namespace pbrt;
    try {
        if ((C != 1 && C != 3) || nb < 1 || W < 1 || H < 1) throw std::runtime_error("bad shape");
        // ---- CreateStatPathIntegrator: denoiseimage=true, filterbuffers "normal" "albedo" ----------------------------
        StatTypeConfigs cfgs;
        cfgs.configs.resize(8);
        {
            auto &cfg = cfgs[Radiance];
            cfg.type = Radiance;
            cfg.index = cfgs.nEnabled++;
            cfg.enable = true;
            cfg.bounceStart = 0;
            cfg.bounceEnd = (unsigned char)nb;
            cfg.nBounces = cfg.bounceEnd - cfg.bounceStart;
            cfg.nChannels = (unsigned char)C;
            cfg.transform = true;
            cfg.maxMoment = 3;
            cfg.cudaGroups.push_back(DenoiseGroup);
        }
        for (int t = 0; t < (nbm > 0 ? 2 : 0); t++) {
            auto &cfg = cfgs[t == 0 ? MISBSDFWinRate : MISLightWinRate];
            cfg.type = t == 0 ? MISBSDFWinRate : MISLightWinRate;
            cfg.index = cfgs.nEnabled++;
            cfg.enable = true;
            cfg.bounceStart = 0;
            cfg.bounceEnd = (unsigned char)nbm;
            cfg.nBounces = (unsigned char)nbm;
            cfg.nChannels = 1;
            cfg.transform = false;
            cfg.maxMoment = 3;
            cfg.cudaGroups.push_back(DenoiseGroup);
        }
        const float sds[2] = {normalSD, albedoSD};
        const unsigned char types[2] = {StatNormal, StatAlbedo};
        for (int g = 0; g < 2; g++) {
            auto &cfg = cfgs[types[g]];
            cfg.enable = true;
            cfg.enableForFilter = true;
            cfg.filterSD = sds[g];
            cfg.type = types[g];
            cfg.index = cfgs.nEnabled++;
            cfg.bounceStart = 0;
            cfg.bounceEnd = 1;
            cfg.nBounces = 1;
            cfg.nChannels = 3;
            cfg.gBuffer = true;
            cfg.transform = false;
            cfg.maxMoment = 1;
        }
        // ---- StatPathIntegrator ctor: film buffer (Film::FilmBuffer = Buffer("film", Mat3(h, w)), film.h:79-90), registry,
        //      estimator, AllocateBuffers ---------------------------------------------------------------------------------
        Buffer filmBuffer("film", Mat3(H, W));
        BufferRegistry reg(filmBuffer);
        const int iNormal = cfgs[StatNormal].index, iAlbedo = cfgs[StatAlbedo].index;
        Estimator est(filmBuffer, cfgs, filterSD, (unsigned char)radius, denoiseFilm != 0, acrr != 0, nbm > 0, 1, reg,
                      Bounds2i(Point2i(0, 0), Point2i(W, H)), nullptr);
        est.AllocateBuffers(reg);
        if (n_registered) *n_registered = (int)reg.buffers.size();
        if (!est.runCUDA) throw std::runtime_error("Estimator::runCUDA is false");

        // ---- the planes the render loop / ReadFile fills -------------------------------------------------------------------
        const size_t px = (size_t)W * H;
        put(est.filmBuffer, film, px * 12);
        for (int j = 0; j < nb; j++) {
            put(est.nBuffers[0][j], n + j * px, px * 4);
            put(est.meanBuffers[0][j], mean + j * px * C, px * C * 4);
            put(est.m2Buffers[0][j], m2 + j * px * C, px * C * 4);
            put(est.m3Buffers[0][j], m3 + j * px * C, px * C * 4);
            put(est.filmBuffers[0][j], film_mean + j * px * C, px * C * 4);
        }
        put(est.filmBuffers[iNormal][0], normal, px * 12);  // features: mean == film-mean (no transform, estimator.cpp:128-136)
        put(est.filmBuffers[iAlbedo][0], albedo, px * 12);
        for (int t = 0; t < (nbm > 0 ? 2 : 0); t++)
            for (int j = 0; j < nbm; j++) {
                const size_t o = ((size_t)t * nbm + j) * px;
                put(est.nBuffers[1 + t][j], mis_n + o, px * 4);
                put(est.meanBuffers[1 + t][j], mis_mean + o, px * 4);  // == film-mean (untransformed type)
                put(est.m2Buffers[1 + t][j], mis_m2 + o, px * 4);
                put(est.m3Buffers[1 + t][j], mis_m3 + o, px * 4);
            }

        // ---- statpath.cpp:406-418 ------------------------------------------------------------------------------------
        double ns = 0;
        for (int r = 0; r < (reps < 1 ? 1 : reps); r++) {
            const auto t0 = std::chrono::steady_clock::now();
            est.Upload();
            est.Denoise();
            est.Download();
            est.Synchronize();
            ns = (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        }
        if (cuda_time_ns) *cuda_time_ns = ns;

        // planes the reference keeps on the device only (not in downloadBuffers): fetch them with Buffer::download
        for (int j = 0; j < nb; j++) {
            est.meanCorrBuffers[0][j].download(est.stream);
            est.discriminatorBuffers[0][j].download(est.stream);
            if (C == 1 && !acrr && nbm == 0) est.filmFilteredBuffers[0][j].download(est.stream);  // only ACRR/SMIS download them (estimator.cpp:236-240)
        }
        est.Synchronize();

        if (film_f) get(est.filmFilteredBuffer.mat, film_f, px * 12);
        for (int j = 0; j < nb; j++) {
            if (film_mean_f) get(est.filmFilteredBuffers[0][j].mat, film_mean_f + j * px * C, px * C * 4);
            if (mean_corr) get(est.meanCorrBuffers[0][j].mat, mean_corr + j * px * C, px * C * 4);
            if (disc) get(est.discriminatorBuffers[0][j].mat, disc + j * px * C, px * C * 4);
        }

        for (int t = 0; t < (nbm > 0 ? 2 : 0); t++)
            for (int j = 0; j < nbm; j++)
                if (mis_f) get(est.filmFilteredBuffers[1 + t][j].mat, mis_f + ((size_t)t * nbm + j) * px, px * 4);

        if (dump_stem) {
            OutputBufferSelection sel(reg, std::regex(dump_regex ? dump_regex : "film.*"), std::string(dump_stem) + ".pfm");
            sel.PrepareOutput();
            sel.Write(dump_suffix ? dump_suffix : "");
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown exception";
        return 2;
    }
}

/* ---- CPU-only checks of the shim's host half (no device needed) ------------------------------------------------------ */

/* Writes `data` ([H][W][C] float32, RGB order) the way OutputBufferSelection::Write does (buffer.cpp:40-53) and reads it
 * back the way StatPathIntegrator::ReadFile does (statpath.cpp:449-454), into `back` (float32) or `back_i32` (a CV_32S
 * plane such as `n`: convertTo rounds half to even). */
extern "C" int smr_shim_pfm_roundtrip(const char *filename, int W, int H, int C, const float *data, float *back,
                                      int32_t *back_i32) {
    try {
        cv::Mat m(H, W, CV_MAKETYPE(CV_32F, C));
        std::memcpy(m.data, data, (size_t)W * H * C * 4);
        if (C == 3) {
            cv::Mat bgr;
            cv::cvtColor(m, bgr, cv::COLOR_RGB2BGR);
            if (!cv::imwrite(filename, bgr)) throw std::runtime_error("imwrite failed");
        } else if (!cv::imwrite(filename, m))
            throw std::runtime_error("imwrite failed");
        cv::Mat dst(H, W, CV_MAKETYPE(back_i32 ? CV_32S : CV_32F, C));
        const uchar *before = dst.data;
        cv::imread(filename, cv::IMREAD_UNCHANGED).convertTo(dst, dst.type());
        if (dst.channels() == 3) cv::cvtColor(dst, dst, cv::COLOR_BGR2RGB);
        if (dst.data != before) throw std::runtime_error("convertTo re-allocated a matching destination");
        std::memcpy(back_i32 ? (void *)back_i32 : (void *)back, dst.data, (size_t)W * H * C * 4);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

/* cv::Mat header semantics the reference relies on: Buffer objects are copied by value and share pixels through the
 * reference count (buffer.cpp:8-10, estimator.cpp:128-144); returns 0 when all of them hold. */
extern "C" int smr_shim_mat_semantics() {
    try {
        cv::Mat a(5, 7, CV_32FC3);
        if (a.rows != 5 || a.cols != 7 || a.channels() != 3 || a.step[0] != 7 * 12 || !a.isContinuous() || !a.u) return 10;
        a.at<cv::Vec3f>(2, 3) = cv::Vec3f(1, 2, 3);
        cv::Mat b = a;  // shares
        if (b.data != a.data || a.u->refcount != 2) return 11;
        {
            cv::Mat c(b);
            cv::Mat d(std::move(c));
            if (a.u->refcount != 3 || c.data || d.data != a.data) return 12;
            cv::Mat e;
            e = d;
            if (a.u->refcount != 4) return 13;
            e = cv::Mat(2, 2, CV_32SC1);  // move-assign drops one reference
            if (a.u->refcount != 3 || e.type() != CV_32SC1) return 14;
        }
        if (a.u->refcount != 2) return 15;
        cv::Mat_<int> n(3, 4);
        for (int i = 0; i < 12; i++) ((int *)n.data)[i] = i - 3;
        cv::Mat f;
        n.convertTo(f, CV_32F);
        if (f.type() != CV_32FC1 || f.at<float>(0, 0) != -3.f || f.at<float>(2, 3) != 8.f) return 16;
        cv::Mat empty(1, 0, CV_8UC(24));  // the 1 x 0 descriptor tables of unused CUDA groups (estimator.cpp:43-46)
        if (empty.data || empty.rows != 1 || empty.cols != 0) return 17;
        std::vector<float> v = {1.f, 2.f, 3.f};
        cv::Mat w(v);  // wraps, no allocation (estimator.cpp:287)
        if (w.rows != 3 || w.cols != 1 || w.data != (uchar *)v.data() || w.u) return 18;
        pbrt::Buffer film("film", a, cv::cuda::GpuMat());  // a Buffer without device memory: names and outMat only
        if (film.channelNames.size() != 3 || film.channelNames[2] != "film.B" || film.outMat.data != a.data) return 19;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
