/* samplelog_hook.h -- forced include (-include) for ONE translation unit of the reference, src/statistics/statpath.cpp,
 * in the sample-logging build of its renderer (oracle/Makefile: _ref/pbrt_ref_cpu_samplelog).  TEST INFRASTRUCTURE ONLY.
 *
 * It declares an explicit specialisation of the one-line forwarder StatTile<Vec3>::AddTransformSampleM3
 * (src/statistics/estimator.h:232) that first hands the radiance sample to oracle/ref_sample_log.cpp and then calls the
 * reference's own AddTransformSample / AddStatSampleM3 (estimator.h:190-226) exactly as the original forwarder does: the
 * arithmetic stays the reference's, statpath.cpp stays unmodified, and the renderer additionally emits the per-pixel
 * sample stream that the accumulation parity tests replay (tests/golden/render_veach_mis_16spp_samples.npz).
 * Only meaningful with trackedbounces = 0 (one radiance tile per image tile). */
#pragma once
#include "statistics/estimator.h"

extern "C" void smr_log_sample(int x, int y, unsigned long long n_before, const float *rgb);

namespace pbrt {
template <>
inline void StatTile<Vec3>::AddTransformSampleM3(const Point2i p, const Vec3 sample) {
    const float rgb[3] = {sample[0], sample[1], sample[2]};
    smr_log_sample(p.x, p.y, GetPixel(p).n, rgb);
    AddTransformSample(p, sample, &StatTile<Vec3>::AddStatSampleM3);
}
}  // namespace pbrt
