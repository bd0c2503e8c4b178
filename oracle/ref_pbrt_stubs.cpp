/*
 * ref_pbrt_stubs.cpp -- what the reference's renderer needs besides its own sources to link WITHOUT its CMake build
 * (oracle/Makefile target `pbrt`: every .cpp of /root/reference/src compiled unmodified from where it lies, except the
 * files named below).  TEST INFRASTRUCTURE ONLY: the resulting binaries (oracle/_ref/pbrt_ref_*) exist to produce REAL
 * statistic dumps from the reference's own CPU path tracer (BASELINE.json configs[0]) and to run its `--denoise` flow
 * on libstatmc_b200.  This file is ours; it replaces
 *   src/core/imageio.cpp            needs OpenEXR (vendored headers, no generated config, no library): PFM only here
 *   src/textures/ptex.cpp           needs the Ptex library: the two creators report "unsupported"
 *   src/display/                    pbrt-v4's display-server client (needs OpenEXR too): no-ops
 *   src/statistics/luts/uberalbedo.cpp   a 64 MiB table the reference's CMakeLists.txt:306-309 downloads: zeros, so the
 *                                   albedo feature of `uber` materials is 0 (no scene used for fixtures has one)
 */
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "core/geometry.h"
#include "core/imageio.h"
#include "core/paramset.h"
#include "core/spectrum.h"
#include "core/texture.h"
#include "core/transform.h"
#include "statistics/luts/uberalbedo.h"

namespace pbrt {

// ---- src/core/imageio.h --------------------------------------------------------------------------------------------
std::unique_ptr<RGBSpectrum[]> ReadImage(const std::string &name, Point2i *) {
    std::fprintf(stderr, "ref_pbrt_stubs: ReadImage(%s) is not available in this build (no OpenEXR / image textures)\n",
                 name.c_str());
    return nullptr;
}

// `rgb` holds outputBounds.Area() RGB triples, top row first (imageio.cpp:80-118).  Only PFM is written (bottom row first,
// little-endian, scale -1); any other extension is skipped with a note -- the statistic planes this build exists for are
// written by OutputBufferSelection::Write (buffer.cpp:40-53), not by this function.
void WriteImage(const std::string &name, const Float *rgb, const Bounds2i &outputBounds, const Point2i &) {
    const Vector2i res = outputBounds.Diagonal();
    const size_t dot = name.find_last_of('.');
    if (dot == std::string::npos || name.substr(dot) != ".pfm") {
        std::fprintf(stderr, "ref_pbrt_stubs: WriteImage(%s) skipped (only .pfm is supported in this build)\n", name.c_str());
        return;
    }
    FILE *f = std::fopen(name.c_str(), "wb");
    if (!f) {
        std::fprintf(stderr, "ref_pbrt_stubs: cannot create %s\n", name.c_str());
        return;
    }
    std::fprintf(f, "PF\n%d %d\n-1.000000\n", res.x, res.y);
    std::vector<float> row((size_t)res.x * 3);
    for (int y = res.y - 1; y >= 0; --y) {
        for (int i = 0; i < res.x * 3; i++) row[i] = (float)rgb[(size_t)y * res.x * 3 + i];
        std::fwrite(row.data(), 4, row.size(), f);
    }
    std::fclose(f);
}

// ---- src/textures/ptex.h ---------------------------------------------------------------------------------------------
Texture<Float> *CreatePtexFloatTexture(const Transform &, const TextureParams &) {
    std::fprintf(stderr, "ref_pbrt_stubs: ptex textures are not available in this build\n");
    return nullptr;
}
Texture<Spectrum> *CreatePtexSpectrumTexture(const Transform &, const TextureParams &) {
    std::fprintf(stderr, "ref_pbrt_stubs: ptex textures are not available in this build\n");
    return nullptr;
}

// ---- src/statistics/luts/uberalbedo.h ----------------------------------------------------------------------------------
Float uberAlbedoLUT[8 * 8 * 8 * 8 * 8 * 8 * 8 * 8];  // zero-initialised (.bss)
unsigned char uberAlbedoLUTNDims = 8;
unsigned char uberAlbedoLUTMaxIndices[8] = {7, 7, 7, 7, 7, 7, 7, 7};
unsigned int uberAlbedoLUTOffsets[256];

}  // namespace pbrt

// ---- src/display/pbrt/util/display.h ----------------------------------------------------------------------------------
namespace pbrtv4 {
void ConnectToDisplayServer(const std::string &) {}
void DisconnectFromDisplayServer() {}
void DisplayStatic(std::string, unsigned short, unsigned short, const float *, std::vector<std::string>) {}
}  // namespace pbrtv4
