// statmc_pfm.hpp -- the PFM dump format StatMC writes with `--writeimages` and reads back with `--denoise`, for the
// B200 host layer (statmc_b200.hpp).  Header-only, no OpenCV.
//
// What the reference does (file:line relative to the reference checkout):
//   write  OutputBufferSelection::Write     src/statistics/buffer.cpp:40-53      "<stem>-<suffix>-<buffer name>.pfm", one
//          file per registered buffer; non-float planes (`n`, CV_32S) are converted to float first (:34-38)
//   read   StatPathIntegrator::ReadFile     src/statistics/statpath.cpp:448-453  imread(IMREAD_UNCHANGED).convertTo(type of
//          the destination plane): float -> int32 rounds to nearest even (cvRound)
//   codec  cv::PFMEncoder / cv::PFMDecoder  src/ext/opencv/modules/imgcodecs/src/grfmt_pfm.cpp:77-258
// The RGB<->BGR swaps of the reference cancel (buffer.cpp:47 + grfmt_pfm.cpp:243-249 on write, :147-149 + statpath.cpp:450-451
// on read), so a file holds plain RGB triples.  Format: "PF\n" (3 channels) or "Pf\n" (1 channel), "<cols> <rows>\n",
// "<scale>\n" (negative = little-endian; the encoder writes "-1"), then rows*cols*channels float32, BOTTOM row first.
// The decoder divides by |scale| (:152-153) and byte-swaps when the sign says big-endian (:138-144).
#ifndef STATMC_PFM_HPP
#define STATMC_PFM_HPP

#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "statmc_b200.hpp"

namespace statmc {
namespace pfm {

struct Header {
    int cols = 0, rows = 0, channels = 0;
    double scale = -1.0;
    long dataOffset = 0;
};

namespace detail {
inline bool readToken(FILE *f, std::string &tok) {  // read_number(): bytes up to the next whitespace (grfmt_pfm.cpp:47-65)
    tok.clear();
    for (int i = 0; i < 2048; i++) {
        const int c = fgetc(f);
        if (c == EOF) return !tok.empty();
        if (std::isspace((unsigned char)c)) return true;
        tok.push_back((char)c);
    }
    return true;
}
inline uint32_t bswap(uint32_t u) { return (u << 24) | ((u & 0xff00u) << 8) | ((u >> 8) & 0xff00u) | (u >> 24); }
}  // namespace detail

// Parses the header; throws statmc::Exception(SMC_ERR_INVALID) on anything the OpenCV decoder rejects.
inline Header readHeader(FILE *f, const std::string &filename) {
    Header h;
    const int p = fgetc(f), t = fgetc(f), nl = fgetc(f);
    if (p != 'P' || (t != 'f' && t != 'F')) throw Exception(SMC_ERR_INVALID, filename + ": not a PFM file (expected Pf / PF)");
    if (nl != '\n') throw Exception(SMC_ERR_INVALID, filename + ": unexpected PFM header (expected a line break after the type)");
    h.channels = t == 'F' ? 3 : 1;
    std::string a, b, c;
    if (!detail::readToken(f, a) || !detail::readToken(f, b) || !detail::readToken(f, c))
        throw Exception(SMC_ERR_INVALID, filename + ": truncated PFM header");
    h.cols = std::atoi(a.c_str());
    h.rows = std::atoi(b.c_str());
    h.scale = std::atof(c.c_str());
    if (h.cols <= 0 || h.rows <= 0) throw Exception(SMC_ERR_INVALID, filename + ": bad PFM dimensions");
    if (!(std::fabs(h.scale) > 0.0)) throw Exception(SMC_ERR_INVALID, filename + ": PFM scale factor must be non-zero");
    h.dataOffset = ftell(f);
    return h;
}

inline Header peek(const std::string &filename) {
    FILE *f = fopen(filename.c_str(), "rb");
    if (!f) throw Exception(SMC_ERR_INVALID, filename + ": cannot open");
    Header h;
    try {
        h = readHeader(f, filename);
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
    return h;
}

// StatPathIntegrator::ReadFile: decode `filename` into the EXISTING plane `mat` (its size, channel count and depth stay;
// a mismatch is an error here -- the reference would silently re-allocate the host matrix and then fail in the upload).
inline void read(const std::string &filename, Mat &mat) {
    FILE *f = fopen(filename.c_str(), "rb");
    if (!f) throw Exception(SMC_ERR_INVALID, filename + ": cannot open");
    try {
        const Header h = readHeader(f, filename);
        if (h.cols != mat.cols || h.rows != mat.rows || h.channels != mat.channels())
            throw Exception(SMC_ERR_INVALID, filename + ": is " + std::to_string(h.cols) + "x" + std::to_string(h.rows) + "x" +
                                                 std::to_string(h.channels) + ", the plane is " + std::to_string(mat.cols) + "x" +
                                                 std::to_string(mat.rows) + "x" + std::to_string(mat.channels()));
        const bool swap = h.scale >= 0.0;  // little-endian host: positive scale = big-endian data (grfmt_pfm.cpp:17-28)
        const float inv = 1.f / (float)std::fabs(h.scale);
        const bool scaled = inv != 1.f;
        const size_t n = (size_t)h.cols * h.channels;
        std::vector<float> row(n);
        for (int y = h.rows - 1; y >= 0; --y) {  // bottom row first
            if (fread(row.data(), 4, n, f) != n) throw Exception(SMC_ERR_INVALID, filename + ": truncated PFM data");
            if (swap) {
                uint32_t *u = reinterpret_cast<uint32_t *>(row.data());
                for (size_t i = 0; i < n; i++) u[i] = detail::bswap(u[i]);
            }
            if (scaled)
                for (size_t i = 0; i < n; i++) row[i] *= inv;
            if (mat.depth() == F32) {
                std::memcpy(mat.ptr<float>(y), row.data(), n * 4);
            } else {  // convertTo(CV_32S): saturate_cast<int>(float) = cvRound, round half to even
                int32_t *d = mat.ptr<int32_t>(y);
                for (size_t i = 0; i < n; i++) {
                    const float v = row[i];
                    d[i] = v >= 2147483648.f ? INT32_MAX : v <= -2147483648.f ? INT32_MIN : (int32_t)std::lrintf(v);
                }
            }
        }
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
}

// OutputBufferSelection::PrepareOutput + Write for one plane (int32 planes are converted to float like buffer.cpp:34-38).
inline void write(const std::string &filename, const Mat &mat) {
    if (mat.channels() != 1 && mat.channels() != 3) throw Exception(SMC_ERR_INVALID, filename + ": PFM needs 1 or 3 channels");
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) throw Exception(SMC_ERR_INVALID, filename + ": cannot create");
    fprintf(f, "P%c\n%d %d\n-1\n", mat.channels() == 3 ? 'F' : 'f', mat.cols, mat.rows);
    const size_t n = (size_t)mat.cols * mat.channels();
    std::vector<float> row(n);
    bool ok = true;
    for (int y = mat.rows - 1; y >= 0 && ok; --y) {
        if (mat.depth() == F32) {
            ok = fwrite(mat.ptr<float>(y), 4, n, f) == n;
        } else {
            const int32_t *s = mat.ptr<int32_t>(y);
            for (size_t i = 0; i < n; i++) row[i] = (float)s[i];
            ok = fwrite(row.data(), 4, n, f) == n;
        }
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw Exception(SMC_ERR_INVALID, filename + ": short write");
}

}  // namespace pfm
}  // namespace statmc
#endif  // STATMC_PFM_HPP
