// smc_denoise -- denoise-only replay of StatMC statistic dumps on the B200 path: the `pbrt --denoise scene.pbrt` flow
// (StatPathIntegrator::Denoise<T>, src/statistics/statpath.cpp:455-550 of the reference) without the renderer.
//
// A StatMC render with `--writeimages` leaves, per iteration, the files "<stem>-<spp>-<buffer>.pfm"
// (OutputBufferSelection::Write, src/statistics/buffer.cpp:40-53): "film", and per statistic type i / bounce j the planes
// "t<i>-b<j>-{n,mean,m2,m3,film-mean,film-m2,...}".  This program rebuilds the integrator's statistic-type table from the same
// parameter names the scene file uses (statpath.cpp:986-1001, 1027-1160), allocates the planes through statmc::Estimator,
// loads every dump file that names one of them, runs Upload -> Denoise -> Download -> Synchronize on the GPU and writes
// the planes selected by --outputregex back as PFM, with the reference's console lines (Iteration / I/O time / CUDA time /
// Output time).
//
//   smc_denoise --stem out/scene [--width W --height H] [--pixelsamples 4] [--iterations 16] [--expiterations true]
//               [--filtersd 10] [--filterradius 20] [--filterbuffers albedo,normal] [--filterbuffersds 0.02,0.1]
//               [--multichannelstats true] [--denoiseimage true] [--acrr false] [--smis false] [--trackedbounces 5]
//               [--calcprodenstats false] [--outputregex "film.*"] [--outstem <stem for the outputs>] [--device 0]
//               [--warmup] [--pipelined]
//   smc_denoise --pfm-copy in.pfm out.pfm [--as-int]     (codec check: decode + re-encode, no GPU work)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <map>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

#include "statmc_b200.hpp"
#include "statmc_pfm.hpp"

using namespace statmc;
namespace fs = std::filesystem;

namespace {

struct Options {
    std::string stem, outStem;
    int width = 0, height = 0;
    unsigned long long spp = 4;          // Sampler "pixelsamples"
    unsigned int nIterations = 16;       // statpath.cpp:986
    bool expIterations = true;           // :987
    int nTrackedBounces = 5;             // :988 (defaults to maxdepth = 5)
    bool multiChannelStats = true;       // :989
    bool acrr = false, smis = false;     // :991-992
    bool calcProDenStats = false;        // :993
    bool denoiseImage = false;           // :997
    float filterSD = 10.f;               // :1000
    int filterRadius = 20;               // :1001
    std::vector<std::string> filterBuffers;
    std::vector<float> filterBufferSDs;
    std::string outputRegex = "film.*";  // :1175
    int device = 0;
    bool warmUp = false, pipelined = false;
};

[[noreturn]] void die(const std::string &msg) {
    std::cerr << "smc_denoise: " << msg << std::endl;
    std::exit(1);
}

bool parseBool(const std::string &v, const std::string &key) {
    if (v == "true" || v == "1") return true;
    if (v == "false" || v == "0") return false;
    die("--" + key + " expects true or false");
}

std::vector<std::string> splitList(const std::string &v) {
    std::vector<std::string> out;
    std::string item;
    std::stringstream ss(v);
    while (std::getline(ss, item, ','))
        if (!item.empty()) out.push_back(item);
    return out;
}

// statpath.cpp:1013-1173: which statistic types exist, in StatTypeIndex order; the estimator numbers the enabled ones t0, t1, ...
StatTypeConfigs makeStatTypeConfigs(const Options &o) {
    StatTypeConfigs cfgs;
    cfgs.configs.resize(8);
    if (o.acrr || o.calcProDenStats || o.denoiseImage) {
        StatTypeConfig &c = cfgs[Radiance];
        c.type = Radiance;
        c.enable = true;
        c.bounceStart = 0;
        c.bounceEnd = (unsigned char)(o.acrr ? o.nTrackedBounces : 1);
        c.nBounces = c.bounceEnd - c.bounceStart;
        c.nChannels = o.multiChannelStats ? 3 : 1;
        if (o.calcProDenStats) c.maxMoment = 2;
        if (o.acrr || o.denoiseImage) {  // denoising needs the Box-Cox statistics up to the third moment
            c.transform = true;
            c.maxMoment = 3;
            c.cudaGroups.push_back(DenoiseGroup);
        }
        if (o.calcProDenStats) c.cudaGroups.push_back(CalculateMeanVarianceGroup);
    }
    if (o.smis) {
        for (unsigned char t : {(unsigned char)MISBSDFWinRate, (unsigned char)MISLightWinRate}) {
            StatTypeConfig &c = cfgs[t];
            c.type = t;
            c.enable = true;
            c.bounceStart = 0;
            c.bounceEnd = c.nBounces = (unsigned char)o.nTrackedBounces;
            c.nChannels = 1;
            c.transform = false;
            c.maxMoment = 3;
            c.cudaGroups.push_back(DenoiseGroup);
        }
    }
    if (o.acrr || o.denoiseImage || o.smis || o.calcProDenStats) {
        struct G {
            const char *name;
            unsigned char type, channels;
        };
        for (const G &g : {G{"materialid", StatMaterialID, 1}, G{"depth", StatDepth, 1}, G{"normal", StatNormal, 3}, G{"albedo", StatAlbedo, 3}}) {
            StatTypeConfig &c = cfgs[g.type];
            for (size_t k = 0; k < o.filterBuffers.size(); k++)
                if (o.filterBuffers[k] == g.name) {
                    c.enable = true;
                    if (o.acrr || o.denoiseImage || o.smis) {
                        c.enableForFilter = true;
                        c.filterSD = o.filterBufferSDs[k];
                    }
                    break;
                }
            if (!c.enable) continue;
            c.type = g.type;
            c.bounceStart = 0;
            c.bounceEnd = c.nBounces = 1;
            c.nChannels = g.channels;
            c.gBuffer = true;
            c.transform = false;
            c.maxMoment = 1;
            if (o.calcProDenStats) {
                c.maxMoment = 2;
                c.cudaGroups.push_back(CalculateMeanVarianceGroup);
            }
        }
    }
    for (auto &c : cfgs.configs)
        if (c.enable) cfgs.nEnabled++;
    return cfgs;
}

struct Named {
    std::string name;
    Buffer *buffer;
};

long long nsSince(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
}

int pfmCopy(const std::string &in, const std::string &out, bool asInt) {
    const pfm::Header h = pfm::peek(in);
    // borrowed pageable memory: the codec check must not need a GPU (statmc::Mat(rows, cols, ...) allocates pinned memory)
    std::vector<unsigned char> mem((size_t)h.rows * h.cols * h.channels * 4);
    Mat m(h.rows, h.cols, h.channels, asInt ? S32 : F32, mem.data(), 0);
    pfm::read(in, m);
    pfm::write(out, m);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    Options o;
    try {
        if (argc >= 4 && std::string(argv[1]) == "--pfm-copy")
            return pfmCopy(argv[2], argv[3], argc > 4 && std::string(argv[4]) == "--as-int");
        for (int i = 1; i < argc; i++) {
            std::string key = argv[i];
            if (key.rfind("--", 0) != 0) die("unexpected argument " + key);
            key = key.substr(2);
            if (key == "warmup") { o.warmUp = true; continue; }
            if (key == "pipelined") { o.pipelined = true; continue; }
            if (i + 1 >= argc) die("--" + key + " needs a value");
            const std::string v = argv[++i];
            if (key == "stem") o.stem = v;
            else if (key == "outstem") o.outStem = v;
            else if (key == "width") o.width = std::atoi(v.c_str());
            else if (key == "height") o.height = std::atoi(v.c_str());
            else if (key == "pixelsamples") o.spp = std::strtoull(v.c_str(), nullptr, 10);
            else if (key == "iterations") o.nIterations = (unsigned int)std::atoi(v.c_str());
            else if (key == "expiterations") o.expIterations = parseBool(v, key);
            else if (key == "trackedbounces") o.nTrackedBounces = std::atoi(v.c_str());
            else if (key == "multichannelstats") o.multiChannelStats = parseBool(v, key);
            else if (key == "acrr") o.acrr = parseBool(v, key);
            else if (key == "smis") o.smis = parseBool(v, key);
            else if (key == "calcprodenstats") o.calcProDenStats = parseBool(v, key);
            else if (key == "denoiseimage") o.denoiseImage = parseBool(v, key);
            else if (key == "filtersd") o.filterSD = (float)std::atof(v.c_str());
            else if (key == "filterradius") o.filterRadius = std::atoi(v.c_str());
            else if (key == "filterbuffers") o.filterBuffers = splitList(v);
            else if (key == "filterbuffersds") { for (auto &s : splitList(v)) o.filterBufferSDs.push_back((float)std::atof(s.c_str())); }
            else if (key == "outputregex") o.outputRegex = v;
            else if (key == "device") o.device = std::atoi(v.c_str());
            else die("unknown option --" + key);
        }
        if (o.stem.empty()) die("--stem is required (the dump files are <stem>-<spp>-<buffer>.pfm)");
        if (o.outStem.empty()) o.outStem = o.stem;
        if (o.filterBuffers.size() != o.filterBufferSDs.size()) die("Size of filterbuffers and filterbuffersds must match.");  // statpath.cpp:1090-1093
        if (o.filterRadius < 0 || o.filterRadius > 255) die("filterradius must be 0..255 (unsigned char in the reference)");
        if (o.nTrackedBounces < 0 || o.nTrackedBounces > 255) die("trackedbounces must be 0..255");

        auto sppOf = [&](unsigned int i) { return o.expIterations ? o.spp << (i - 1) : (unsigned long long)i * o.spp; };
        // image size: the reference takes it from the scene's Film; here from the options or the first dump file found
        if (o.width <= 0 || o.height <= 0) {
            for (unsigned int i = 1; i <= o.nIterations && o.width <= 0; i++) {
                const std::string prefix = o.stem + "-" + std::to_string(sppOf(i)) + "-";
                for (const char *probe : {"film.pfm", "t0-b0-n.pfm", "t0-b0-mean.pfm"})
                    if (fs::exists(prefix + probe)) {
                        const pfm::Header h = pfm::peek(prefix + probe);
                        o.width = h.cols;
                        o.height = h.rows;
                        break;
                    }
            }
            if (o.width <= 0) die("no dump file found for stem " + o.stem + " (give --width/--height or check --pixelsamples/--iterations)");
        }
        if (o.width > SMC_MAX_DIM || o.height > SMC_MAX_DIM) die("image larger than 65535 (unsigned short in the reference)");

        Stream stream(o.device);
        Buffer filmBuffer(stream, "film", Mat(o.height, o.width, 3));
        Estimator estimator(stream, filmBuffer, makeStatTypeConfigs(o), o.filterSD, (unsigned char)o.filterRadius, o.denoiseImage, o.acrr,
                            o.smis);
        estimator.AllocateBuffers();

        // BufferRegistry (estimator.cpp:20-34): every plane by name, in allocation order, plus film / film-f
        std::vector<Named> registry;
        registry.push_back({"film", &estimator.filmBuffer});
        registry.push_back({"film-f", &estimator.filmFilteredBuffer});
        for (auto *v : {&estimator.nBuffers, &estimator.meanBuffers, &estimator.m2Buffers, &estimator.m3Buffers, &estimator.filmBuffers,
                        &estimator.filmM2Buffers, &estimator.meanCorrBuffers, &estimator.discriminatorBuffers,
                        &estimator.filmVarBuffers, &estimator.filmFilteredBuffers})
            for (auto &perType : *v)
                for (Buffer &b : perType) registry.push_back({b.name, &b});
        const std::regex outRe(o.outputRegex);
        std::vector<Named> outputs;
        for (const Named &n : registry)
            if (std::regex_match(n.name, outRe)) outputs.push_back(n);

        // the suffixes --denoise reads (statpath.cpp:504-511) -> plane vectors
        const std::map<std::string, std::vector<std::vector<Buffer>> *> readable = {
            {"n", &estimator.nBuffers},
            {"mean", &estimator.meanBuffers},
            {"m2", &estimator.m2Buffers},
            {"m3", &estimator.m3Buffers},
            {"film-m2", &estimator.filmM2Buffers},
            {"mean-corr", &estimator.meanCorrBuffers},
            {"discriminator", &estimator.discriminatorBuffers},
            {"film-mean", &estimator.filmBuffers}};
        const std::regex idRe("^t([0-9]+)-b([0-9]+)-(.*)$");

        auto denoiseLoop = [&](unsigned int nIterations) {
            for (unsigned int i = 1; i <= nIterations; i++) {
                auto begin = std::chrono::steady_clock::now();
                const unsigned long long currentSPP = sppOf(i);
                const std::string prefix = o.stem + "-" + std::to_string(currentSPP) + "-";
                int nRead = 0;
                if (fs::exists(prefix + "film.pfm")) {
                    pfm::read(prefix + "film.pfm", estimator.filmBuffer.mat);
                    nRead++;
                }
                // cv::glob(prefix + "*.pfm") -- sorted, non-recursive
                std::vector<std::string> files;
                const fs::path dir = fs::path(prefix).parent_path().empty() ? fs::path(".") : fs::path(prefix).parent_path();
                const std::string base = fs::path(prefix).filename().string();
                if (fs::is_directory(dir))
                    for (const auto &e : fs::directory_iterator(dir)) {
                        const std::string fn = e.path().filename().string();
                        if (fn.size() > base.size() + 4 && fn.compare(0, base.size(), base) == 0 && fn.compare(fn.size() - 4, 4, ".pfm") == 0)
                            files.push_back(fn);
                    }
                std::sort(files.begin(), files.end());
                for (const std::string &fn : files) {
                    const std::string id = fn.substr(base.size(), fn.size() - 4 - base.size());
                    std::smatch m;
                    if (!std::regex_match(id, m, idRe)) continue;
                    const size_t typeIndex = (size_t)std::stoul(m[1]), bounceIndex = (size_t)std::stoul(m[2]);
                    auto it = readable.find(m[3]);
                    if (it == readable.end()) continue;
                    auto &vec = *it->second;
                    if (typeIndex >= vec.size() || bounceIndex >= vec[typeIndex].size()) {  // the reference indexes unchecked
                        std::cerr << "smc_denoise: skipping " << fn << " (no statistic type " << typeIndex << " / bounce " << bounceIndex
                                  << " in this configuration)" << std::endl;
                        continue;
                    }
                    Buffer &b = vec[typeIndex][bounceIndex];
                    pfm::read((dir / fn).string(), b.mat);
                    nRead++;
                }
                if (nRead == 0) std::cerr << "smc_denoise: no input planes found for " << prefix << "*.pfm" << std::endl;
                std::cout << "Iteration: " << i << std::endl;
                std::cout << "I/O time [ns]: " << nsSince(begin) << std::endl;

                begin = std::chrono::steady_clock::now();
                // statpath.cpp:529-532.  CalculateMeanVars is an addition (the reference computes the estimator variance in its
                // render loop only, estimator.cpp:491-569); it must run before Download fetches its result plane.
                if (o.pipelined && estimator.runCUDA && !o.calcProDenStats) {
                    estimator.DenoiseHost();
                } else {
                    estimator.Upload();
                    if (estimator.runCUDA) estimator.Denoise();
                    if (o.calcProDenStats) estimator.CalculateMeanVars();
                    estimator.Download();
                }
                // The reference never downloads mean-corr / discriminator (estimator.cpp:163-178), so selecting them for output
                // there writes the untouched host planes; here they are fetched when asked for.
                for (const Named &n : outputs) {
                    const auto ends = [&](const char *suf) {
                        const size_t l = std::strlen(suf);
                        return n.name.size() >= l && n.name.compare(n.name.size() - l, l, suf) == 0;
                    };
                    if ((ends("-mean-corr") || ends("-discriminator")) && !estimator.downloadBuffers.count(n.buffer)) n.buffer->download(stream);
                }
                estimator.Synchronize();
                std::cout << "CUDA time [ns]: " << nsSince(begin) << std::endl;

                begin = std::chrono::steady_clock::now();
                for (const Named &n : outputs) pfm::write(o.outStem + "-" + std::to_string(currentSPP) + "-" + n.name + ".pfm", n.buffer->mat);
                std::cout << "Output time [ns]: " << nsSince(begin) << std::endl;
            }
        };
        if (o.warmUp) {
            std::cout << "==== Warm-Up Start ====" << std::endl;
            denoiseLoop(1);
            std::cout << "==== Warm-Up End ====" << std::endl;
        }
        denoiseLoop(o.nIterations);
    } catch (const Exception &e) {
        std::cerr << "smc_denoise: error " << e.code << ": " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
