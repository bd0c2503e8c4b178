// test_estimator.cpp -- exercises the C++ host layer (include/statmc_b200.hpp) the way StatPathIntegrator uses
// src/statistics/: build an Estimator for RGB radiance + normal + albedo, feed sample batches, Upload / Denoise / Download /
// Synchronize, then replay the dumped statistics through a second Estimator (the `--denoise` flow, statpath.cpp:456-550)
// and through the kernel-level stat_denoiser::filter<float3> call with device pointer tables (samples/stat_denoiser/main.cpp).
// Reads its inputs from a file written by tests/test_cpp_host_gpu.py and writes every result plane back for comparison with
// the oracle there.   usage: test_estimator <in.bin> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "statmc_b200.hpp"

using namespace statmc;

struct Header {
    int W, H, S, radius;
    float sd, normal_sd, albedo_sd;
};

static std::vector<float> read_floats(FILE *f, size_t n) {
    std::vector<float> v(n);
    if (fread(v.data(), 4, n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(2);
    }
    return v;
}

static void write_mat(FILE *f, const Mat &m) {
    for (int y = 0; y < m.rows; y++) fwrite(m.ptr<unsigned char>(y), 1, (size_t)m.cols * m.channels() * 4, f);
}

static StatTypeConfigs make_configs(const Header &h) {
    StatTypeConfigs cfgs;
    StatTypeConfig rad;  // statpath.cpp:1027-1054
    rad.type = Radiance; rad.enable = true; rad.nBounces = 1; rad.bounceStart = 0; rad.bounceEnd = 1; rad.nChannels = 3;
    rad.transform = true; rad.maxMoment = 3; rad.cudaGroups = {DenoiseGroup, CalculateMeanVarianceGroup};
    StatTypeConfig nrm;  // statpath.cpp:1128-1160: features are M1, untransformed, registered as G-buffers
    nrm.type = StatNormal; nrm.enable = true; nrm.nBounces = 1; nrm.bounceStart = 0; nrm.bounceEnd = 1; nrm.nChannels = 3;
    nrm.transform = false; nrm.maxMoment = 1; nrm.gBuffer = true; nrm.enableForFilter = true; nrm.filterSD = h.normal_sd;
    StatTypeConfig alb = nrm;
    alb.type = StatAlbedo; alb.filterSD = h.albedo_sd;
    StatTypeConfig off;  // a disabled type must be dropped (estimator.h:268-269)
    off.type = StatDepth; off.enable = false;
    cfgs.configs = {rad, off, nrm, alb};
    return cfgs;
}

int main(int argc, char **argv) {
    if (argc != 3) return 2;
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 2;
    Header h;
    if (fread(&h, sizeof(h), 1, fi) != 1) return 2;
    const size_t px = (size_t)h.W * h.H;
    std::vector<float> rad = read_floats(fi, px * 3 * h.S), nrm = read_floats(fi, px * 3 * h.S),
                       alb = read_floats(fi, px * 3 * h.S), film = read_floats(fi, px * 3);
    fclose(fi);
    FILE *fo = fopen(argv[2], "wb");
    int failures = 0;
    try {
        Stream stream(0);
        // ---- flow A: accumulate on the device, denoise, download --------------------------------------------------------
        Buffer filmBuffer(stream, "film", Mat(h.H, h.W, 3));
        std::memcpy(filmBuffer.mat.ptr(), film.data(), px * 12);
        Estimator est(stream, filmBuffer, make_configs(h), h.sd, (unsigned char)h.radius, true);
        est.AllocateBuffers();
        if (est.statTypeConfigs.nEnabled != 3 || est.gBuffers.size() != 2 || !est.runCUDA || est.rgbBufferCounts[DenoiseGroup] != 1 ||
            est.nBuffers[0][0].name != "t0-b0-n" || est.filmFilteredBuffers[0][0].mat.ptr() != est.filmFilteredBuffer.mat.ptr() ||
            est.meanBuffers[1][0].mat.ptr() != est.filmBuffers[1][0].mat.ptr()) {
            fprintf(stderr, "AllocateBuffers: unexpected layout\n");
            failures++;
        }
        const int s0 = h.S / 2;  // two batches: the streaming update continues across them (statpath.cpp:172-190)
        const size_t off = px * 3 * s0;
        est.AddSamples(0, 0, rad.data(), s0);
        est.AddSamples(1, 0, nrm.data(), s0);
        est.AddSamples(2, 0, alb.data(), s0);
        est.AddSamples(0, 0, rad.data() + off, h.S - s0);
        est.AddSamples(1, 0, nrm.data() + off, h.S - s0);
        est.AddSamples(2, 0, alb.data() + off, h.S - s0);
        est.Upload();
        est.Denoise();
        est.CalculateMeanVars();
        est.Download();
        // the statistics live on the device: fetch them for the dump (the reference's --writeimages, buffer.cpp:40-53)
        for (auto *v : {&est.nBuffers, &est.meanBuffers, &est.m2Buffers, &est.m3Buffers, &est.filmBuffers, &est.filmM2Buffers,
                        &est.meanCorrBuffers, &est.discriminatorBuffers})
            for (auto &per_type : *v)
                for (Buffer &b : per_type) b.download(stream);
        est.Synchronize();
        write_mat(fo, est.filmFilteredBuffer.mat);
        write_mat(fo, est.nBuffers[0][0].mat);
        write_mat(fo, est.meanBuffers[0][0].mat);
        write_mat(fo, est.m2Buffers[0][0].mat);
        write_mat(fo, est.m3Buffers[0][0].mat);
        write_mat(fo, est.filmBuffers[0][0].mat);
        write_mat(fo, est.filmM2Buffers[0][0].mat);
        write_mat(fo, est.meanCorrBuffers[0][0].mat);
        write_mat(fo, est.discriminatorBuffers[0][0].mat);
        write_mat(fo, est.filmVarBuffers[0][0].mat);
        write_mat(fo, est.filmBuffers[1][0].mat);  // normal feature mean
        write_mat(fo, est.filmBuffers[2][0].mat);  // albedo feature mean

        // ---- flow B: replay the dumped statistics from host planes through a second estimator, pipelined ------------------
        Buffer filmBuffer2(stream, "film", Mat(h.H, h.W, 3));
        std::memcpy(filmBuffer2.mat.ptr(), film.data(), px * 12);
        Estimator rep(stream, filmBuffer2, make_configs(h), h.sd, (unsigned char)h.radius, true);
        rep.AllocateBuffers();
        auto copy = [&](Buffer &dst, Buffer &src) { std::memcpy(dst.mat.ptr(), src.mat.ptr(), src.mat.bytes()); };
        copy(rep.nBuffers[0][0], est.nBuffers[0][0]);
        copy(rep.meanBuffers[0][0], est.meanBuffers[0][0]);
        copy(rep.m2Buffers[0][0], est.m2Buffers[0][0]);
        copy(rep.m3Buffers[0][0], est.m3Buffers[0][0]);
        copy(rep.filmBuffers[1][0], est.filmBuffers[1][0]);
        copy(rep.filmBuffers[2][0], est.filmBuffers[2][0]);
        rep.DenoiseHost();
        rep.Synchronize();
        write_mat(fo, rep.filmFilteredBuffer.mat);

        // ---- flow C: kernel-level API with device-resident PtrStepSzb tables (samples/stat_denoiser/main.cpp:19-44, 151-180) --
        auto table = [&](std::vector<PtrStepSzb> v) {
            GpuMat t(stream, 1, (int)(v.size() * sizeof(PtrStepSzb) / 4), 1, S32);
            t.upload(v.data(), 0);
            stream.waitForCompletion();  // `v` is a temporary
            return t;
        };
        GpuMat out3(stream, h.H, h.W, 3), dummy(stream, h.H, h.W, 3), mc(stream, h.H, h.W, 3), dc(stream, h.H, h.W, 3);
        GpuMat tn = table({rep.nBuffers[0][0].gpuMat}), tmean = table({rep.meanBuffers[0][0].gpuMat}),
               tm2 = table({rep.m2Buffers[0][0].gpuMat}), tm3 = table({rep.m3Buffers[0][0].gpuMat}),
               tfilm = table({rep.filmBuffers[0][0].gpuMat}), tg = table({rep.gBuffers[0].gpuMat, rep.gBuffers[1].gpuMat}),
               tmc = table({mc}), tdc = table({dc}), tout = table({dummy});
        const unsigned char chc[4] = {3, 3, 0, 0};
        GpuMat gch(stream, 1, 1, 1, S32), gdr(stream, 1, 2, 1, F32);
        gch.upload(chc, 0);
        gdr.upload(rep.gBufferDRFactors.data(), 0);
        stat_denoiser::filter<float3>(1, (unsigned short)h.W, (unsigned short)h.H, rep.filterDSFactor, rep.filterRadius, true, tn, tmean,
                                      tm2, tm3, tfilm, rep.filmBuffer.gpuMat, tg, gch, gdr, 2, tmc, tdc, tout, out3, stream);
        Mat out3h(h.H, h.W, 3);
        out3.download(out3h, stream);
        stat_denoiser::synchronize(stream);
        write_mat(fo, out3h);

        // ---- error behaviour: invalid arguments throw (the reference: cv::error -> cv::Exception) --------------------------------
        bool threw = false;
        try {
            stat_denoiser::filter<float3>(0, (unsigned short)h.W, (unsigned short)h.H, -0.005f, 5, true, tn, tmean, tm2, tm3, tfilm,
                                          rep.filmBuffer.gpuMat, tg, gch, gdr, 2, tmc, tdc, tout, out3, stream);
        } catch (const Exception &e) {
            threw = e.code == SMC_ERR_INVALID;
        }
        if (!threw) {
            fprintf(stderr, "ptrCount == 0 did not throw\n");
            failures++;
        }
    } catch (const Exception &e) {
        fprintf(stderr, "statmc::Exception %d: %s\n", e.code, e.what());
        failures++;
    }
    fclose(fo);
    printf("test_estimator: %s\n", failures ? "FAILED" : "ok");
    return failures ? 1 : 0;
}
