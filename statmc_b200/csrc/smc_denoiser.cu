// smc_denoiser.cu -- the denoise plan: validation, device descriptor tables, record storage, spatial table,
// kernel selection and launch order.  Host-side replacement for cv::cuda::stat_denoiser::filter<T>
// (stat_denoiser.cu:397-475) and for what Estimator::AllocateBuffers does with GpuMat pointer tables
// (estimator.cpp:35-84, 271-288).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "smc_internal.h"

struct smc_denoiser {
    smc_context *ctx = nullptr;
    int C = 3, ptr_count = 1, W = 0, H = 0, radius = 0, denoise_film = 0, mode = 0, n_gbufs = 0, NG = 0;
    int NGX = 0, gext_stride = 0;  // flattened G-buffer channels beyond the record's seven, and their side array
    float *d_gext = nullptr;
    int row_begin = 0, row_end = 0;
    int skip_top = 0, skip_bottom = 0;
    int kernel_pref = 0;
    float ds_factor = 0.f;
    // device memory owned by the plan
    SmcPtrStepSz *d_tables = nullptr;  // [n | mean | m2 | m3 | film_ptrs | mean_corr | disc | out | accepted][ptr_count] + gbufs
    unsigned char *d_gch = nullptr;
    float *d_gf = nullptr;
    float *d_sw = nullptr;
    int2 *d_rowrange = nullptr;
    unsigned char *d_rec = nullptr;
    bool tables_external = false;  // smc_filter_device_tables: tables live in caller memory
    // resolved pointers
    const SmcPtrStepSz *t_n = nullptr, *t_mean = nullptr, *t_m2 = nullptr, *t_m3 = nullptr, *t_film = nullptr,
                       *t_gbufs = nullptr, *t_mc = nullptr, *t_disc = nullptr, *t_out = nullptr, *t_acc = nullptr;
    SmcPtrStepSz film{}, film_filtered{};
    int padX = 0, rec_pitch = 0, rec_rows = 0, sw_stride = 0;
    size_t rec_image_stride = 0;
    bool use_stream = false, use_sym = false;
    int py = 4;
    // symmetric kernel: forward spatial table, scratch (partial mirror sums), forward-sum plane
    float2 *d_sym_sw = nullptr;
    int2 *d_sym_rowrange = nullptr;
    float sym_sw_special = 0.f;
    // three scalar images per record (see alloc_records): the records, the prepass grid and the filter run over `rec_images`
    // = ceil(ptr_count / 3) record images
    int tri = 0, rec_images = 1;
    float2 *d_sym_scratch2 = nullptr, *d_sym_fwd2 = nullptr;
    float4 *d_sym_scratch = nullptr, *d_sym_fwd = nullptr;
    int *d_sym_scratch_cnt = nullptr, *d_sym_fwd_cnt = nullptr;
    size_t sym_scratch_elems = 0, sym_fwd_elems = 0;
    char kernel_name[64] = "generic";
    // host-pipelined run (smc_denoiser_run_host): host copy of the descriptor tables, copy streams, event pool
    std::vector<SmcPtrStepSz> h_tables;
    std::vector<unsigned char> h_gch;
    std::vector<float> h_gf;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> events;
    int *d_tile_counter = nullptr;
    int *d_sync_counters = nullptr;  // [0] prepass, [1] filter: blocks retired (in-kernel halo protocol)
    // peer halos (multi-GPU, one process per GPU or several plans in one process): flags live behind the records in the
    // same allocation so that one IPC handle covers both.  flags: [0] ready_from_up [1] ready_from_down [2] free_from_up
    // [3] free_from_down, each holding the step number the neighbour has reached.
    int *d_flags = nullptr;
    size_t flags_offset = 0;
    struct Peer {
        unsigned char *rec = nullptr;  // peer's record array (mapped)
        int *flags = nullptr;
        size_t image_stride = 0;
        int H = 0;
        void *ipc_base = nullptr;      // non-null when opened with cudaIpcOpenMemHandle
    } peer[2];                         // 0 = rank above, 1 = rank below
    int step = 0;                      // prepass count (the protocol's clock)
    unsigned long long *d_trace = nullptr;  // SMC_STREAM_TRACE=<file>: per-CTA timeline of the last streaming launch
};

static int taps_in_window(int r) {
    int c = 0;
    for (int dy = -r; dy < r; dy++)
        for (int dx = -r; dx < r; dx++)
            if (dy * dy + dx * dx <= r * r) c++;
    return c;
}

// Spatial table: sw[dy][dx] = dS2 * dSFactor * log2(e) for taps of the window [-r, r) x [-r, r) with
// dS2 = dy^2 + dx^2 <= r^2 (SET_OUTER / SET_INNER, stat_denoiser.cu:24-37), -inf elsewhere (including the margins).
static int build_spatial_table(smc_denoiser *d) {
    const int r = d->radius, MY = SMC_SW_MARGIN_Y, MX = SMC_SW_MARGIN_X;
    const int rows = 2 * r + 2 * MY;
    d->sw_stride = 2 * r + 2 * MX;
    std::vector<float> sw((size_t)rows * d->sw_stride, -INFINITY);
    std::vector<int2> rr(rows);
    for (int tr = 0; tr < rows; tr++) {
        const int dy = tr - r - MY;
        rr[tr] = make_int2(1 << 20, -(1 << 20));
        if (dy < -r || dy > r - 1) continue;
        int lo = 1 << 20, hi = -(1 << 20);
        for (int dx = -r; dx <= r - 1; dx++) {
            const int dS2 = dy * dy + dx * dx;
            if (dS2 > r * r) continue;
            // float(dS2) * dSFactor as the reference forms it, then the change of base in double, rounded once
            sw[(size_t)tr * d->sw_stride + (dx + r + MX)] = (float)((double)dS2 * (double)d->ds_factor * 1.4426950408889634);
            lo = std::min(lo, dx);
            hi = std::max(hi, dx);
        }
        if (lo <= hi) rr[tr] = make_int2(lo, hi + 1);  // +1: the thread's second column sees dx = j - 1
    }
    SMC_CUDA(cudaMalloc(&d->d_sw, sw.size() * sizeof(float)));
    SMC_CUDA(cudaMalloc(&d->d_rowrange, rr.size() * sizeof(int2)));
    SMC_CUDA(cudaMemcpyAsync(d->d_sw, sw.data(), sw.size() * sizeof(float), cudaMemcpyHostToDevice, d->ctx->stream));
    SMC_CUDA(cudaMemcpyAsync(d->d_rowrange, rr.data(), rr.size() * sizeof(int2), cudaMemcpyHostToDevice,
                             d->ctx->stream));
    SMC_CUDA(cudaStreamSynchronize(d->ctx->stream));  // sources are stack/pageable
    return SMC_OK;
}

// Forward spatial table of the symmetric kernel: rows dy = -MY .. r + MY, columns dx = -r - MX .. r + MX (+1 pad);
// finite for the forward offsets that are taps of BOTH positions' windows: dy == 0 and dx in [1, r-1], or dy in [1, r-1] and
// dS2 <= r^2.  (0, r) and (r, 0) -- taps of the lower / right position only -- are handled apart with sym_sw_special.
static int build_sym_table(smc_denoiser *d) {
    SmcFilterParams p;
    std::memset(&p, 0, sizeof(p));
    p.radius = d->radius; p.W = d->W; p.row_begin = d->row_begin; p.row_end = d->row_end; p.ptr_count = d->rec_images;
    p.padX = d->padX; p.rec_pitch = d->rec_pitch; p.C = d->C; p.sm_count = d->ctx->sm_count;
    p.tri = d->tri; p.images = d->ptr_count;
    SmcSymParams g;
    size_t smem = 0;
    if (!smc_filter_sym_geometry(p, g, smem)) return SMC_OK;  // no symmetric variant for this plan
    const int r = d->radius;
    // entries are (-sw, 0): the kernel starts the sum of squared G-buffer differences from them with one packed FMA
    std::vector<float2> sw((size_t)g.sw_rows * g.sw_stride, make_float2(INFINITY, 0.f));
    std::vector<int2> rr(g.sw_rows);
    auto val = [&](int dS2) { return (float)((double)dS2 * (double)d->ds_factor * 1.4426950408889634); };
    for (int tr = 0; tr < g.sw_rows; tr++) {
        const int dy = tr - g.sw_my;
        rr[tr] = make_int2(1 << 20, -(1 << 20));
        if (dy < 0 || dy > r - 1) continue;
        int lo = 1 << 20, hi = -(1 << 20);
        for (int dx = (dy == 0 ? 1 : -r); dx <= r - 1; dx++) {
            const int dS2 = dy * dy + dx * dx;
            if (dS2 > r * r) continue;
            sw[(size_t)tr * g.sw_stride + (dx + r + g.sw_mx)] = make_float2(-val(dS2), 0.f);
            lo = std::min(lo, dx);
            hi = std::max(hi, dx);
        }
        if (lo <= hi) rr[tr] = make_int2(lo, hi + 1);  // +1: the lane's second column sees dx = j - 1
    }
    d->sym_sw_special = val(r * r);
    SMC_CUDA(cudaMalloc(&d->d_sym_sw, sw.size() * sizeof(float2)));
    SMC_CUDA(cudaMalloc(&d->d_sym_rowrange, rr.size() * sizeof(int2)));
    SMC_CUDA(cudaMemcpyAsync(d->d_sym_sw, sw.data(), sw.size() * sizeof(float2), cudaMemcpyHostToDevice, d->ctx->stream));
    SMC_CUDA(cudaMemcpyAsync(d->d_sym_rowrange, rr.data(), rr.size() * sizeof(int2), cudaMemcpyHostToDevice, d->ctx->stream));
    SMC_CUDA(cudaStreamSynchronize(d->ctx->stream));  // sources are stack/pageable
    return SMC_OK;
}

static void fill_filter_params(const smc_denoiser *d, SmcFilterParams &p);

static int alloc_records(smc_denoiser *d) {
    const int r = d->radius;
    // >= 2r + 2 columns of replicated border: the symmetric kernel's virtual centres sit up to r columns outside the image and
    // read r columns beyond themselves
    d->padX = ((2 * r + 2 + 15) / 16) * 16;
    d->rec_pitch = ((d->W + 2 * d->padX + 15) / 16) * 16;
    d->rec_rows = d->H + 2 * r + 4;  // +4: rows only ever paired with out-of-range centre rows of the last tile
    d->rec_image_stride = (size_t)d->rec_rows * smc_rec_row_bytes(d->rec_pitch);
    // Several scalar images over the same G-buffers (Estimator with ptrCount > 1: the ACRR radiance / albedo / ... statistics,
    // one grid.z slice each in the reference, stat_denoiser.cu:422): three of them share one record -- image 3z + k in the slots
    // of channel k of an RGB record -- and the symmetric kernel evaluates the G-buffer weight of a pair once for the three.
    // Only where the symmetric kernel runs (Welch test, no RGB film, no acceptance counts, no external record halos);
    // SMC_SYM_TRIPLE=0 keeps one record image per image.
    d->tri = 0;
    d->rec_images = d->ptr_count;
    {
        int pref = d->kernel_pref;
        if (pref == 0 && getenv("SMC_FILTER_KERNEL")) pref = -1;
        const char *e = getenv("SMC_SYM_TRIPLE");
        if (d->C == 1 && d->ptr_count >= 2 && !d->denoise_film && !d->t_acc && !d->skip_top && !d->skip_bottom &&
            (pref == 0 || pref == 3) && !(e && atoi(e) == 0)) {
            SmcFilterParams p;
            d->tri = 1;
            d->rec_images = (d->ptr_count + 2) / 3;
            fill_filter_params(d, p);
            if (!smc_filter_sym_supported(p)) {
                d->tri = 0;
                d->rec_images = d->ptr_count;
            }
        }
    }
    d->flags_offset = ((d->rec_image_stride * d->rec_images + 255) / 256) * 256;
    // [records | halo flags | two lists of non-finite values (SmcNfList, step parity)]
    const size_t bytes = d->flags_offset + SMC_NF_OFFSET + 2 * sizeof(SmcNfList);
    cudaError_t e = cudaMalloc(&d->d_rec, bytes);
    if (e != cudaSuccess)
        SMC_FAIL(SMC_ERR_NOMEM, "cudaMalloc(%zu bytes of records) failed: %s", bytes, cudaGetErrorString(e));
    SMC_CUDA(cudaMemsetAsync(d->d_rec, 0, bytes, d->ctx->stream));
    d->d_flags = (int *)(d->d_rec + d->flags_offset);
    if (d->NGX > 0) {
        d->gext_stride = ((d->NGX + 3) / 4) * 4;
        const size_t eb = (size_t)d->rec_rows * d->rec_pitch * d->gext_stride * sizeof(float);
        if (cudaMalloc(&d->d_gext, eb) != cudaSuccess) {
            cudaGetLastError();
            SMC_FAIL(SMC_ERR_NOMEM, "cudaMalloc(%zu bytes of G-buffer extension) failed", eb);
        }
        SMC_CUDA(cudaMemsetAsync(d->d_gext, 0, eb, d->ctx->stream));
    }
    return SMC_OK;
}

static void fill_filter_params(const smc_denoiser *d, SmcFilterParams &p) {
    p.W = d->W; p.H = d->H; p.C = d->C; p.NG = d->NG; p.radius = d->radius; p.mode = d->mode;
    p.ptr_count = d->rec_images; p.denoise_film = d->denoise_film; p.sm_count = d->ctx->sm_count;
    p.tri = d->tri; p.images = d->ptr_count;
    p.row_begin = d->row_begin; p.row_end = d->row_end;
    p.padX = d->padX; p.rec_pitch = d->rec_pitch; p.rec_image_stride = d->rec_image_stride; p.rec = d->d_rec;
    p.sw = d->d_sw; p.sw_stride = d->sw_stride; p.sw_margin_y = SMC_SW_MARGIN_Y; p.sw_margin_x = SMC_SW_MARGIN_X;
    p.out_ptrs = d->t_out; p.film_filtered = d->film_filtered; p.accepted = d->t_acc; p.trace = d->d_trace; p.tile_counter = d->d_tile_counter;
    p.gext = d->d_gext; p.NGX = d->NGX; p.gext_stride = d->gext_stride;
    std::memset(&p.halo, 0, sizeof(p.halo));
}

static int select_kernel(smc_denoiser *d) {
    SmcFilterParams p;
    fill_filter_params(d, p);
    const char *nm = nullptr;
    const bool ok = smc_filter_stream_supported(p, d->ctx->sm_count, &nm);
    if (d->kernel_pref == 2 && !ok)
        SMC_FAIL(SMC_ERR_UNSUPPORTED, "streaming kernel requested but not available for C=%d NG=%d r=%d", d->C, d->NG,
                 d->radius);
    // kernel choice: 0 = auto (symmetric where it applies, else one-sided streaming, else generic), 1 = generic,
    // 2 = one-sided streaming, 3 = symmetric.  SMC_FILTER_KERNEL=stream|generic overrides `auto` for A/B runs.
    const bool sym_ok = d->d_sym_sw != nullptr && smc_filter_sym_supported(p);
    if (d->kernel_pref == 3 && !sym_ok)
        SMC_FAIL(SMC_ERR_UNSUPPORTED, "symmetric kernel requested but not available for C=%d mode=%d r=%d", d->C, d->mode,
                 d->radius);
    int pref = d->kernel_pref;
    if (pref == 0)
        if (const char *e = getenv("SMC_FILTER_KERNEL")) pref = !strcmp(e, "stream") ? 2 : !strcmp(e, "generic") ? 1 : 0;
    d->use_sym = sym_ok && (pref == 0 || pref == 3);
    d->use_stream = !d->use_sym && ok && pref != 1;
    if (d->tri && !d->use_sym) SMC_FAIL(SMC_ERR_UNSUPPORTED, "three-image records were laid out but the symmetric kernel is not selected");
    // output rows per thread: 2 x 2 pixels per thread wastes less work at the rim of the window (x1.06 at r = 20,
    // x1.23 at r = 6, against x1.13 / x1.46 for 2 x 4) and leaves room for 3 CTAs per SM; measured faster at every
    // radius tried (profiles/r1_variants.md).  SMC_STREAM_PY=4 selects the 2 x 4 variant for experiments.
    d->py = 2;
    if (const char *e = getenv("SMC_STREAM_PY")) d->py = atoi(e) == 4 ? 4 : 2;
    if (!d->d_tile_counter) SMC_CUDA(cudaMalloc(&d->d_tile_counter, sizeof(int)));
    if (!d->d_sync_counters) {
        SMC_CUDA(cudaMalloc(&d->d_sync_counters, 2 * sizeof(int)));
        SMC_CUDA(cudaMemsetAsync(d->d_sync_counters, 0, 2 * sizeof(int), d->ctx->stream));
    }
    if (getenv("SMC_STREAM_TRACE") && !d->d_trace) {
        SMC_CUDA(cudaMalloc(&d->d_trace, 4096 * 4 * sizeof(unsigned long long)));
        SMC_CUDA(cudaMemset(d->d_trace, 0, 4096 * 4 * sizeof(unsigned long long)));
    }
    snprintf(d->kernel_name, sizeof(d->kernel_name), "%s", d->use_sym ? "sym" : d->use_stream ? "stream" : "generic");
    return SMC_OK;
}

static int validate_common(int channels, int ptr_count, int width, int height, int radius, int n_gbufs) {
    if (channels != 1 && channels != 3) SMC_FAIL(SMC_ERR_INVALID, "channels must be 1 or 3, got %d", channels);
    if (ptr_count < 1 || ptr_count > SMC_MAX_DIM) SMC_FAIL(SMC_ERR_INVALID, "ptr_count %d out of range", ptr_count);
    if (width < 1 || height < 1 || width > SMC_MAX_DIM || height > SMC_MAX_DIM)
        SMC_FAIL(SMC_ERR_INVALID, "size %d x %d out of range (reference: unsigned short)", width, height);
    if (radius < 0 || radius > SMC_MAX_RADIUS)
        SMC_FAIL(SMC_ERR_INVALID, "radius %d out of range (reference: unsigned char)", radius);
    if (n_gbufs < 0 || n_gbufs > 255) SMC_FAIL(SMC_ERR_INVALID, "n_gbufs %d out of range", n_gbufs);
    return SMC_OK;
}

static bool plane_ok(const smc_plane *arr, int count) {
    if (!arr) return false;
    for (int i = 0; i < count; i++)
        if (!arr[i].dev) return false;
    return true;
}

extern "C" int smc_denoiser_create(smc_context *ctx, const smc_filter_desc *desc, smc_denoiser **out) {
    if (!ctx || !desc || !out) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    *out = nullptr;
    int rc = validate_common(desc->channels, desc->ptr_count, desc->width, desc->height, desc->radius, desc->n_gbufs);
    if (rc) return rc;
    if (desc->membership != SMC_MEMBER_WELCH && desc->membership != SMC_MEMBER_MOON)
        SMC_FAIL(SMC_ERR_INVALID, "bad membership %d", desc->membership);
    const int pc = desc->ptr_count;
    if (!plane_ok(desc->n, pc) || !plane_ok(desc->mean, pc) || !plane_ok(desc->m2, pc))
        SMC_FAIL(SMC_ERR_INVALID, "n/mean/m2 planes are required for every image");
    if (desc->membership == SMC_MEMBER_WELCH && !plane_ok(desc->m3, pc))
        SMC_FAIL(SMC_ERR_INVALID, "m3 planes are required for the Welch membership");
    if (!plane_ok(desc->film_ptrs, pc) && !(desc->channels == 3 && desc->denoise_film && pc == 1))
        SMC_FAIL(SMC_ERR_INVALID, "film_ptrs planes are required");
    if (!plane_ok(desc->film_filtered_ptrs, pc) && !(desc->channels == 3 && desc->denoise_film && pc == 1))
        SMC_FAIL(SMC_ERR_INVALID, "film_filtered_ptrs planes are required");
    if (desc->denoise_film && (!desc->film.dev || !desc->film_filtered.dev))
        SMC_FAIL(SMC_ERR_INVALID, "denoise_film needs film and film_filtered");
    int ng = 0;
    for (int g = 0; g < desc->n_gbufs; g++) {
        if (!desc->gbufs || !desc->gbufs[g].dev || !desc->gbuf_channels || !desc->gbuf_dr_factors)
            SMC_FAIL(SMC_ERR_INVALID, "G-buffer %d incomplete", g);
        // dr2 silently ignores buffers whose channel count is neither 1 nor 3 (stat_denoiser.cu:103-110): so do we
        if (desc->gbuf_channels[g] != 1 && desc->gbuf_channels[g] != 3) continue;
        if (!(desc->gbuf_dr_factors[g] <= 0.f))
            SMC_FAIL(SMC_ERR_INVALID, "G-buffer %d: drFactor %g must be <= 0 (-0.5/sd^2)", g, desc->gbuf_dr_factors[g]);
        ng += desc->gbuf_channels[g];
    }
    if (ng > SMC_MAX_GBUF_CHANNELS)
        SMC_FAIL(SMC_ERR_UNSUPPORTED, "%d flattened G-buffer channels; this build handles %d", ng, SMC_MAX_GBUF_CHANNELS);
    if (ng > SMC_REC_GBUF_CHANNELS && (desc->halo_top_external || desc->halo_bottom_external))
        SMC_FAIL(SMC_ERR_UNSUPPORTED, "more than %d G-buffer channels are not supported with external record halos",
                 SMC_REC_GBUF_CHANNELS);
    int rb = desc->row_begin, re = desc->row_end;
    if (rb == 0 && re == 0) re = desc->height;
    if (rb < 0 || re > desc->height || rb > re) SMC_FAIL(SMC_ERR_INVALID, "bad row range [%d, %d)", rb, re);

    SMC_CUDA(cudaSetDevice(ctx->device));
    smc_denoiser *d = new (std::nothrow) smc_denoiser;
    if (!d) SMC_FAIL(SMC_ERR_NOMEM, "out of host memory");
    d->ctx = ctx; d->C = desc->channels; d->ptr_count = pc; d->W = desc->width; d->H = desc->height;
    d->radius = desc->radius; d->denoise_film = desc->denoise_film ? 1 : 0; d->mode = desc->membership;
    d->n_gbufs = desc->n_gbufs; d->NG = std::min(ng, SMC_REC_GBUF_CHANNELS); d->NGX = ng - d->NG;
    d->row_begin = rb; d->row_end = re; d->ds_factor = desc->ds_factor;
    d->skip_top = desc->halo_top_external ? 1 : 0; d->skip_bottom = desc->halo_bottom_external ? 1 : 0;
    d->kernel_pref = desc->kernel;

    // descriptor tables: 9 per-image families + G-buffers, one allocation
    const int fam = 9;
    std::vector<SmcPtrStepSz> h((size_t)fam * pc + std::max(desc->n_gbufs, 1));
    const smc_plane *src[fam] = {desc->n, desc->mean, desc->m2, desc->m3, desc->film_ptrs, desc->mean_corr,
                                 desc->disc, desc->film_filtered_ptrs, desc->accepted};
    for (int f = 0; f < fam; f++)
        for (int i = 0; i < pc; i++) {
            SmcPtrStepSz &e = h[(size_t)f * pc + i];
            e.data = src[f] ? (unsigned char *)src[f][i].dev : nullptr;
            e.step = src[f] ? src[f][i].step : 0;
            e.cols = d->W;
            e.rows = d->H;
        }
    for (int g = 0; g < desc->n_gbufs; g++) {
        SmcPtrStepSz &e = h[(size_t)fam * pc + g];
        e.data = (unsigned char *)desc->gbufs[g].dev;
        e.step = desc->gbufs[g].step;
        e.cols = d->W;
        e.rows = d->H;
    }
    auto fail = [&](int code) {
        smc_denoiser_destroy(d);
        return code;
    };
    if (cudaMalloc(&d->d_tables, h.size() * sizeof(SmcPtrStepSz)) != cudaSuccess ||
        cudaMalloc(&d->d_gch, std::max(desc->n_gbufs, 1)) != cudaSuccess ||
        cudaMalloc(&d->d_gf, sizeof(float) * std::max(desc->n_gbufs, 1)) != cudaSuccess) {
        smc_set_error("cudaMalloc(descriptor tables) failed");
        return fail(SMC_ERR_NOMEM);
    }
    cudaMemcpyAsync(d->d_tables, h.data(), h.size() * sizeof(SmcPtrStepSz), cudaMemcpyHostToDevice, ctx->stream);
    if (desc->n_gbufs > 0) {
        cudaMemcpyAsync(d->d_gch, desc->gbuf_channels, desc->n_gbufs, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d->d_gf, desc->gbuf_dr_factors, sizeof(float) * desc->n_gbufs, cudaMemcpyHostToDevice,
                        ctx->stream);
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        smc_set_error("uploading descriptor tables failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(SMC_ERR_CUDA);
    }
    d->t_n = d->d_tables + 0 * pc; d->t_mean = d->d_tables + 1 * pc; d->t_m2 = d->d_tables + 2 * pc;
    d->t_m3 = d->d_tables + 3 * pc; d->t_film = d->d_tables + 4 * pc;
    d->t_mc = desc->mean_corr ? d->d_tables + 5 * pc : nullptr;
    d->t_disc = desc->disc ? d->d_tables + 6 * pc : nullptr;
    d->t_out = d->d_tables + 7 * pc;
    d->t_acc = desc->accepted ? d->d_tables + 8 * pc : nullptr;
    d->t_gbufs = d->d_tables + (size_t)fam * pc;
    d->h_tables = h;
    if (desc->n_gbufs > 0) {
        d->h_gch.assign(desc->gbuf_channels, desc->gbuf_channels + desc->n_gbufs);
        d->h_gf.assign(desc->gbuf_dr_factors, desc->gbuf_dr_factors + desc->n_gbufs);
    }
    d->film = SmcPtrStepSz{(unsigned char *)desc->film.dev, desc->film.step, d->W, d->H};
    d->film_filtered = SmcPtrStepSz{(unsigned char *)desc->film_filtered.dev, desc->film_filtered.step, d->W, d->H};

    if ((rc = alloc_records(d)) || (rc = build_spatial_table(d)) || (rc = build_sym_table(d)) || (rc = select_kernel(d)))
        return fail(rc);
    *out = d;
    return SMC_OK;
}

extern "C" void smc_denoiser_destroy(smc_denoiser *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    if (d->d_trace) {
        if (const char *path = getenv("SMC_STREAM_TRACE")) {
            std::vector<unsigned long long> h(4096 * 4);
            if (cudaMemcpy(h.data(), d->d_trace, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess)
                if (FILE *f = fopen(path, "w")) {
                    for (int i = 0; i < 4096 && h[4 * i]; i++)
                        fprintf(f, "%d %llu %llu %llu %llu\n", i, h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
                    fclose(f);
                }
        }
        cudaFree(d->d_trace);
    }
    if (!d->tables_external) cudaFree(d->d_tables);
    cudaFree(d->d_gch);
    cudaFree(d->d_gf);
    cudaFree(d->d_sw);
    cudaFree(d->d_rowrange);
    cudaFree(d->d_sym_sw);
    cudaFree(d->d_sym_rowrange);
    cudaFree(d->d_sym_scratch);
    cudaFree(d->d_sym_scratch2);
    cudaFree(d->d_sym_fwd2);
    cudaFree(d->d_sym_scratch_cnt);
    cudaFree(d->d_sym_fwd);
    cudaFree(d->d_sym_fwd_cnt);
    for (int w = 0; w < 2; w++)
        if (d->peer[w].ipc_base) cudaIpcCloseMemHandle(d->peer[w].ipc_base);
    cudaFree(d->d_rec);
    cudaFree(d->d_gext);
    cudaFree(d->d_tile_counter);
    cudaFree(d->d_sync_counters);
    for (cudaEvent_t e : d->events) cudaEventDestroy(e);
    if (d->s_in) cudaStreamDestroy(d->s_in);
    if (d->s_out) cudaStreamDestroy(d->s_out);
    delete d;
}

static bool has_peers(const smc_denoiser *d);
static SmcNfList *nf_list(const smc_denoiser *d, const int *flags);

// prepass over image rows [y0, y1); the replicated rows above row 0 / below row H-1 go with the first / last rows
static int prepass_rows(smc_denoiser *d, int y0, int y1, const SmcHaloSync *halo = nullptr) {
    SMC_CUDA(cudaSetDevice(d->ctx->device));
    SmcPrepassParams p;
    p.nf = nf_list(d, d->d_flags);
    // a frame's first rows start its list of non-finite values (row chunks of one frame come in ascending order)
    if (y0 == 0) SMC_CUDA(cudaMemsetAsync(&p.nf->count, 0, sizeof(int), d->ctx->stream));
    std::memset(&p.halo, 0, sizeof(p.halo));
    p.halo_blocks = 0;
    p.W = d->W; p.H = d->H; p.C = d->C; p.ptr_count = d->rec_images; p.radius = d->radius; p.mode = d->mode;
    p.triple = d->tri; p.images = d->ptr_count;
    p.denoise_film = d->denoise_film; p.padX = d->padX; p.rec_pitch = d->rec_pitch;
    p.rec_image_stride = d->rec_image_stride; p.rec = d->d_rec; p.skip_top = d->skip_top; p.skip_bottom = d->skip_bottom;
    p.pr_begin = y0 == 0 ? 0 : y0 + d->radius;
    p.pr_end = y1 == d->H ? d->H + 2 * d->radius : y1 + d->radius;
    p.n = d->t_n; p.mean = d->t_mean; p.m2 = d->t_m2; p.m3 = d->t_m3; p.film_ptrs = d->t_film; p.film = d->film;
    p.gbufs = d->t_gbufs; p.NG = d->NG; p.NGX = d->NGX; p.gext = d->d_gext; p.gext_stride = d->gext_stride;
    for (int g = 0, k = 0; g < d->n_gbufs; g++) {
        if (d->h_gch[g] != 1 && d->h_gch[g] != 3) continue;  // ignored, as in dr2 (stat_denoiser.cu:103-110)
        for (int c = 0; c < d->h_gch[g] && k < SMC_MAX_GBUF_CHANNELS; c++, k++) {
            p.g_buf[k] = (unsigned char)g; p.g_ch[k] = (unsigned char)c; p.g_nch[k] = d->h_gch[g];
            p.g_scale[k] = sqrtf(-d->h_gf[g] * 1.4426950408889634f);
        }
    }
    p.mean_corr = d->t_mc; p.disc = d->t_disc; p.lut = d->ctx->d_lut;
    const size_t row_bytes = smc_rec_row_bytes(d->rec_pitch);
    const smc_denoiser::Peer &up = d->peer[0], &dn = d->peer[1];
    // bottom halo of the rank above: its padded rows [H_up + r, H_up + 2r); top halo of the rank below: padded rows [0, r)
    p.peer_up_halo = up.rec ? up.rec + (size_t)(up.H + d->radius) * row_bytes : nullptr;
    p.peer_down_halo = dn.rec ? dn.rec : nullptr;
    p.peer_up_image_stride = up.image_stride;
    p.peer_down_image_stride = dn.image_stride;
    if (halo) {
        p.halo = *halo;
        // blocks whose row goes into a neighbour: rows [0, r) when a rank sits above, [H - r, H) when one sits below
        const int r = d->radius, H = d->H;
        int rows = 0;
        for (int y = 0; y < H; y++) rows += ((up.rec && y < r) || (dn.rec && y >= H - r)) ? 1 : 0;
        p.halo_blocks = rows * ((d->rec_pitch + 255) / 256) * d->rec_images;
    }
    return smc_launch_prepass(d->ctx, p);
}

static bool has_peers(const smc_denoiser *d) { return d->peer[0].rec || d->peer[1].rec; }

// the list of non-finite values of the current frame: with peer halos the two lists alternate with the protocol's step, so
// that a neighbour can still read the entries of its halo rows while this rank lists the next frame
static SmcNfList *nf_list(const smc_denoiser *d, const int *flags) {
    return (SmcNfList *)((unsigned char *)flags + SMC_NF_OFFSET) + (has_peers(d) ? (d->step & 1) : 0);
}

// the listed non-finite values of this frame go to the centres of rows [p.row_begin, p.row_end) they are member taps of
static int nonfinite_fixup(const smc_denoiser *d, SmcFilterParams p, const SmcHaloSync &release) {
    // with peer halos this kernel, not the filter, releases the neighbours' halo rows: it still reads records of the halo
    p.halo = release;
    SmcNfSources src;
    std::memset(&src, 0, sizeof(src));
    const int r = d->radius;
    src.list[0] = nf_list(d, d->d_flags);
    src.row_lo[0] = 0; src.row_hi[0] = 1 << 30; src.row_shift[0] = 0;
    if (d->peer[0].rec) {  // rank above: its last r rows (padded rows [H_up, H_up + r)) are this rank's padded rows [0, r)
        src.list[1] = nf_list(d, d->peer[0].flags);
        src.row_lo[1] = d->peer[0].H; src.row_hi[1] = d->peer[0].H + r; src.row_shift[1] = -d->peer[0].H;
    }
    if (d->peer[1].rec) {  // rank below: its first r rows (padded rows [r, 2r)) are this rank's padded rows [H + r, H + 2r)
        src.list[2] = nf_list(d, d->peer[1].flags);
        src.row_lo[2] = r; src.row_hi[2] = 2 * r; src.row_shift[2] = d->H;
    }
    return smc_launch_nonfinite_fixup(d->ctx, p, src);
}

// filter over output rows [y0, y1)
static int filter_rows(smc_denoiser *d, int y0, int y1, const SmcHaloSync *halo = nullptr) {
    SMC_CUDA(cudaSetDevice(d->ctx->device));
    if (y1 <= y0) return SMC_OK;
    SmcFilterParams p;
    fill_filter_params(d, p);
    SmcHaloSync release;
    std::memset(&release, 0, sizeof(release));
    if (halo) {  // the filter kernel waits for the halo rows; the fix-up kernel after it releases them
        p.halo = *halo;
        p.halo.signal0 = p.halo.signal1 = nullptr;
        release = *halo;
        release.wait0 = release.wait1 = nullptr;
    }
    p.row_begin = y0;
    p.row_end = y1;
    if (d->use_sym) {
        SmcSymParams g;
        size_t smem = 0;
        if (!smc_filter_sym_geometry(p, g, smem)) SMC_FAIL(SMC_ERR_UNSUPPORTED, "symmetric filter: geometry not supported");
        // scratch for the partial mirror sums and the forward sums of this row range (grown on demand; a launch over fewer
        // rows needs less)
        const size_t se = smc_filter_sym_scratch_elems(p, g), fe = (size_t)d->rec_images * (y1 - y0) * d->W;
        if (se > d->sym_scratch_elems || fe > d->sym_fwd_elems) {
            // both arrays grow to the largest need seen so far (row chunks of different lengths alternate in the host pipeline,
            // and a shorter chunk with shorter work units can need MORE scratch rows than a longer one: never shrink)
            const size_t se2 = std::max(se, d->sym_scratch_elems), fe2 = std::max(fe, d->sym_fwd_elems);
            SMC_CUDA(cudaStreamSynchronize(d->ctx->stream));
            cudaFree(d->d_sym_scratch); cudaFree(d->d_sym_scratch_cnt); cudaFree(d->d_sym_fwd); cudaFree(d->d_sym_fwd_cnt);
            cudaFree(d->d_sym_scratch2); cudaFree(d->d_sym_fwd2);
            d->d_sym_scratch = d->d_sym_fwd = nullptr;
            d->d_sym_scratch2 = d->d_sym_fwd2 = nullptr;
            d->d_sym_scratch_cnt = d->d_sym_fwd_cnt = nullptr;
            d->sym_scratch_elems = d->sym_fwd_elems = 0;
            if (cudaMalloc(&d->d_sym_scratch, se2 * sizeof(float4)) != cudaSuccess ||
                cudaMalloc(&d->d_sym_fwd, fe2 * sizeof(float4)) != cudaSuccess ||
                (d->tri && (cudaMalloc(&d->d_sym_scratch2, se2 * sizeof(float2)) != cudaSuccess ||
                            cudaMalloc(&d->d_sym_fwd2, fe2 * sizeof(float2)) != cudaSuccess)) ||
                (d->t_acc && (cudaMalloc(&d->d_sym_scratch_cnt, se2 * sizeof(int)) != cudaSuccess ||
                              cudaMalloc(&d->d_sym_fwd_cnt, fe2 * sizeof(int)) != cudaSuccess))) {
                cudaGetLastError();
                SMC_FAIL(SMC_ERR_NOMEM, "cudaMalloc(%zu MB of symmetric-filter scratch) failed",
                         (se2 * 16 + fe2 * 16) >> 20);
            }
            d->sym_scratch_elems = se2;
            d->sym_fwd_elems = fe2;
        }
        g.sw = d->d_sym_sw; g.rowrange = d->d_sym_rowrange; g.sw_special = d->sym_sw_special;
        g.scratch = d->d_sym_scratch; g.scratch_cnt = d->d_sym_scratch_cnt; g.fwd = d->d_sym_fwd; g.fwd_cnt = d->d_sym_fwd_cnt;
        g.scratch2 = d->d_sym_scratch2; g.fwd2 = d->d_sym_fwd2;
        g.unit_counter = d->d_tile_counter;
        const char *nm = nullptr;
        const int rc = smc_launch_filter_sym(d->ctx, p, g, smem, &nm);
        if (nm) snprintf(d->kernel_name, sizeof(d->kernel_name), "%s", nm);
        return rc ? rc : nonfinite_fixup(d, p, release);
    }
    if (d->use_stream) {
        const char *nm = nullptr;
        const int rc = smc_launch_filter_stream(d->ctx, p, d->d_rowrange, d->py, &nm);
        if (nm) snprintf(d->kernel_name, sizeof(d->kernel_name), "%s", nm);
        return rc ? rc : nonfinite_fixup(d, p, release);
    }
    snprintf(d->kernel_name, sizeof(d->kernel_name), "generic<C=%d,NG=%d,%s>", d->C, d->NG, d->mode ? "moon" : "welch");
    const int rc = smc_launch_filter_generic(d->ctx, p);
    return rc ? rc : nonfinite_fixup(d, p, release);
}

static int check_rows(const smc_denoiser *d, int y0, int y1, int lo, int hi) {
    if (!d) SMC_FAIL(SMC_ERR_INVALID, "NULL plan");
    if (y0 < lo || y1 > hi || y0 > y1) SMC_FAIL(SMC_ERR_INVALID, "bad row range [%d, %d) (allowed [%d, %d))", y0, y1, lo, hi);
    return SMC_OK;
}

extern "C" int smc_denoiser_prepass(smc_denoiser *d) {
    if (!d) SMC_FAIL(SMC_ERR_INVALID, "NULL plan");
    if (!has_peers(d)) return prepass_rows(d, 0, d->H);
    // Halo protocol, step k: the neighbours must have finished FILTERING step k-1 before their halo rows are overwritten
    // (they signal free_from_* = k-1 into OUR flags); after the prepass (own records + halo rows stored into the
    // neighbours' arrays) tell them their halos of step k are ready.  Both ends live inside the prepass kernel: the blocks
    // that store into a neighbour wait for its flag, the last of them to retire signals (SmcHaloSync).
    d->step++;
    SmcHaloSync h;
    h.wait0 = d->peer[0].rec ? d->d_flags + 2 : nullptr;
    h.wait1 = d->peer[1].rec ? d->d_flags + 3 : nullptr;
    h.wait_value = d->step - 1;
    // the rank above sees us as "down", the rank below as "up"
    h.signal0 = d->peer[0].flags ? d->peer[0].flags + 1 : nullptr;
    h.signal1 = d->peer[1].flags ? d->peer[1].flags + 0 : nullptr;
    h.signal_value = d->step;
    h.done_counter = d->d_sync_counters + 0;
    return prepass_rows(d, 0, d->H, &h);
}

extern "C" int smc_denoiser_prepass_rows(smc_denoiser *d, int row_begin, int row_end) {
    int rc = check_rows(d, row_begin, row_end, 0, d ? d->H : 0);
    if (rc) return rc;
    if (row_begin == row_end) return SMC_OK;
    return prepass_rows(d, row_begin, row_end);
}

extern "C" int smc_denoiser_filter(smc_denoiser *d) {
    if (!d) SMC_FAIL(SMC_ERR_INVALID, "NULL plan");
    if (!has_peers(d)) return filter_rows(d, d->row_begin, d->row_end);
    // wait until both neighbours have stored this step's halo rows into our array, filter, then release their halos
    SmcHaloSync h;
    h.wait0 = d->peer[0].rec ? d->d_flags + 0 : nullptr;
    h.wait1 = d->peer[1].rec ? d->d_flags + 1 : nullptr;
    h.wait_value = d->step;
    h.signal0 = d->peer[0].flags ? d->peer[0].flags + 3 : nullptr;
    h.signal1 = d->peer[1].flags ? d->peer[1].flags + 2 : nullptr;
    h.signal_value = d->step;
    h.done_counter = d->d_sync_counters + 1;
    SmcFilterParams fp;
    fill_filter_params(d, fp);
    if (d->use_sym || (d->use_stream && smc_filter_stream_syncs_halo(fp, d->py)))
        return filter_rows(d, d->row_begin, d->row_end, &h);  // the kernel waits and signals itself
    int rc = smc_launch_halo_wait(d->ctx, h.wait0, h.wait1, h.wait_value);
    if (rc) return rc;
    if ((rc = filter_rows(d, d->row_begin, d->row_end))) return rc;
    return smc_launch_halo_signal(d->ctx, h.signal0, h.signal1, h.signal_value);
}

extern "C" int smc_denoiser_filter_rows(smc_denoiser *d, int row_begin, int row_end) {
    int rc = check_rows(d, row_begin, row_end, d ? d->row_begin : 0, d ? d->row_end : 0);
    if (rc) return rc;
    return filter_rows(d, row_begin, row_end);
}

extern "C" int smc_denoiser_run(smc_denoiser *d) {
    int rc = smc_denoiser_prepass(d);
    if (rc) return rc;
    return smc_denoiser_filter(d);
}

// ---------------------------------------------------------------------------------------------------------
// Host-pipelined run: Estimator::Upload -> Denoise -> Download (estimator.cpp:409-489) as ONE call in which the
// PCIe copies overlap the kernels.  The image is cut into row chunks; chunk k is uploaded on a copy stream while
// chunk k-1 is prepassed and the rows whose whole window is already packed (y <= chunk_end - radius) are filtered on
// the context stream, and finished output rows are downloaded on a third stream.  Results are identical to
// upload-all / run / download-all: the same kernels run over the same rows, only in row order.
// ---------------------------------------------------------------------------------------------------------
static cudaEvent_t get_event(smc_denoiser *d, size_t i) {
    while (d->events.size() <= i) {
        cudaEvent_t e = nullptr;
        // SMC_PIPE_TRACE=1: events carry timestamps so that smc_denoiser_run_host can print its own timeline
        static const bool timing = getenv("SMC_PIPE_TRACE") != nullptr;
        if (cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) return nullptr;
        d->events.push_back(e);
    }
    return d->events[i];
}

static int auto_chunk_rows(const smc_denoiser *d) {
    const int rows = d->row_end - d->row_begin;
    int chunk;
    if (d->use_stream) {
        // whole waves of the persistent grid per chunk: tiles(chunk) = k * grid, about 8 chunks per frame
        SmcFilterParams p;
        fill_filter_params(d, p);
        int tile_w = 256;
        const int grid = smc_filter_stream_resident_ctas(p, d->py, d->ctx->sm_count, &tile_w);
        const int tiles_x = (d->W + tile_w - 1) / tile_w;
        const long long total = (long long)tiles_x * ((rows + d->py - 1) / d->py);
        long long k = (total / std::max(grid, 1) + 4) / 8;
        if (k < 1) k = 1;
        chunk = (int)(k * grid / tiles_x) * d->py;
    } else if (d->use_sym) {
        // every launch of the symmetric kernel re-evaluates the `radius` rows above its range as virtual centres (about
        // 0.42 radius rows of work): about six chunks per frame, none shorter than four radii
        chunk = std::max(((rows / 6 + 7) / 8) * 8, 4 * d->radius);
    } else {
        chunk = ((rows / 8 + 7) / 8) * 8;
    }
    return std::max(chunk, std::max(2 * d->radius, 16));
}

// rows [bounds[k], bounds[k+1]) form chunk k of the host pipeline
static void chunk_schedule(smc_denoiser *d, int chunk_rows, std::vector<int> &bounds) {
    smc_context *ctx = d->ctx;
    const int H = d->H, r = d->radius;
    // Chunk boundaries: uniform, except for the last one (below).  (With the earlier, slower filter a schedule that ramped
    // up from a small first chunk and down to a small last one was measured SLOWER, 14.7 vs 14.1 ms at 4K: a filter launch
    // over few rows fills the persistent grid badly -- one tile is ~0.35 ms of work for a warp -- and compute, not PCIe, was
    // then the critical path.  A small FIRST chunk buys nothing in either regime: the bus is busy from t = 0 anyway.)
    // Chunk schedule: uniform chunks of whole waves of the persistent filter grid (about 8 per frame).  Timeline at 4K
    // (SMC_PIPE_TRACE=1, profiles/r1g_pipe_trace.txt): the bus delivers a 296-row chunk every 1.71 ms (50.5 GB/s against
    // 55.6 GB/s for one raw 630 MB copy), the filter needs 1.43 ms for it, so in steady state compute waits for PCIe, and
    // after the last byte has arrived (12.4 ms) the last full chunk's filter and the remainder's still run (14.05 ms).
    // SMC_PIPE_TAPER=1 selects a tapered schedule (two-wave first chunk, two-wave and one-wave chunks at the end) meant to
    // shorten that drain; it is exact (tests pass with it) but measured 13.93 ms against 13.74 ms uniform -- the extra
    // copies and small launches cost what the shorter drain saves -- so it is off by default.
    const bool auto_chunks = chunk_rows == 0;
    if (chunk_rows == 0) chunk_rows = auto_chunk_rows(d);
    int unit = 0;             // rows of one wave of the streaming grid
    if (auto_chunks && d->use_stream) {
        SmcFilterParams fp;
        fill_filter_params(d, fp);
        int tile_w = 256;
        const int grid = smc_filter_stream_resident_ctas(fp, d->py, ctx->sm_count, &tile_w);
        const int tiles_x = (d->W + tile_w - 1) / tile_w;
        unit = (grid / std::max(tiles_x, 1)) * d->py;
    }
    const char *tp = getenv("SMC_PIPE_TAPER");
    const bool taper = unit >= 2 * r && chunk_rows >= 4 * unit && H >= 3 * chunk_rows && (tp && atoi(tp) == 1);
    if (taper) {
        // whole waves everywhere but in the first chunk, which takes the remainder: its filter runs while the (longer)
        // upload of the second chunk is in flight, so a partly filled last wave costs nothing there
        const int prelast = 2 * unit, last = std::max(unit - (r - 1), 16) & ~1;
        const int body = H - prelast - last - 2 * unit;   // rows of the middle chunks + the first chunk's remainder
        const int nw = body / unit;                       // waves to distribute over the middle chunks
        const int first = 2 * unit + (body - nw * unit);
        const int nmid = std::max(1, (nw * unit + chunk_rows / 2) / chunk_rows);
        bounds.push_back(0);
        int a = first;
        for (int k = 0; k < nmid; k++) {
            bounds.push_back(a);
            a = first + (int)((long long)nw * (k + 1) / nmid) * unit;
        }
        bounds.push_back(H - prelast - last);
        bounds.push_back(H - last);
        bounds.push_back(H);
    } else if (auto_chunks && d->use_sym && !(tp && atoi(tp) == 0)) {
        // Symmetric kernel: uniform chunks while the bus is the critical path, then halving chunks.  Whatever is filtered
        // and downloaded after the last byte has arrived is pure latency: a 360-row last chunk costs 1.2 ms + 0.5 ms at 4K, a
        // 90-row one 0.35 ms + 0.12 ms; the extra launches run under the uploads.  (SMC_PIPE_TAPER=0: uniform chunks.)
        const int min_chunk = std::max(2 * r + 8, 48);
        int a = 0;
        bounds.push_back(0);
        while (H - a > 2 * chunk_rows) {
            a += chunk_rows;
            bounds.push_back(a);
        }
        while (H - a > 2 * min_chunk) {
            a += (((H - a) / 2 + 7) / 8) * 8;
            bounds.push_back(a);
        }
        bounds.push_back(H);
    } else {
        for (int a = 0; a < H; a += chunk_rows) bounds.push_back(a);
        if (bounds.back() < H) bounds.push_back(H);
    }

}

// one plane's rows moving between host and device inside the pipeline
struct Xfer {
    unsigned char *dev;
    size_t dev_step;
    smc_plane host;
    size_t row_bytes;
};

// The pipeline proper: chunk k is uploaded on the copy stream while chunk k-1 is prepassed and the rows whose window is complete
// are filtered on the context stream; finished rows of `downs` go back on a third stream (the first n_aux_downs of them -- mean-
// corr / discriminator planes -- are final right after the prepass).
static int run_pipeline(smc_denoiser *d, const std::vector<Xfer> &ups, const std::vector<Xfer> &downs, size_t n_aux_downs,
                        int chunk_rows) {
    smc_context *ctx = d->ctx;
    if (!d->s_in) SMC_CUDA(cudaStreamCreateWithFlags(&d->s_in, cudaStreamNonBlocking));
    if (!d->s_out) SMC_CUDA(cudaStreamCreateWithFlags(&d->s_out, cudaStreamNonBlocking));
    const int H = d->H, r = d->radius;
    std::vector<int> bounds;
    chunk_schedule(d, chunk_rows, bounds);
    size_t ev = 0;
    cudaEvent_t e_begin = get_event(d, ev++);
    if (!e_begin) SMC_FAIL(SMC_ERR_CUDA, "cudaEventCreate failed");
    // the copy streams start after whatever is already queued on the context stream
    SMC_CUDA(cudaEventRecord(e_begin, ctx->stream));
    SMC_CUDA(cudaStreamWaitEvent(d->s_in, e_begin, 0));
    SMC_CUDA(cudaStreamWaitEvent(d->s_out, e_begin, 0));

    const bool trace = getenv("SMC_PIPE_TRACE") != nullptr;
    struct ChunkEv { cudaEvent_t up, pre, filt, down; int a, b, f0, f1; };
    std::vector<ChunkEv> tr;
    int filtered_to = d->row_begin;  // output rows < filtered_to are done
    for (size_t ci = 0; ci + 1 < bounds.size(); ci++) {
        const int a = bounds[ci], b = bounds[ci + 1];
        for (const Xfer &x : ups)
            SMC_CUDA(cudaMemcpy2DAsync(x.dev + (size_t)a * x.dev_step, x.dev_step,
                                       (const char *)x.host.dev + (size_t)a * x.host.step, x.host.step, x.row_bytes,
                                       b - a, cudaMemcpyHostToDevice, d->s_in));
        cudaEvent_t e_up = get_event(d, ev++), e_f = get_event(d, ev++);
        if (!e_up || !e_f) SMC_FAIL(SMC_ERR_CUDA, "cudaEventCreate failed");
        SMC_CUDA(cudaEventRecord(e_up, d->s_in));
        SMC_CUDA(cudaStreamWaitEvent(ctx->stream, e_up, 0));
        int rc = prepass_rows(d, a, b);
        if (rc) return rc;
        cudaEvent_t e_pre = nullptr;
        if (trace) {
            e_pre = get_event(d, ev++);
            if (e_pre) cudaEventRecord(e_pre, ctx->stream);
        }
        // output row y reads record rows y-r .. y+r-1: complete once rows < b are packed  <=>  y <= b - r
        const int f_end = b == H ? d->row_end : std::min(d->row_end, std::max(filtered_to, b - r + 1));
        const int f_begin = filtered_to;
        if (f_end > f_begin) {
            rc = filter_rows(d, f_begin, f_end);
            if (rc) return rc;
            filtered_to = f_end;
        }
        SMC_CUDA(cudaEventRecord(e_f, ctx->stream));
        SMC_CUDA(cudaStreamWaitEvent(d->s_out, e_f, 0));
        for (size_t i = 0; i < downs.size(); i++) {
            const Xfer &x = downs[i];
            const int y0 = i < n_aux_downs ? a : f_begin, y1 = i < n_aux_downs ? b : f_end;
            if (y1 <= y0) continue;
            SMC_CUDA(cudaMemcpy2DAsync((char *)x.host.dev + (size_t)y0 * x.host.step, x.host.step,
                                       x.dev + (size_t)y0 * x.dev_step, x.dev_step, x.row_bytes, y1 - y0,
                                       cudaMemcpyDeviceToHost, d->s_out));
        }
        if (trace) {
            cudaEvent_t e_dn = get_event(d, ev++);
            if (e_dn) cudaEventRecord(e_dn, d->s_out);
            tr.push_back({e_up, e_pre, e_f, e_dn, a, b, f_begin, f_end});
        }
    }
    // join: smc_synchronize(ctx) (or anything queued later on the context stream) covers the downloads
    cudaEvent_t e_end = get_event(d, ev++);
    if (!e_end) SMC_FAIL(SMC_ERR_CUDA, "cudaEventCreate failed");
    SMC_CUDA(cudaEventRecord(e_end, d->s_out));
    SMC_CUDA(cudaStreamWaitEvent(ctx->stream, e_end, 0));
    if (trace) {  // blocking: a diagnostic, not a mode to run in
        cudaEventSynchronize(e_end);
        float t_end = 0.f;
        cudaEventElapsedTime(&t_end, e_begin, e_end);
        fprintf(stderr, "smc_denoiser_run_host timeline (ms after the first enqueue), total %.3f\n", t_end);
        for (const ChunkEv &c : tr) {
            float u = 0, p = 0, f = 0, dn = 0;
            cudaEventElapsedTime(&u, e_begin, c.up);
            if (c.pre) cudaEventElapsedTime(&p, e_begin, c.pre);
            cudaEventElapsedTime(&f, e_begin, c.filt);
            if (c.down) cudaEventElapsedTime(&dn, e_begin, c.down);
            fprintf(stderr, "  rows [%4d,%4d) uploaded %.3f  prepassed %.3f  rows [%4d,%4d) filtered %.3f  downloaded %.3f\n", c.a, c.b,
                    u, p, c.f0, c.f1, f, dn);
        }
    }
    return SMC_OK;
}

extern "C" int smc_denoiser_run_host(smc_denoiser *d, const smc_host_io *io, int chunk_rows) {
    if (!d || !io) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    if (d->tables_external) SMC_FAIL(SMC_ERR_UNSUPPORTED, "plan was built from device tables: host planes unknown");
    if (d->skip_top || d->skip_bottom)
        SMC_FAIL(SMC_ERR_UNSUPPORTED, "record-halo exchange plans cannot be pipelined from the host in one call");
    if (chunk_rows < 0) SMC_FAIL(SMC_ERR_INVALID, "chunk_rows < 0");
    smc_context *ctx = d->ctx;
    SMC_CUDA(cudaSetDevice(ctx->device));
    const int pc = d->ptr_count;
    std::vector<Xfer> ups, downs;
    auto add = [&](std::vector<Xfer> &v, const SmcPtrStepSz *dev, const smc_plane *host, size_t px_bytes) {
        if (!host || !host->dev || !dev->data) return;
        for (const Xfer &x : v)
            if (x.dev == dev->data) return;  // aliased planes (mean == film-mean when !transform) move once
        Xfer x{dev->data, dev->step, *host, (size_t)d->W * px_bytes};
        if (x.host.step == 0) x.host.step = x.row_bytes;
        v.push_back(x);
    };
    const SmcPtrStepSz *T = d->h_tables.data();
    const size_t cb = (size_t)d->C * 4;
    for (int i = 0; i < pc; i++) {
        add(ups, T + 0 * pc + i, io->n ? io->n + i : nullptr, 4);
        add(ups, T + 1 * pc + i, io->mean ? io->mean + i : nullptr, cb);
        add(ups, T + 2 * pc + i, io->m2 ? io->m2 + i : nullptr, cb);
        add(ups, T + 3 * pc + i, io->m3 ? io->m3 + i : nullptr, cb);
        add(ups, T + 4 * pc + i, io->film_ptrs ? io->film_ptrs + i : nullptr, cb);
        add(downs, T + 5 * pc + i, io->mean_corr ? io->mean_corr + i : nullptr, cb);
        add(downs, T + 6 * pc + i, io->disc ? io->disc + i : nullptr, cb);
    }
    add(ups, &d->film, &io->film, 12);
    for (int g = 0; g < d->n_gbufs; g++)
        add(ups, T + 9 * (size_t)pc + g, io->gbufs ? io->gbufs + g : nullptr, (size_t)d->h_gch[g] * 4);
    const size_t n_aux_downs = downs.size();  // mean-corr / discriminator rows are final right after the prepass
    for (int i = 0; i < pc; i++) add(downs, T + 7 * pc + i, io->film_filtered_ptrs ? io->film_filtered_ptrs + i : nullptr, cb);
    add(downs, &d->film_filtered, &io->film_filtered, 12);

    return run_pipeline(d, ups, downs, n_aux_downs, chunk_rows);
}

// ---------------------------------------------------------------------------------------------------------
// Peer halos: the multi-GPU exchange as stores into the neighbours' record arrays (see smc_prepass.cu)
// ---------------------------------------------------------------------------------------------------------
static int peer_check(smc_denoiser *d, int which, int peer_H, int peer_radius, int peer_pitch, int peer_pc) {
    if (!d) SMC_FAIL(SMC_ERR_INVALID, "NULL plan");
    if (which != 0 && which != 1) SMC_FAIL(SMC_ERR_INVALID, "which must be 0 (rank above) or 1 (rank below)");
    if (which == 0 ? !d->skip_top : !d->skip_bottom)
        SMC_FAIL(SMC_ERR_INVALID, "the plan was not created with halo_%s_external", which == 0 ? "top" : "bottom");
    if (peer_radius != d->radius || peer_pitch != d->rec_pitch || peer_pc != d->ptr_count)
        SMC_FAIL(SMC_ERR_INVALID, "peer plan has a different radius / width / image count");
    if (peer_H < d->radius || d->H < d->radius)
        SMC_FAIL(SMC_ERR_UNSUPPORTED, "bands (%d and %d rows) must be at least `radius` = %d rows", d->H, peer_H, d->radius);
    if (d->peer[which].rec) SMC_FAIL(SMC_ERR_INVALID, "peer %d already attached", which);
    return SMC_OK;
}

extern "C" int smc_denoiser_peer_export(smc_denoiser *d, smc_peer_info *out) {
    if (!d || !out) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    std::memset(out, 0, sizeof(*out));
    SMC_CUDA(cudaSetDevice(d->ctx->device));
    cudaIpcMemHandle_t h;
    SMC_CUDA(cudaIpcGetMemHandle(&h, d->d_rec));
    static_assert(sizeof(h) == sizeof(out->ipc_handle), "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(out->ipc_handle, &h, sizeof(h));
    out->image_stride = d->rec_image_stride;
    out->flags_offset = d->flags_offset;
    out->height = d->H; out->radius = d->radius; out->rec_pitch = d->rec_pitch; out->ptr_count = d->ptr_count;
    out->device = d->ctx->device;
    return SMC_OK;
}

extern "C" int smc_denoiser_peer_attach(smc_denoiser *d, int which, const smc_peer_info *info) {
    if (!info) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    int rc = peer_check(d, which, info->height, info->radius, info->rec_pitch, info->ptr_count);
    if (rc) return rc;
    SMC_CUDA(cudaSetDevice(d->ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, info->ipc_handle, sizeof(h));
    void *base = nullptr;
    SMC_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    smc_denoiser::Peer &p = d->peer[which];
    p.ipc_base = base;
    p.rec = (unsigned char *)base;
    p.flags = (int *)((unsigned char *)base + info->flags_offset);
    p.image_stride = info->image_stride;
    p.H = info->height;
    return SMC_OK;
}

extern "C" int smc_denoiser_peer_attach_local(smc_denoiser *d, int which, smc_denoiser *other) {
    if (!other) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    int rc = peer_check(d, which, other->H, other->radius, other->rec_pitch, other->ptr_count);
    if (rc) return rc;
    if (other->ctx->device != d->ctx->device) {
        SMC_CUDA(cudaSetDevice(d->ctx->device));
        int can = 0;
        SMC_CUDA(cudaDeviceCanAccessPeer(&can, d->ctx->device, other->ctx->device));
        if (!can) SMC_FAIL(SMC_ERR_UNSUPPORTED, "device %d cannot access device %d", d->ctx->device, other->ctx->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(other->ctx->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SMC_CUDA(e);
        cudaGetLastError();
    }
    smc_denoiser::Peer &p = d->peer[which];
    p.rec = other->d_rec;
    p.flags = other->d_flags;
    p.image_stride = other->rec_image_stride;
    p.H = other->H;
    return SMC_OK;
}

extern "C" int smc_denoiser_halo(smc_denoiser *d, int z, int which, void **dev, size_t *bytes) {
    if (!d || !dev || !bytes) SMC_FAIL(SMC_ERR_INVALID, "NULL argument");
    if (z < 0 || z >= d->rec_images) SMC_FAIL(SMC_ERR_INVALID, "record image %d out of range", z);
    const int r = d->radius;
    if (r > d->H) SMC_FAIL(SMC_ERR_UNSUPPORTED, "band of %d rows is shorter than the radius %d", d->H, r);
    const size_t row_bytes = smc_rec_row_bytes(d->rec_pitch);
    unsigned char *img = d->d_rec + (size_t)z * d->rec_image_stride;
    int first;  // record row (record row k holds y = k - r)
    switch (which) {
        case 0: first = r; break;                 // own rows 0 .. r-1
        case 1: first = d->H; break;              // own rows H-r .. H-1  -> record rows H .. H+r-1
        case 2: first = 0; break;                 // halo above: y = -r .. -1
        case 3: first = d->H + r; break;          // halo below: y = H .. H+r-1
        default: SMC_FAIL(SMC_ERR_INVALID, "which must be 0..3");
    }
    *dev = img + (size_t)first * row_bytes;
    *bytes = (size_t)r * row_bytes;
    return SMC_OK;
}

extern "C" uint64_t smc_denoiser_pairs(const smc_denoiser *d) {
    if (!d) return 0;
    return (uint64_t)(d->row_end - d->row_begin) * (uint64_t)d->W * (uint64_t)taps_in_window(d->radius) *
           (uint64_t)d->ptr_count;
}

extern "C" size_t smc_denoiser_record_bytes(const smc_denoiser *d) { return d ? d->rec_image_stride * d->rec_images : 0; }
extern "C" const char *smc_denoiser_kernel_name(const smc_denoiser *d) { return d ? d->kernel_name : ""; }

// ---------------------------------------------------------------------------------------------------------
// One-shot entry point with the reference's device-resident descriptor tables (cudaimgproc.hpp:756-777).
// ---------------------------------------------------------------------------------------------------------
extern "C" int smc_filter_device_tables(smc_context *ctx, int channels, int ptr_count, int width, int height,
                                        float ds_factor, int radius, int denoise_film, const void *n_ptrs,
                                        const void *mean_ptrs, const void *m2_ptrs, const void *m3_ptrs,
                                        const void *film_ptrs, const void *film_data, size_t film_step,
                                        const void *gbuf_ptrs, const void *gbuf_channel_counts,
                                        const void *gbuf_dr_factors, int n_gbufs, void *mean_corr_ptrs,
                                        void *disc_ptrs, void *film_filtered_ptrs, void *film_filtered_data,
                                        size_t film_filtered_step, void *stream) {
    return smc_filter_device_tables_host(ctx, channels, ptr_count, width, height, ds_factor, radius, denoise_film, n_ptrs,
                                         mean_ptrs, m2_ptrs, m3_ptrs, film_ptrs, film_data, film_step, gbuf_ptrs,
                                         gbuf_channel_counts, gbuf_dr_factors, n_gbufs, mean_corr_ptrs, disc_ptrs,
                                         film_filtered_ptrs, film_filtered_data, film_filtered_step, stream, nullptr, 0);
}

// The same call when (some of) the planes the tables point at are still on the host: `uploads` lists copies that have not been
// issued yet (Estimator::Upload deferred by the link shim).  Planes of `height` rows travel inside the row-chunked pipeline of
// smc_denoiser_run_host -- chunk k on the copy stream while chunk k-1 is prepassed and filtered -- so that PCIe overlaps the
// kernels although the reference calls Upload(); Denoise(); one after the other (statpath.cpp:406-418); anything else in the
// list is copied up front.
extern "C" int smc_filter_device_tables_host(smc_context *ctx, int channels, int ptr_count, int width, int height,
                                             float ds_factor, int radius, int denoise_film, const void *n_ptrs,
                                             const void *mean_ptrs, const void *m2_ptrs, const void *m3_ptrs,
                                             const void *film_ptrs, const void *film_data, size_t film_step,
                                             const void *gbuf_ptrs, const void *gbuf_channel_counts,
                                             const void *gbuf_dr_factors, int n_gbufs, void *mean_corr_ptrs,
                                             void *disc_ptrs, void *film_filtered_ptrs, void *film_filtered_data,
                                             size_t film_filtered_step, void *stream, const smc_host_rows *uploads,
                                             int n_uploads) {
    if (!ctx) SMC_FAIL(SMC_ERR_INVALID, "ctx == NULL");
    int rc = validate_common(channels, ptr_count, width, height, radius, n_gbufs);
    if (rc) return rc;
    if (!n_ptrs || !mean_ptrs || !m2_ptrs || !m3_ptrs || !film_ptrs || !film_filtered_ptrs)
        SMC_FAIL(SMC_ERR_INVALID, "NULL descriptor table");
    if (denoise_film && (!film_data || !film_filtered_data)) SMC_FAIL(SMC_ERR_INVALID, "denoiseFilm needs film buffers");
    if (n_gbufs > 0 && (!gbuf_ptrs || !gbuf_channel_counts || !gbuf_dr_factors))
        SMC_FAIL(SMC_ERR_INVALID, "NULL G-buffer table");
    if (n_uploads < 0 || (n_uploads > 0 && !uploads)) SMC_FAIL(SMC_ERR_INVALID, "bad upload list");
    SMC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<Xfer> ups;
    for (int i = 0; i < n_uploads; i++) {
        const smc_host_rows &u = uploads[i];
        if (!u.dev || !u.host || u.rows <= 0 || u.row_bytes == 0) SMC_FAIL(SMC_ERR_INVALID, "upload %d incomplete", i);
        if (u.rows == height) {
            ups.push_back(Xfer{(unsigned char *)u.dev, u.dev_step, smc_plane{(void *)u.host, u.host_step ? u.host_step : u.row_bytes},
                               u.row_bytes});
        } else {
            SMC_CUDA(cudaMemcpy2DAsync(u.dev, u.dev_step, u.host, u.host_step ? u.host_step : u.row_bytes, u.row_bytes, u.rows,
                                       cudaMemcpyHostToDevice, s));
        }
    }

    // The kernel variant and the record packing depend on the G-buffer channel counts and range factors, which live on the
    // device.  They are read back on EVERY call (a few bytes; the reference's flow blocks in Synchronize() right after
    // Denoise anyway, statpath.cpp:406-418) and are part of the plan's cache key, so a second Estimator whose tables land on
    // recycled addresses with other contents can never run with a stale plan.
    std::vector<unsigned char> gch(std::max(n_gbufs, 1));
    std::vector<float> gf(std::max(n_gbufs, 1));
    if (n_gbufs > 0) {
        SMC_CUDA(cudaMemcpyAsync(gch.data(), gbuf_channel_counts, n_gbufs, cudaMemcpyDeviceToHost, s));
        SMC_CUDA(cudaMemcpyAsync(gf.data(), gbuf_dr_factors, sizeof(float) * n_gbufs, cudaMemcpyDeviceToHost, s));
        SMC_CUDA(cudaStreamSynchronize(s));
    }
    std::vector<unsigned char> key(sizeof(int) * 8 + sizeof(float) + (size_t)n_gbufs * (1 + sizeof(float)));
    {
        unsigned char *k = key.data();
        const int ints[8] = {channels, ptr_count, width, height, radius, denoise_film, n_gbufs, 0};
        std::memcpy(k, ints, sizeof(ints)); k += sizeof(ints);
        std::memcpy(k, &ds_factor, sizeof(float)); k += sizeof(float);
        if (n_gbufs > 0) {
            std::memcpy(k, gch.data(), n_gbufs); k += n_gbufs;
            std::memcpy(k, gf.data(), sizeof(float) * n_gbufs);
        }
    }
    const int slot = channels == 3 ? 1 : 0;
    smc_denoiser *d = ctx->cached[slot];
    if (!d || key != ctx->cached_key[slot]) {
        if (d) smc_denoiser_destroy(d);
        ctx->cached[slot] = nullptr;
        int ng = 0;
        for (int g = 0; g < n_gbufs; g++) {
            if (gch[g] != 1 && gch[g] != 3) continue;  // ignored, as in dr2 (stat_denoiser.cu:103-110)
            if (!(gf[g] <= 0.f)) SMC_FAIL(SMC_ERR_INVALID, "G-buffer %d: drFactor %g must be <= 0", g, gf[g]);
            ng += gch[g];
        }
        if (ng > SMC_MAX_GBUF_CHANNELS) SMC_FAIL(SMC_ERR_UNSUPPORTED, "%d flattened G-buffer channels", ng);
        d = new (std::nothrow) smc_denoiser;
        if (!d) SMC_FAIL(SMC_ERR_NOMEM, "out of host memory");
        d->ctx = ctx; d->C = channels; d->ptr_count = ptr_count; d->W = width; d->H = height; d->radius = radius;
        d->denoise_film = denoise_film ? 1 : 0; d->mode = SMC_MEMBER_WELCH; d->n_gbufs = n_gbufs;
        d->NG = std::min(ng, SMC_REC_GBUF_CHANNELS); d->NGX = ng - d->NG;
        d->row_begin = 0; d->row_end = height; d->ds_factor = ds_factor; d->tables_external = true;
        if (cudaMalloc(&d->d_gch, std::max(n_gbufs, 1)) != cudaSuccess ||
            cudaMalloc(&d->d_gf, sizeof(float) * std::max(n_gbufs, 1)) != cudaSuccess) {
            smc_denoiser_destroy(d);
            SMC_FAIL(SMC_ERR_NOMEM, "cudaMalloc failed");
        }
        if (n_gbufs > 0) {
            d->h_gch.assign(gch.begin(), gch.begin() + n_gbufs);
            d->h_gf.assign(gf.begin(), gf.begin() + n_gbufs);
            cudaMemcpy(d->d_gch, gch.data(), n_gbufs, cudaMemcpyHostToDevice);
            cudaMemcpy(d->d_gf, gf.data(), sizeof(float) * n_gbufs, cudaMemcpyHostToDevice);
        }
        if ((rc = alloc_records(d)) || (rc = build_spatial_table(d)) || (rc = build_sym_table(d)) ||
            (rc = select_kernel(d))) {
            smc_denoiser_destroy(d);
            return rc;
        }
        SMC_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->cached[slot] = d;
        ctx->cached_key[slot] = key;
    }
    d->t_n = (const SmcPtrStepSz *)n_ptrs; d->t_mean = (const SmcPtrStepSz *)mean_ptrs;
    d->t_m2 = (const SmcPtrStepSz *)m2_ptrs; d->t_m3 = (const SmcPtrStepSz *)m3_ptrs;
    d->t_film = (const SmcPtrStepSz *)film_ptrs; d->t_gbufs = (const SmcPtrStepSz *)gbuf_ptrs;
    d->t_mc = (const SmcPtrStepSz *)mean_corr_ptrs; d->t_disc = (const SmcPtrStepSz *)disc_ptrs;
    d->t_out = (const SmcPtrStepSz *)film_filtered_ptrs; d->t_acc = nullptr;
    d->film = SmcPtrStepSz{(unsigned char *)film_data, film_step, width, height};
    d->film_filtered = SmcPtrStepSz{(unsigned char *)film_filtered_data, film_filtered_step, width, height};
    // run on the caller's stream
    cudaStream_t saved = ctx->stream;
    ctx->stream = s;
    rc = ups.empty() ? smc_denoiser_run(d) : run_pipeline(d, ups, std::vector<Xfer>(), 0, 0);
    ctx->stream = saved;
    return rc;
}
