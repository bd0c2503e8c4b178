// smc_internal.h -- shared declarations of libstatmc_b200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/statmc_b200.h"

// ---- error plumbing -------------------------------------------------------------------------------------
void smc_set_error(const char *fmt, ...);

#define SMC_FAIL(code, ...)         \
    do {                            \
        smc_set_error(__VA_ARGS__); \
        return (code);              \
    } while (0)

#define SMC_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            smc_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);      \
            return e__ == cudaErrorMemoryAllocation ? SMC_ERR_NOMEM : SMC_ERR_CUDA;                          \
        }                                                                                                    \
    } while (0)

#define SMC_CHECK_LAUNCH(ctx)                                                                                \
    do {                                                                                                     \
        (ctx)->launches++;                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                                \
        if (e__ != cudaSuccess) {                                                                            \
            smc_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__);  \
            return SMC_ERR_CUDA;                                                                             \
        }                                                                                                    \
    } while (0)

// ---- plane descriptor as it sits in device tables: layout-compatible with cv::cuda::PtrStepSzb ----------
// (src/ext/opencv/modules/core/include/opencv2/core/cuda_types.hpp:103-135: {T* data; size_t step; int cols; int rows;})
struct SmcPtrStepSz {
    unsigned char *data;
    size_t step;
    int cols;
    int rows;
};
static_assert(sizeof(SmcPtrStepSz) == 24, "must match cv::cuda::PtrStepSzb");

// ---- packed per-pixel record (64 B) -----------------------------------------------------------------------
// float slots:  0 m.x  1 m.y  2 d.x  3 d.y | 4 m.z  5 d.z  6 V.z  7 g6 (1.0f when NG < 7) | 8 V.x  9 V.y 10 g0 11 g1 | 12 g2 13 g3 14 g4 15 g5
//   m = Johnson-corrected mean (Welch mode) or raw mean (Moon mode); d = discriminator (Welch) or CI half-width (Moon)
//   V = value to average (film or film-mean); g* = G-buffer channels pre-scaled by sqrt(-drFactor * log2(e))
//   channels == 1: m -> slot 0, d -> slot 2, scalar value -> slot 4, film RGB (denoiseFilm, image 0) -> slots 8, 9, 6
// Records are stored two to a 144-byte "line": 2 x 64 B + 16 B of padding.  The odd line stride makes eight
// consecutive lines start in eight different 16-byte bank groups, so when the lanes of a warp read records that are
// two apart (the streaming filter's mapping) every LDS.128 is bank-conflict-free, and -- unlike an XOR swizzle --
// the chunk addresses of a record stay an affine function of its index: no per-load integer math in the inner loop.
// A record row is laid out identically in HBM and in shared memory, so a row segment moves with one bulk copy.
#define SMC_REC_GBUF_CHANNELS 7  // G-buffer channels inside the record; further ones live in the extension array
#define SMC_REC_FLOATS 16
#define SMC_REC_BYTES 64
#define SMC_LINE_BYTES 144
static const int kGSlot[7] = {10, 11, 12, 13, 14, 15, 7};

__host__ __device__ inline uint32_t smc_rec_offset(uint32_t pcol) {
    // byte offset, within a record row, of the record at padded column `pcol`
    return (pcol >> 1) * SMC_LINE_BYTES + (pcol & 1u) * SMC_REC_BYTES;
}
__host__ __device__ inline uint32_t smc_rec_chunk_offset(uint32_t pcol, uint32_t chunk) {
    return smc_rec_offset(pcol) + chunk * 16u;
}
// bytes of a record row of `rec_pitch` records (rec_pitch is even)
__host__ __device__ inline size_t smc_rec_row_bytes(int rec_pitch) { return (size_t)(rec_pitch / 2) * SMC_LINE_BYTES; }

struct smc_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    uint64_t launches = 0;
    double alpha = 0.005;
    float h_lut[SMC_T_LUT_ENTRIES];
    float *d_lut = nullptr;  // device copy of the active table
    unsigned long long *d_accum_fallback = nullptr;  // samples the accumulate kernel redid on its scalar IEEE path
    // scratch plan cached for smc_filter_device_tables
    // (one slot per element type: Estimator::Denoise calls filter<float> and filter<float3> back to back when both
    // groups are populated, estimator.cpp:434-488, and neither plan should evict the other)
    struct smc_denoiser *cached[2] = {nullptr, nullptr};
    std::vector<unsigned char> cached_key[2];
};

struct smc_buffer {
    smc_context *ctx;
    void *dev;
    size_t step;
    int rows, cols, channels, dtype;
    size_t elem_bytes;  // bytes per pixel
};

// Cross-GPU halo protocol folded into the kernels (multi-GPU peer halos; all null / zero otherwise).  A kernel first WAITS --
// one thread per block spins with acquire loads at system scope until both flags have reached `wait_value` -- and, when its
// last block retires (device-scope counter), SIGNALS: release stores of `signal_value` at system scope into the neighbours'
// flags.  Replaces four single-thread kernels per step.
struct SmcHaloSync {
    const int *wait0, *wait1;
    int wait_value;
    int *signal0, *signal1;
    int signal_value;
    int *done_counter;  // blocks retired so far; reset by the last one
};

__device__ __forceinline__ void smc_halo_wait(const SmcHaloSync &h) {  // call from ONE thread, then barrier
    for (int k = 0; k < 2; k++) {
        const int *f = k ? h.wait1 : h.wait0;
        if (!f) continue;
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v < h.wait_value) __nanosleep(100);
        } while (v < h.wait_value);
    }
}
// call from ONE thread of every participating block after the block's work (and a block barrier); `total` blocks take part
__device__ __forceinline__ void smc_halo_signal_last(const SmcHaloSync &h, int total) {
    if (!h.signal0 && !h.signal1) return;
    __threadfence_system();
    const int done = atomicAdd(h.done_counter, 1) + 1;
    if (done == total) {
        *h.done_counter = 0;
        __threadfence_system();
        if (h.signal0) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(h.signal0), "r"(h.signal_value) : "memory");
        if (h.signal1) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(h.signal1), "r"(h.signal_value) : "memory");
    }
}

// Non-finite values (a NaN or +-Inf radiance in the plane being filtered).  The reference SKIPS rejected taps and taps outside
// the disc (stat_denoiser.cu:247-268), so a non-finite value only reaches the centres it is a member tap of; the streaming
// kernels instead give such taps weight 0, and 0 * Inf = NaN would reach every centre of the window.  The prepass therefore
// stores 0 in the record for a non-finite value and lists the position; after the filter a fix-up kernel adds w * value to
// exactly the centres whose reference loop would have added it (smc_filter_fixup.cu).  The lists live behind the halo flags in
// the record allocation, two of them (step parity) so that a neighbouring rank can read the entries of its halo rows while
// this rank already lists the next frame.
#define SMC_NF_CAP 16384
struct SmcNfEntry {
    int pr, pc, z, mask;  // padded record row / column, record image, bit j set: value j is non-finite
    float v[4];           // values j = 0..2: record slots 8, 9, 6 (RGB value / film, or the three images of a triple);
                          // j = 3: slot 4, the scalar value of a one-image scalar record
};
struct SmcNfList {
    int count, pad[7];
    SmcNfEntry e[SMC_NF_CAP];
};
#define SMC_NF_OFFSET 256  // bytes from the halo flags to SmcNfList[2]
// what the fix-up kernel reads: this rank's list and, with peer halos, the neighbours' (entries of the rows that were stored
// into this rank's halo, shifted to this rank's padded rows)
struct SmcNfSources {
    const SmcNfList *list[3];
    int row_lo[3], row_hi[3], row_shift[3];  // take entries with row_lo <= pr < row_hi; this rank's padded row = pr + row_shift
};

// filter parameters shared by the kernels (passed by value)
struct SmcFilterParams {
    int W, H;              // local plane size
    int C;                 // 1 or 3
    int NG;                // flattened G-buffer channels (0..7)
    int radius;
    int mode;              // smc_membership
    int ptr_count;
    int denoise_film;
    int sm_count;          // SMs of the device (work partitioning heuristics)
    // three scalar images per record (symmetric kernel only): ptr_count counts record images = triples, `images` the plan's
    // images (out_ptrs has `images` entries, image 3z + k sits in channel k's slots of record image z)
    int tri, images;
    int row_begin, row_end;
    int padX;              // record columns left of x = 0
    int rec_pitch;         // records per record row
    size_t rec_image_stride;  // bytes between images
    const unsigned char *rec; // record array base (image 0, record row 0 = y = -radius)
    const float *sw;       // spatial table: (2r + 2*SW_MARGIN_Y) rows x sw_stride floats; -inf where invalid
    int sw_stride, sw_margin_y, sw_margin_x;
    const SmcPtrStepSz *out_ptrs;  // film_filtered_ptrs table (device)
    SmcPtrStepSz film_filtered;    // "film-f"
    const SmcPtrStepSz *accepted;  // optional, may be null
    // G-buffer channels beyond the seven of the record (generic kernel only): [record row][padded column][gext_stride] floats,
    // pre-scaled like the record's; shared by all images
    const float *gext;
    int NGX, gext_stride;
    SmcHaloSync halo;              // wait for the neighbours' halo records, release ours when done (sym / per-warp kernels)
    int *tile_counter;             // streaming kernel: global tile counter (dynamic scheduling), reset before each launch
    unsigned long long *trace;     // debug (SMC_STREAM_TRACE): per CTA {start ns, end ns, smid, tiles}; normally null
};

// parameters of the symmetric filter kernel (smc_filter_sym.cu), passed by value next to SmcFilterParams
struct SmcSymParams {
    int xorg;          // first centre column of strip 0 (even, <= -radius): strips cover [xorg, W + radius)
    int n_strips;
    int ystart;        // first centre row = row_begin - radius
    int n_trows;       // tile rows (two centre rows each) covering [ystart, row_end)
    int n_units_y;     // units per strip: n_big runs of u_big tiles, then runs of u_small tiles (the tail of the work queue)
    int n_big, u_big, u_small;
    int units_total;   // n_units_y * n_strips * ptr_count
    int scratch_rows;  // scratch rows per (image, strip): sum over units of (2 * tiles + radius)
    int seg_rec;       // records per streamed row segment (even) = entries per scratch row
    int slot_bytes, macc_bytes, warp_bytes, nwarps;
    int sw_stride, sw_rows, sw_my, sw_mx;  // forward spatial table: rows dy = -sw_my .. radius + sw_my
    float sw_special;  // table value of dS2 = radius^2: the offsets (0, r) and (r, 0), booked to the record only
    const float2 *sw;  // (-sw, 0) per forward offset; +inf outside the disc
    const int2 *rowrange;
    float4 *scratch;   // [image][strip][scratch_rows][seg_rec]: partial mirror sums (x, y, z, den)
    int *scratch_cnt;  // same indexing: partial accepted-tap counts (only with an `accepted` plane)
    float4 *fwd;       // [image][row_end - row_begin][W]: forward sums
    int *fwd_cnt;
    // three scalar images per record (copied from SmcFilterParams by smc_filter_sym_geometry): the scratch / forward entries
    // are (num0, num1, den0, den1) and the third image's (num2, den2) goes to scratch2 / fwd2
    int tri, images;
    float2 *scratch2, *fwd2;
    int *unit_counter;
};

struct SmcPrepassParams {
    int W, H, C, ptr_count, radius, mode, denoise_film;
    int triple, images;  // scalar statistics, three images per record: ptr_count counts triples, `images` the images
    int padX, rec_pitch;
    size_t rec_image_stride;
    unsigned char *rec;
    int skip_top, skip_bottom;  // halo rows filled externally
    int pr_begin, pr_end;       // padded record rows [pr_begin, pr_end) to produce (padded row pr holds y = pr - radius)
    // Peer halos (multi-GPU): when set, the records of the top / bottom `radius` own rows are ALSO stored straight into the
    // neighbouring rank's record array over NVLink (peer-mapped memory): into the bottom halo of the rank above
    // (its padded rows [H_up + r, H_up + 2r)) and the top halo of the rank below (its padded rows [0, r)).
    unsigned char *peer_up_halo, *peer_down_halo;   // address of the first halo row of image 0 in the peer's array, or null
    size_t peer_up_image_stride, peer_down_image_stride;
    // blocks that store into a neighbour wait until it has released its halo rows (wait0: rank above, wait1: rank below) and
    // the last of them tells both neighbours that their halos are complete
    SmcHaloSync halo;
    int halo_blocks;
    const SmcPtrStepSz *n, *mean, *m2, *m3, *film_ptrs;
    SmcPtrStepSz film;
    const SmcPtrStepSz *gbufs;           // G-buffer plane descriptors (device table)
    // flattened G-buffer channels (by value): channel k of the record comes from plane g_buf[k], component g_ch[k] of
    // g_nch[k], multiplied by g_scale[k] = sqrtf(-drFactor * log2(e))
    int NG;                       // channels packed into the record (<= 7)
    int NGX, gext_stride;         // further channels, written to `gext` (see SmcFilterParams)
    float *gext;
    unsigned char g_buf[SMC_MAX_GBUF_CHANNELS], g_ch[SMC_MAX_GBUF_CHANNELS], g_nch[SMC_MAX_GBUF_CHANNELS];
    float g_scale[SMC_MAX_GBUF_CHANNELS];
    const SmcPtrStepSz *mean_corr, *disc;  // optional tables (device) or null
    const float *lut;
    SmcNfList *nf;  // where non-finite values are listed (see SmcNfEntry)
};

int smc_launch_prepass(smc_context *ctx, const SmcPrepassParams &p);
int smc_launch_halo_signal(smc_context *ctx, int *f0, int *f1, int value);
int smc_launch_halo_wait(smc_context *ctx, const int *f0, const int *f1, int value);
int smc_launch_filter_generic(smc_context *ctx, const SmcFilterParams &p);
// adds the listed non-finite values to the outputs of rows [row_begin, row_end) (after any of the filter kernels)
int smc_launch_nonfinite_fixup(smc_context *ctx, const SmcFilterParams &p, const SmcNfSources &src);
// returns SMC_ERR_UNSUPPORTED when the configuration has no streaming instantiation.
// rowrange: device array, one {jlo, jhi} per spatial-table row; py: output rows per thread (2 or 4)
int smc_launch_filter_stream(smc_context *ctx, const SmcFilterParams &p, const int2 *rowrange, int py,
                             const char **name);
// workers (CTAs, or warps for the per-warp variant) of the persistent streaming grid that are resident at once, and
// the width in pixels of the tile one worker owns
int smc_filter_stream_resident_ctas(const SmcFilterParams &p, int py, int sm_count, int *tile_w);
bool smc_filter_stream_supported(const SmcFilterParams &p, int sm_count, const char **name);
// true when the streaming variant that would run handles SmcFilterParams::halo itself (the per-warp kernel)
bool smc_filter_stream_syncs_halo(const SmcFilterParams &p, int py);
// symmetric kernel: every unordered pair evaluated once (RGB statistics, Welch membership, any radius 2..255, any NG)
bool smc_filter_sym_supported(const SmcFilterParams &p);
bool smc_filter_sym_geometry(const SmcFilterParams &p, SmcSymParams &g, size_t &smem);
size_t smc_filter_sym_scratch_elems(const SmcFilterParams &p, const SmcSymParams &g);
int smc_launch_filter_sym(smc_context *ctx, const SmcFilterParams &p, const SmcSymParams &g, size_t smem, const char **name);
#define SMC_SW_MARGIN_Y 3
#define SMC_SW_MARGIN_X 2
