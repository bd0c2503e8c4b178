/*
 * statmc_b200.h -- C ABI of libstatmc_b200.so: the B200-native (sm_100a) replacement for the data-parallel hot
 * path of cg-tuwien/StatMC: (1) per-pixel streaming moment accumulation and (2) the statistical denoiser.
 *
 * This layer replaces, for that path only, the cv::cuda::GpuMat buffer management and the
 * cv::cuda::stat_denoiser::* kernels the reference's src/statistics/ calls.  Reference interfaces are cited
 * per entry point as file:line relative to the reference checkout:
 *   EST.h   = src/statistics/estimator.h            EST.cpp = src/statistics/estimator.cpp
 *   BUF.h   = src/statistics/buffer.h
 *   CIP.hpp = src/ext/opencv_contrib/modules/cudaimgproc/include/opencv2/cudaimgproc.hpp
 *   SD.cu   = src/ext/opencv_contrib/modules/cudaimgproc/src/cuda/stat_denoiser.cu
 *
 * Conventions
 *   - plain C types only; every function returns an int status (SMC_OK == 0) unless noted; nothing throws.
 *     smc_last_error() returns a thread-local description of the last failure.
 *   - device work is asynchronous on the context's CUDA stream; smc_synchronize() is the only blocking call
 *     (mirrors Estimator::Upload/Denoise/Download/Synchronize, EST.cpp:409-425, 571-573).
 *   - image planes use the reference's layout: row-major, interleaved channels (CV_32FC3 = 12 B/px,
 *     CV_32FC1, CV_32SC1) with an arbitrary byte pitch; a plane descriptor is {device pointer, pitch} here and
 *     {data, step, cols, rows} (= cv::cuda::PtrStepSzb, 24 bytes) in the device-table entry points.
 *   - there is NO CPU fallback: every compute entry point fails with SMC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef STATMC_B200_H
#define STATMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMC_VERSION 100 /* 0.1.0 */

enum smc_status {
    SMC_OK = 0,
    SMC_ERR_INVALID = 1,     /* bad argument (the reference would hit UB or cv::error) */
    SMC_ERR_CUDA = 2,        /* CUDA runtime / launch failure, or no usable device */
    SMC_ERR_NOMEM = 3,       /* device or pinned-host allocation failed */
    SMC_ERR_UNSUPPORTED = 4  /* valid in the reference but outside this build's limits */
};

enum smc_dtype { SMC_F32 = 0, SMC_I32 = 1 };

/* membership function of the filter: SD.cu:40-42 MEMFNC (compile-time switch there, run-time here) */
enum smc_membership {
    SMC_MEMBER_WELCH = 0, /* MEMFNC 0: Johnson-corrected Welch test via discriminators, SD.cu:81-88 */
    SMC_MEMBER_MOON = 1   /* MEMFNC 1: Moon et al. 2013 CI test on raw means, SD.cu:125-144 */
};

/* Reference limits kept: width,height <= 65535 (unsigned short, SD.cu:18-20), radius <= 255 (unsigned char,
 * SD.cu:214), ptr_count <= 65535 (grid.z, SD.cu:422). */
#define SMC_MAX_DIM 65535
#define SMC_MAX_RADIUS 255
#define SMC_MAX_GBUF_CHANNELS 23 /* flattened G-buffer channels handled by this build (reference: unbounded).  Up to 7 travel
                                  * inside the packed per-pixel record (every kernel); channels 8..23 -- e.g. filterbuffers
                                  * [materialid depth normal albedo] = 8 -- go through a side array and the generic kernel.
                                  * G-buffers whose channel count is neither 1 nor 3 are ignored, as in dr2 (SD.cu:101-111) */
#define SMC_T_LUT_ENTRIES 1024  /* SD.cu:43 */

typedef struct smc_context smc_context;   /* one per (process, GPU): device, stream, t-quantile table */
typedef struct smc_buffer smc_buffer;     /* one device plane (+ optional pinned staging) */
typedef struct smc_denoiser smc_denoiser; /* a denoise plan: descriptor tables + packed record storage */

/* {device pointer of row 0, byte pitch}.  dev == NULL means "absent" where the field is optional. */
typedef struct smc_plane {
    void *dev;
    size_t step;
} smc_plane;

const char *smc_last_error(void);
int smc_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Context.  Replaces cv::cuda::stat_denoiser::setup() (CIP.hpp:739, SD.cu:352-355) and the Estimator's
 * cv::cuda::Stream member (EST.h:326).
 * --------------------------------------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device` with its own non-blocking stream. */
int smc_context_create(int device, smc_context **out);
/* Same, but all work is issued on the caller's stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream);
 * the stream is borrowed, not owned. */
int smc_context_create_on_stream(int device, void *cuda_stream, smc_context **out);
void smc_context_destroy(smc_context *ctx);
/* cv::cuda::stat_denoiser::synchronize(Stream&) (CIP.hpp:741-743, SD.cu:357-359); also surfaces asynchronous
 * kernel errors, which the reference never checks. */
int smc_synchronize(smc_context *ctx);
void *smc_context_stream(smc_context *ctx); /* the cudaStream_t in use */
int smc_context_device(smc_context *ctx);
/* Number of kernels of this library launched on the context since creation (bench.py's gpu_launches). */
uint64_t smc_context_launch_count(smc_context *ctx);
/* Diagnostic: (pixel, sample) updates that smc_accumulate redid on its scalar IEEE path since the last call (inputs outside
 * the proven range of the fast path: zero/denormal-scale differences, non-finite values, n >= 2^22).  Synchronises the
 * context's stream and resets the counter.  Results are bit-identical on either path; this only explains throughput. */
uint64_t smc_accumulate_fallback_samples(smc_context *ctx);

/* Two-sided significance level of the Student-t quantile table (SD.cu:53-67: "SET DIFFERENT LUTS HERE").
 * Default 0.005 (t_005_quantiles, SD.cu:56,67).  The nine levels the reference carries as text are served from
 * bit-identical built-in tables; any other alpha in (0,1) is computed in double precision on the host
 * (entry i = float(t_quantile(1 - alpha/2, df = i + 1))). */
int smc_set_alpha(smc_context *ctx, double alpha);
double smc_get_alpha(smc_context *ctx);
/* Copies the 1024-entry float32 table currently in use to `out`. */
int smc_get_t_table(smc_context *ctx, float *out);
/* Host-side double-precision Student-t quantile and CDF used to build tables (exposed for tests). */
double smc_t_quantile(double p, double df);
double smc_t_cdf(double t, double df);
/* Device Student-t CDF, float32, evaluated for `count` (t, df) pairs resident on the device -- the "fast,
 * accuracy-checked Student-t CDF" of BASELINE.json; used to validate tables (cdf(table[i], i+1) == 1 - alpha/2)
 * and by the optional soft-membership extension.  t, df, out: device pointers. */
int smc_student_t_cdf(smc_context *ctx, const float *t, const float *df, float *out, size_t count);

/* ---------------------------------------------------------------------------------------------------------
 * Buffers.  Replace cv::cuda::GpuMat(rows, cols, type) + Buffer::upload/download (BUF.h:24-63) and
 * Estimator::Upload/Download (EST.cpp:409-425).
 * --------------------------------------------------------------------------------------------------------- */
/* Allocates a rows x cols plane of `channels` interleaved elements of `dtype`; rows are pitched to 256 bytes
 * when rows > 1 && cols > 1 (GpuMat uses cudaMallocPitch in that case, gpu_mat.cu:112-123), contiguous otherwise.
 * The plane is zero-filled. */
int smc_buffer_create(smc_context *ctx, int rows, int cols, int channels, int dtype, smc_buffer **out);
void smc_buffer_destroy(smc_buffer *buf);
void *smc_buffer_dev(const smc_buffer *buf);
size_t smc_buffer_step(const smc_buffer *buf);
smc_plane smc_buffer_plane(const smc_buffer *buf);
/* Async 2-D copies on the context stream (GpuMat::upload/download = cudaMemcpy2DAsync, gpu_mat.cu:224-234).
 * `host_step` is the host row pitch in bytes (0 = tightly packed).  Truly asynchronous only for pinned memory. */
int smc_buffer_upload(smc_buffer *buf, const void *host, size_t host_step);
int smc_buffer_download(const smc_buffer *buf, void *host, size_t host_step);
/* Row-range variants (row-band pipelining and multi-GPU sharding): rows [row0, row0+nrows) of the device plane
 * <-> `host`, which points at the first of those rows. */
int smc_buffer_upload_rows(smc_buffer *buf, int row0, int nrows, const void *host, size_t host_step);
int smc_buffer_download_rows(const smc_buffer *buf, int row0, int nrows, void *host, size_t host_step);
int smc_buffer_fill_zero(smc_buffer *buf);
/* Stream-ordered device-to-device copy of `bytes` bytes (same or peer device; used for the record-halo exchange
 * when the ranks of a box live in one process). */
int smc_memcpy_device(smc_context *ctx, void *dst, const void *src, size_t bytes);
/* Pinned host memory (the reference uploads from pageable cv::Mat memory; pinned makes the copies overlap). */
int smc_host_alloc(size_t bytes, void **out);
void smc_host_free(void *p);
int smc_host_register(void *p, size_t bytes);
int smc_host_unregister(void *p);

/* ---------------------------------------------------------------------------------------------------------
 * Stage 1: moment accumulation.  Replaces StatTile<T>::Add[Transform]SampleM{1,2,3} (EST.h:162-232) applied
 * to a batch of samples, and Estimator::Merge[Transform]Tile (EST.cpp:341-388), whose only job is to copy the
 * running totals into the planes -- here the planes ARE the running state, resident in HBM.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct smc_moments {
    int width, height, channels; /* channels: 1 (Float) or 3 (Vec3) */
    smc_plane n;                 /* CV_32SC1; the reference's `n` plane (EST.cpp:121,347) */
    smc_plane mean, m2, m3;      /* CV_32FC(channels) */
    smc_plane film_mean, film_m2;/* CV_32FC(channels); may alias mean/m2 when !transform (EST.cpp:128-136) */
} smc_moments;

/* Applies `nsamples` samples per pixel, in order, to the state (sequential continuation: exactly the reference's
 * arithmetic and order).
 *   samples    : device pointer, [nsamples][height][width][channels] float32, tightly packed
 *   transform  : 1 = AddTransformSample (Box-Cox lambda=.5 + raw film moments, EST.h:212-226), 0 = AddSample (:206-211)
 *   max_moment : 1, 2 or 3 (AddStatSampleM1/M2/M3, EST.h:162-205)
 *   row_begin,row_end: rows of the planes to update (0,0 = all); samples cover exactly those rows. */
int smc_accumulate(smc_context *ctx, const smc_moments *state, const float *samples, int nsamples, int transform,
                   int max_moment, int row_begin, int row_end);
/* NEW capability (no reference counterpart): pairwise Chan/Pebay combination
 * dst <- dst (+) src of two independently accumulated moment sets (n, mean, M2, M3 and film mean/M2). */
int smc_merge_moments(smc_context *ctx, const smc_moments *dst, const smc_moments *src);
/* calculate_mean_vars_kernel (SD.cu:148-159) / cv::cuda::stat_denoiser::calculateMeanVars<T> (CIP.hpp:745-754):
 * out = m2 / (n (n-1)) per pixel.  (The CPU loop the reference ships instead, EST.cpp:524-568, reads n once per row and
 * scales RGB planes by the reciprocal -- OpenCV's Vec3f / float -- so it is up to 1 ulp away; not replicated.) */
int smc_calculate_mean_vars(smc_context *ctx, int width, int height, int channels, smc_plane n, smc_plane m2,
                            smc_plane out);

/* The reference's own argument list (CIP.hpp:745-754): every *_ptrs argument is a DEVICE-resident array of ptr_count
 * 24-byte {data, step, cols, rows} descriptors (cv::cuda::PtrStepSzb); `stream` is a cudaStream_t. */
int smc_calculate_mean_vars_device_tables(smc_context *ctx, int channels, int ptr_count, int width, int height,
                                          const void *n_ptrs, const void *m2_ptrs, void *mean_var_ptrs, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Stage 2: statistical denoiser.  Replaces cv::cuda::stat_denoiser::filter<T> (CIP.hpp:756-799; SD.cu:397-475:
 * johnson_mean_corrs_kernel -> mean_discriminators_kernel -> filter_kernel) and Estimator::Denoise (EST.cpp:427-489).
 *
 * Non-finite input: a NaN / +-Inf value in a filtered plane (film, film_ptrs[z]) reaches the output pixels whose reference
 * loop adds it -- the centres that accept the pixel as a tap inside the disc, and the pixel itself -- and no others, as the
 * `continue`s of SD.cu:247-268 have it (up to 16384 such values per frame and plan; beyond that the excess ones turn their
 * whole window NaN).  Non-finite statistics only make the membership test fail, as in the reference.  With several scalar
 * images and no denoise_film, three images share one internal record (same results as one by one).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct smc_filter_desc {
    int channels;      /* 1 = filter<float>, 3 = filter<float3> */
    int ptr_count;     /* images per launch sharing the G-buffers (grid.z, SD.cu:422) */
    int width, height; /* plane size in pixels (for a row band: the rows present locally, halos included) */
    float ds_factor;   /* -0.5 / filterSD^2 (EST.h:259) */
    int radius;        /* window [c-r, c+r) x [c-r, c+r) clipped to dS2 <= r^2 (SD.cu:24-36, 247-248) */
    int denoise_film;  /* denoiseFilm: for image 0, filter `film` into `film_filtered` (SD.cu:251-253, 319-344) */
    int membership;    /* enum smc_membership */
    /* per-image planes, arrays of ptr_count entries (host memory; copied at plan creation) */
    const smc_plane *n, *mean, *m2, *m3; /* CV_32SC1, CV_32FC(channels) x3 */
    const smc_plane *film_ptrs;          /* value to average per image: CV_32FC(channels) ("film-mean") */
    smc_plane film;                      /* CV_32FC3 pbrt film image; required iff denoise_film */
    /* G-buffers shared by all images */
    int n_gbufs;
    const smc_plane *gbufs;         /* n_gbufs planes, CV_32FC1 or CV_32FC3 */
    const uint8_t *gbuf_channels;   /* 1 or 3 each (SD.cu:90-112) */
    const float *gbuf_dr_factors;   /* -0.5 / sd_g^2 each (EST.cpp:16), must be <= 0 */
    /* outputs */
    const smc_plane *mean_corr, *disc;   /* optional (array may be NULL, entries may have dev == NULL): the
                                            API-visible "mean-corr"/"discriminator" planes (SD.cu:181, 205) */
    const smc_plane *film_filtered_ptrs; /* ptr_count planes CV_32FC(channels) ("film-mean-f") */
    smc_plane film_filtered;             /* CV_32FC3 "film-f"; required iff denoise_film */
    const smc_plane *accepted;           /* optional debug output, ptr_count CV_32SC1 planes: taps accepted per pixel */
    /* Row-band sharding.  Output rows [row_begin, row_end) are produced (0,0 = all rows).  Reads clamp to
     * [0, height-1] like BrdReplicate, so a band that carries >= radius halo rows above and >= radius-1 below
     * (or sits on a true image border) produces exactly the rows of the unsharded run. */
    int row_begin, row_end;
    /* Record-halo exchange mode (multi-GPU, statistics resident per band, no raw halo rows): when a flag is
     * set, the `radius` record rows above row 0 / below row height-1 are NOT replicated from the edge row by the
     * prepass; the caller fills them from the neighbouring rank via smc_denoiser_halo() before smc_denoiser_filter(). */
    int halo_top_external, halo_bottom_external;
    int kernel; /* 0 = auto (symmetric kernel where it applies -- RGB statistics, Welch membership, radius >= 2 --, else the
                 * one-sided streaming kernel, else generic), 1 = force the generic (any-radius, any-config) kernel,
                 * 2 = force the one-sided streaming kernel, 3 = force the symmetric kernel.  The symmetric kernel evaluates every
                 * unordered pair once; its accept/reject decisions are identical, its sums differ in summation order only */
} smc_filter_desc;

/* Builds the plan: validates, uploads descriptor tables, allocates the packed per-pixel record array
 * ((height + 2 radius) x (width + 2 pad) x 64 B per image). */
int smc_denoiser_create(smc_context *ctx, const smc_filter_desc *desc, smc_denoiser **out);
void smc_denoiser_destroy(smc_denoiser *d);
/* prepass (Johnson correction + discriminator + record packing; SD.cu:162-206) over all local rows */
int smc_denoiser_prepass(smc_denoiser *d);
/* filter (SD.cu:208-345) over rows [row_begin, row_end) */
int smc_denoiser_filter(smc_denoiser *d);
/* prepass + filter: the whole of cv::cuda::stat_denoiser::filter<T> */
int smc_denoiser_run(smc_denoiser *d);
/* Row-range forms (row-band pipelining; overlapping the halo exchange with interior rows).  Prepass rows are plane
 * rows [row_begin, row_end) of [0, height); the replicated border rows are produced together with row 0 / row height-1.
 * Filter rows must lie inside the plan's [row_begin, row_end) and need the prepass of rows [y - radius, y + radius). */
int smc_denoiser_prepass_rows(smc_denoiser *d, int row_begin, int row_end);
int smc_denoiser_filter_rows(smc_denoiser *d, int row_begin, int row_end);

/* Host mirrors of the planes a plan was created with: {dev = HOST pointer of row 0, step = host pitch in bytes, 0 =
 * tightly packed}.  Arrays have the same length and order as in smc_filter_desc; NULL arrays / NULL entries are skipped. */
typedef struct smc_host_io {
    const smc_plane *n, *mean, *m2, *m3, *film_ptrs; /* uploaded (Estimator::uploadBuffers, EST.cpp:163-178) */
    smc_plane film;                                  /* uploaded */
    const smc_plane *gbufs;                          /* uploaded */
    const smc_plane *film_filtered_ptrs;             /* downloaded (Estimator::downloadBuffers) */
    smc_plane film_filtered;                         /* downloaded */
    const smc_plane *mean_corr, *disc;               /* optional downloads */
} smc_host_io;
/* Estimator::Upload -> Denoise -> Download (EST.cpp:409-489) as ONE pipelined call: host planes are uploaded in row
 * chunks on a copy stream while earlier chunks are prepassed and filtered on the context stream and finished output rows
 * are downloaded on a third stream, so PCIe transfers overlap the kernels.  Results are identical to upload-all, run,
 * download-all.  Asynchronous: the context stream joins the copy streams, smc_synchronize() completes everything.
 * Host memory should be pinned (smc_host_alloc / smc_host_register); pageable memory works but does not overlap.
 * chunk_rows = 0 picks whole waves of the filter grid (about 8 chunks per frame). */
int smc_denoiser_run_host(smc_denoiser *d, const smc_host_io *io, int chunk_rows);
/* Record rows for halo exchange, image z.  which: 0 = own top `radius` rows (send up), 1 = own bottom `radius`
 * rows (send down), 2 = halo above row 0 (receive from the rank above), 3 = halo below the last row (receive from
 * the rank below).  Each region is one contiguous block of *bytes on the device. */
int smc_denoiser_halo(smc_denoiser *d, int z, int which, void **dev, size_t *bytes);
/* Peer halos: the same exchange WITHOUT a separate copy step.  Each rank exports its record array once
 * (smc_denoiser_peer_export; the 128-byte smc_peer_info travels to the neighbours by any means, e.g. an all-gather) and
 * attaches the rank above (which = 0) and below (which = 1).  From then on smc_denoiser_prepass() stores the records of its
 * top / bottom `radius` rows straight into the neighbours' halo rows over NVLink (peer-mapped memory, CUDA IPC) from inside
 * the prepass kernel, and prepass / filter order themselves across GPUs with release/acquire flags in device memory:
 * no NCCL call, no host synchronisation, nothing to do between prepass and filter.  The plan must have been created with
 * halo_top_external / halo_bottom_external for the attached sides; every rank must call prepass and filter once per step.
 * smc_denoiser_peer_attach_local is the same for two plans living in one process (same or peer-accessible devices). */
typedef struct smc_peer_info {
    unsigned char ipc_handle[64]; /* cudaIpcMemHandle_t of the record array */
    uint64_t image_stride, flags_offset;
    int32_t height, radius, rec_pitch, ptr_count, device;
    int32_t reserved[7];
} smc_peer_info;
int smc_denoiser_peer_export(smc_denoiser *d, smc_peer_info *out);
int smc_denoiser_peer_attach(smc_denoiser *d, int which, const smc_peer_info *info);
int smc_denoiser_peer_attach_local(smc_denoiser *d, int which, smc_denoiser *other);
/* Algorithmic work of one filter pass: pair evaluations = rows x width x taps(radius) x ptr_count. */
uint64_t smc_denoiser_pairs(const smc_denoiser *d);
size_t smc_denoiser_record_bytes(const smc_denoiser *d);
/* name of the filter kernel variant the plan selected ("stream<3,6,4>", "generic") -- for reports */
const char *smc_denoiser_kernel_name(const smc_denoiser *d);

/* One-shot form with the reference's own argument list: every *_ptrs argument is a DEVICE-resident array of
 * 24-byte {data, step, cols, rows} descriptors (cv::cuda::PtrStepSzb), gbuf_channel_counts a device uchar array,
 * gbuf_dr_factors a device float array, exactly what Estimator::AllocateBuffers uploads (EST.cpp:35-84, 271-288)
 * and what filter<T> receives (CIP.hpp:756-777).  The tables are dereferenced on the device; no host sync.
 * `stream` is a cudaStream_t.  Scratch (the record array) is cached on the context between calls. */
int smc_filter_device_tables(smc_context *ctx, int channels, int ptr_count, int width, int height, float ds_factor,
                             int radius, int denoise_film, const void *n_ptrs, const void *mean_ptrs,
                             const void *m2_ptrs, const void *m3_ptrs, const void *film_ptrs, const void *film_data,
                             size_t film_step, const void *gbuf_ptrs, const void *gbuf_channel_counts,
                             const void *gbuf_dr_factors, int n_gbufs, void *mean_corr_ptrs, void *disc_ptrs,
                             void *film_filtered_ptrs, void *film_filtered_data, size_t film_filtered_step,
                             void *stream);

/* The same call while (some of) the planes behind the tables are still on the host: `uploads` lists host -> device copies
 * that have NOT been issued yet (the link shim defers Estimator::Upload's GpuMat::upload calls, EST.cpp:409-432).  Planes of
 * `height` rows travel inside the row-chunked pipeline of smc_denoiser_run_host, so that PCIe overlaps the kernels although the
 * reference calls Upload(); Denoise(); Download(); one after the other (statpath.cpp:406-418); other entries are copied up
 * front.  Host memory should be page-locked (smc_host_alloc) and must stay unchanged until the stream has been synchronised.
 * n_uploads == 0 is smc_filter_device_tables. */
typedef struct smc_host_rows {
    void *dev;         /* destination plane */
    size_t dev_step;   /* bytes per device row */
    const void *host;  /* source rows */
    size_t host_step;  /* bytes per host row (0: tightly packed) */
    size_t row_bytes;  /* bytes to copy per row */
    int rows;
} smc_host_rows;
int smc_filter_device_tables_host(smc_context *ctx, int channels, int ptr_count, int width, int height, float ds_factor,
                                  int radius, int denoise_film, const void *n_ptrs, const void *mean_ptrs,
                                  const void *m2_ptrs, const void *m3_ptrs, const void *film_ptrs, const void *film_data,
                                  size_t film_step, const void *gbuf_ptrs, const void *gbuf_channel_counts,
                                  const void *gbuf_dr_factors, int n_gbufs, void *mean_corr_ptrs, void *disc_ptrs,
                                  void *film_filtered_ptrs, void *film_filtered_data, size_t film_filtered_step,
                                  void *stream, const smc_host_rows *uploads, int n_uploads);

#ifdef __cplusplus
}
#endif
#endif /* STATMC_B200_H */
