/*
 * statmc_oracle.c -- CPU restatement of StatMC's data-parallel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (statmc_b200/, the C-ABI library)
 * may include, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * Parity pin: the reference ships NO golden vectors, KATs or tests for this path
 * (SURVEY.md section 4), so this restatement is pinned two ways instead:
 *   (1) the t-quantile table it uses is asserted bit-equal (float32) to the reference's
 *       own table text (tests/golden/t_quantiles.json, made by tools/gen_t_quantiles.py);
 *   (2) on the GPU box it is compared against the reference's own CUDA kernels compiled
 *       UNMODIFIED from /root/reference into oracle/_ref/ (see oracle/Makefile,
 *       tests/test_reference_cuda.py).
 *
 * Reference files restated here (paths relative to /root/reference):
 *   EST.h  = src/statistics/estimator.h
 *   EST.cpp= src/statistics/estimator.cpp
 *   SD.cu  = src/ext/opencv_contrib/modules/cudaimgproc/src/cuda/stat_denoiser.cu
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).  -ffp-contract=off
 * matters: every float expression below is meant to round exactly where it is written;
 * the places where the reference's *CUDA* build fuses a multiply-add (nvcc -fmad=true default)
 * are written as explicit fmaf() and marked [nvcc-fma].
 *
 * Layouts follow the reference: planes are row-major, interleaved channels (CV_32FC3 = 12 B/px,
 * CV_32FC1, CV_32SC1), addressed with a byte pitch ("step") like cv::cuda::PtrStepSz
 * (src/ext/opencv/modules/core/include/opencv2/core/cuda_types.hpp:103-135).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SMO_LUT_SIZE 1024 /* SD.cu:43 T_QUANTILE_LUT_SIZE */
#define SMO_LUT_MAX 1023  /* SD.cu:44 T_QUANTILE_MAX_INDEX */

typedef struct smo_plane {
    void *data;  /* first byte of row 0 */
    size_t step; /* bytes between rows */
} smo_plane;

static inline float *rowf(const smo_plane *p, int y) { return (float *)((char *)p->data + (size_t)y * p->step); }
static inline int *rowi(const smo_plane *p, int y) { return (int *)((char *)p->data + (size_t)y * p->step); }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------
 * Stage 1: streaming moment accumulation.   EST.h:135-145 (boxCox), :162-226 (StatTile<T>).
 *
 * One call applies `nsamples` samples per pixel, in order, to persistent per-pixel state --
 * exactly what the reference does by keeping StatTile objects alive across iterations
 * (statpath.cpp:172-190) and calling Add[Transform]SampleM{1,2,3} once per sample
 * (statpath.cpp:355-371).  Per channel arithmetic; `n` is shared by the channels of a pixel.
 *
 *   samples  : [nsamples][npix][C] float32 (sample-major; the reference has no batch layout,
 *              its samples arrive one at a time from the path tracer)
 *   n        : [npix] int64 (reference: uint64_t n, stored to the int32 `n` plane by MergeTile,
 *              EST.cpp:347)
 *   mean,m2,m3,film_mean,film_m2 : [npix][C] float32
 *   transform: 1 = AddTransformSample (EST.h:212-226): Box-Cox(lambda=.5) for mean/m2/m3 and a second
 *              M2-level update of film_mean/film_m2 on the raw sample using the already-incremented n;
 *              0 = AddSample (EST.h:206-211): film_mean = mean, film_m2 = m2 after the update.
 *   max_moment: 1, 2 or 3 selects AddStatSampleM1/M2/M3 (EST.h:162-205).
 *   use_sqrt : 0 = powf(s, .5f) as the reference (EST.h:136,215); 1 = sqrtf(s), the form the CUDA kernel
 *              uses (differs from glibc powf by <= 1 ulp on rare inputs); lets tests separate that one
 *              deviation from everything else, which must be bit-exact.
 * ------------------------------------------------------------------------------------------ */
static inline float box_cox_half(float s, int use_sqrt) {
    /* EST.h:135-137: (std::pow(val, lambda) - 1.f) / lambda with lambda = .5f (EST.h:215) */
    const float p = use_sqrt ? sqrtf(s) : powf(s, .5f);
    return (p - 1.f) / .5f;
}

void smo_accumulate(int64_t npix, int C, int nsamples, const float *samples, int transform, int max_moment,
                    int use_sqrt, int64_t *n, float *mean, float *m2, float *m3, float *film_mean, float *film_m2) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npix; p++) {
        uint64_t np = (uint64_t)n[p];
        for (int s = 0; s < nsamples; s++) {
            const float *smp = samples + ((size_t)s * (size_t)npix + (size_t)p) * (size_t)C;
            np++; /* EST.h:168,181,196: n++ (once per sample, shared by channels: Vec3 ops are element-wise) */
            const float nf = (float)np; /* `d / n` with uint64 n: usual arithmetic conversion to float */
            for (int c = 0; c < C; c++) {
                const size_t i = (size_t)p * (size_t)C + (size_t)c;
                const float raw = smp[c];
                const float x = transform ? box_cox_half(raw, use_sqrt) : raw;
                const float d = x - mean[i];
                const float dN = d / nf;
                if (max_moment >= 3) {
                    const float d2 = d * d;
                    const float dN2 = dN * dN;
                    mean[i] += dN;
                    m2[i] += d * (d - dN);
                    /* EST.h:204: m3 += -3.f*dN*m2 + d*(d2 - dN2), with the m2 updated on the line above */
                    m3[i] += ((-3.f * dN) * m2[i]) + (d * (d2 - dN2));
                } else if (max_moment == 2) {
                    mean[i] += dN;
                    m2[i] += d * (d - dN);
                } else {
                    mean[i] += dN;
                }
                if (transform) {
                    /* EST.h:217-225, raw sample, n already incremented */
                    const float fD = raw - film_mean[i];
                    const float fDN = fD / nf;
                    film_mean[i] += fDN;
                    film_m2[i] += fD * (fD - fDN);
                } else {
                    film_mean[i] = mean[i]; /* EST.h:209-210 */
                    film_m2[i] = m2[i];
                }
            }
        }
        n[p] = (int64_t)np;
    }
}

/* float64 accumulation of the same sample stream (two-pass-free Welford in double), for information
 * only: it says how far the float32 streaming update is from the exact moments. Not a parity target. */
void smo_accumulate_f64(int64_t npix, int C, int nsamples, const float *samples, int transform, int64_t *n,
                        double *mean, double *m2, double *m3, double *film_mean, double *film_m2) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < npix; p++) {
        uint64_t np = (uint64_t)n[p];
        for (int s = 0; s < nsamples; s++) {
            const float *smp = samples + ((size_t)s * (size_t)npix + (size_t)p) * (size_t)C;
            np++;
            const double nf = (double)np;
            for (int c = 0; c < C; c++) {
                const size_t i = (size_t)p * (size_t)C + (size_t)c;
                const double raw = smp[c];
                const double x = transform ? (sqrt(raw) - 1.0) / 0.5 : raw;
                const double d = x - mean[i], dN = d / nf;
                mean[i] += dN;
                m2[i] += d * (d - dN);
                m3[i] += -3.0 * dN * m2[i] + d * (d * d - dN * dN);
                if (transform) {
                    const double fD = raw - film_mean[i], fDN = fD / nf;
                    film_mean[i] += fDN;
                    film_m2[i] += fD * (fD - fDN);
                } else {
                    film_mean[i] = mean[i];
                    film_m2[i] = m2[i];
                }
            }
        }
        n[p] = (int64_t)np;
    }
}

/* MergeTile / MergeTransformTile, EST.cpp:341-352, 374-388: the "merge" of the reference is a plain copy
 * of the running totals of one tile into the global planes at offset y*W + x (no pairwise combine).
 * tile_* are [th*tw][C] (tile-local, row-major), the planes are [H*W][C] with n stored as int32. */
void smo_merge_tile(int W, int C, int tx0, int ty0, int tw, int th, int transform, const int64_t *tile_n,
                    const float *tile_mean, const float *tile_m2, const float *tile_m3, const float *tile_film_mean,
                    const float *tile_film_m2, int *n, float *mean, float *m2, float *m3, float *film_mean,
                    float *film_m2) {
    for (int y = 0; y < th; y++)
        for (int x = 0; x < tw; x++) {
            const size_t t = (size_t)y * tw + x;
            const size_t o = (size_t)(ty0 + y) * W + (tx0 + x); /* EST.cpp:343 */
            n[o] = (int)tile_n[t];                             /* EST.cpp:347 */
            for (int c = 0; c < C; c++) {
                mean[o * C + c] = tile_mean[t * C + c];
                m2[o * C + c] = tile_m2[t * C + c];
                m3[o * C + c] = tile_m3[t * C + c];
                if (transform) { /* EST.cpp:385-386 */
                    film_mean[o * C + c] = tile_film_mean[t * C + c];
                    film_m2[o * C + c] = tile_film_m2[t * C + c];
                }
            }
        }
}

/* calculate_mean_vars: per-pixel-correct form of SD.cu:148-159: meanVar = m2 / (n*(n-1)).
 * per_row_n_bug = 1 reproduces the shipped CPU loop EST.cpp:524-568 instead, which reads n once per ROW
 * (`float nPF = (float)*nP;` outside the column loop) and divides by ((nPF-1)*nPF). */
void smo_calculate_mean_vars(int W, int H, int C, const smo_plane *n, const smo_plane *m2, const smo_plane *out,
                             int per_row_n_bug) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        const int *nr = rowi(n, y);
        const float *mr = rowf(m2, y);
        float *orow = rowf(out, y);
        for (int x = 0; x < W; x++) {
            const float nf = (float)(per_row_n_bug ? nr[0] : nr[x]);
            const float den = per_row_n_bug ? ((nf - 1.f) * nf) : (nf * (nf - 1.f));
            for (int c = 0; c < C; c++) orow[x * C + c] = mr[x * C + c] / den;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage 2a: prepass.   johnson_mean_corrs_kernel SD.cu:162-182 (+ johnson_mean_corr :114-123),
 *                      mean_discriminators_kernel SD.cu:184-206.
 * float32, operation order as written there; the reference's SASS has no fused multiply-add on these values.
 * n < 2 makes index 2n-3 negative, an out-of-bounds LUT read in the reference (UB); here t is then
 * taken as NaN-irrelevant: with n = 1 the reference's 0/0 makes disc NaN whatever t is, and for
 * n <= 0 we define t = lut[0] (documented in DESIGN.md; the CUDA path does the same).
 * ------------------------------------------------------------------------------------------ */
static inline float johnson_corr(float nF, float s2, float m3) {
    return s2 > FLT_EPSILON ? (m3 / nF) / (6.f * s2 * nF) : 0.f; /* SD.cu:115 */
}

static inline float lut_t(const float *lut, int idx) {
    /* SD.cu:200-202: t = lut[1023]; if (2n-3 < 1024) t = lut[2n-3];  (signed compare) */
    if (idx < 0) idx = 0; /* reference: OOB read; see header comment */
    return idx < SMO_LUT_SIZE ? lut[idx] : lut[SMO_LUT_MAX];
}

void smo_prepass(int W, int H, int C, const float *lut, const smo_plane *n, const smo_plane *mean,
                 const smo_plane *m2, const smo_plane *m3, const smo_plane *mean_corr, const smo_plane *disc) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        const int *nr = rowi(n, y);
        const float *me = rowf(mean, y), *s2r = rowf(m2, y), *s3r = rowf(m3, y);
        float *mc = rowf(mean_corr, y), *dr = rowf(disc, y);
        for (int x = 0; x < W; x++) {
            const int ni = nr[x];
            const float nF = (float)ni; /* __int2float_rn */
            const float t = lut_t(lut, 2 * ni - 3);
            for (int c = 0; c < C; c++) {
                const int i = x * C + c;
                const float s2 = s2r[i] / (nF - 1.f);                    /* SD.cu:179 */
                const float m = me[i] + johnson_corr(nF, s2, s3r[i]);    /* SD.cu:181 */
                mc[i] = m;
                /* SD.cu:205: mean*mean - t*t*m2/(nF*(nF-1.f)).  Checked in the SASS of the reference build
                 * (cuobjdump of oracle/_ref/libstatmc_ref.so, mean_discriminators_kernel): FMUL m*m ... FADD m2, -q,
                 * i.e. nvcc does NOT fuse this one (the IEEE division sits between the two), so plain ops here. */
                const float q = ((t * t) * s2r[i]) / (nF * (nF - 1.f));
                dr[i] = m * m - q;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage 2b: membership-gated cross-bilateral filter.   filter_kernel<T> SD.cu:208-274,
 *           filter_kernel<float3> SD.cu:276-345, helpers :78-112, macros :17-37.
 *
 * Quirks kept: half-open window [c-r, c+r); dS2 from UNCLAMPED offsets with
 * `dS2 > r*r -> skip`; every read of a tap uses clamp-to-edge coordinates (BrdReplicate); the centre
 * tap (by unclamped coordinates) always has weight 1; RGB membership is the AND of the 3 channels;
 * 3-channel G-buffers sum the three squared differences first, then multiply by the factor.
 *
 * Membership is decided in float32 exactly as the reference does (it is a hard threshold on float32
 * planes).  Weights and sums are evaluated in `real` = float (mimicking the CUDA build's contractions)
 * or double (the "double-precision transcription" BASELINE.json asks for).
 *
 * value/out planes: what is averaged and where it goes.  The caller resolves the reference's routing:
 *   RGB kernel, denoiseFilm && z==0 : value = film,        out = filmFiltered   (SD.cu:319-344)
 *   otherwise                       : value = filmPtrs[z], out = filmFilteredPtrs[z]
 *   scalar kernel with denoiseFilm && z==0 additionally filters `film` (3ch) with the scalar gate;
 *   call this function twice (value_channels = 1, then 3) for that case (SD.cu:251-273).
 * accepted (optional): int32 plane receiving the number of taps that contributed (centre included) --
 *   used by tests to demand identical membership decisions from the CUDA path.
 * mode: 0 = Welch/discriminator test (MEMFNC==0); 1 = Moon et al. CI test on raw means (MEMFNC==1,
 *   SD.cu:125-144,256,325), which needs n/mean/m2 planes instead of mean_corr/disc.
 * ------------------------------------------------------------------------------------------ */
typedef struct smo_filter_args {
    int W, H;
    int C;              /* statistic channels: 1 or 3 */
    int value_channels; /* channels of value/out: 1 or 3 */
    int radius;
    float ds_factor; /* -0.5/sd^2, EST.h:259 */
    int n_gbufs;
    const smo_plane *gbufs;     /* [n_gbufs] float planes */
    const uint8_t *gbuf_channels; /* 1 or 3 each */
    const float *gbuf_dr_factors; /* -0.5/sd_g^2, EST.cpp:16 */
    const smo_plane *mean_corr, *disc; /* mode 0 */
    const smo_plane *n, *mean, *m2;    /* mode 1 */
    const float *lut;                  /* mode 1 */
    const smo_plane *value;
    const smo_plane *out;
    const smo_plane *accepted; /* may be NULL */
    int mode;
} smo_filter_args;

static inline int member_welch(int C, const float *mC, const float *dC, const float *mI, const float *dI) {
    /* SD.cu:81-88 */
    for (int c = 0; c < C; c++)
        if (!(dC[c] + dI[c] <= 2.f * mC[c] * mI[c])) return 0;
    return 1;
}

static inline int member_moon(int C, int nC, const float *meanC, const float *m2C, const float *meanI,
                              const float *lut) {
    /* SD.cu:125-144: t index n-2 (signed compare), se = t*sqrtf(m2/(n(n-1))), meanI in [meanC-se, meanC+se] */
    const float t = lut_t(lut, nC - 2);
    const float nCF = (float)nC;
    for (int c = 0; c < C; c++) {
        const float se = t * sqrtf(m2C[c] / (nCF * (nCF - 1.f)));
        if (!(meanI[c] >= meanC[c] - se && meanI[c] <= meanC[c] + se)) return 0;
    }
    return 1;
}

#define SMO_DEFINE_FILTER(NAME, real, EXPFN, FMA)                                                                     \
    void NAME(const smo_filter_args *a) {                                                                             \
        const int W = a->W, H = a->H, C = a->C, VC = a->value_channels, rad = a->radius;                              \
        const int rad2 = rad * rad;                                                                                   \
        _Pragma("omp parallel for schedule(dynamic, 1)") for (int yC = 0; yC < H; yC++) {                             \
            for (int xC = 0; xC < W; xC++) {                                                                          \
                const float *mC = NULL, *dC = NULL, *meanC = NULL, *m2C = NULL;                                       \
                int nC = 0;                                                                                           \
                if (a->mode == 0) {                                                                                   \
                    mC = rowf(a->mean_corr, yC) + xC * C;                                                             \
                    dC = rowf(a->disc, yC) + xC * C;                                                                  \
                } else {                                                                                              \
                    nC = rowi(a->n, yC)[xC];                                                                          \
                    meanC = rowf(a->mean, yC) + xC * C;                                                               \
                    m2C = rowf(a->m2, yC) + xC * C;                                                                   \
                }                                                                                                     \
                real num[3] = {0, 0, 0};                                                                              \
                real den = 0;                                                                                         \
                int acc = 0;                                                                                          \
                for (int yI = yC - rad; yI < yC + rad; yI++)     /* SD.cu:247, SET_OUTER :24-28: half-open */        \
                    for (int xI = xC - rad; xI < xC + rad; xI++) { /* SD.cu:248 */                                    \
                        const int y = clampi(yI, 0, H - 1), x = clampi(xI, 0, W - 1); /* SD.cu:31-32 */               \
                        const int dS2 = (yC - yI) * (yC - yI) + (xC - xI) * (xC - xI); /* SD.cu:33-34 unclamped */    \
                        if (dS2 > rad2) continue;                                      /* SD.cu:36 */                 \
                        const float *v = rowf(a->value, y) + x * VC;                                                  \
                        if (xI == xC && yI == yC) { /* SD.cu:78,250-254: centre, weight 1 */                          \
                            for (int c = 0; c < VC; c++) num[c] += (real)v[c];                                        \
                            den += (real)1;                                                                           \
                            acc++;                                                                                    \
                            continue;                                                                                 \
                        }                                                                                             \
                        int ok;                                                                                       \
                        if (a->mode == 0)                                                                             \
                            ok = member_welch(C, mC, dC, rowf(a->mean_corr, y) + x * C, rowf(a->disc, y) + x * C);    \
                        else                                                                                          \
                            ok = member_moon(C, nC, meanC, m2C, rowf(a->mean, y) + x * C, a->lut);                    \
                        if (!ok) continue;                                                                            \
                        /* dr2, SD.cu:90-112 */                                                                       \
                        real d2 = 0;                                                                                  \
                        for (int g = 0; g < a->n_gbufs; g++) {                                                        \
                            const int gc = a->gbuf_channels[g];                                                       \
                            const float *gC = rowf(&a->gbufs[g], yC) + xC * gc;                                       \
                            const float *gI = rowf(&a->gbufs[g], y) + x * gc;                                         \
                            if (gc == 3) {                                                                            \
                                real t = 0;                                                                           \
                                for (int c = 0; c < 3; c++) {                                                         \
                                    const real df = (real)gC[c] - (real)gI[c];                                        \
                                    t = FMA(df, df, t); /* [nvcc-fma] d2Tmp += d*d */                                 \
                                }                                                                                     \
                                d2 = FMA(t, (real)a->gbuf_dr_factors[g], d2); /* [nvcc-fma] d2 += d2Tmp*f */          \
                            } else if (gc == 1) {                                                                     \
                                const real df = (real)gC[0] - (real)gI[0];                                            \
                                d2 = FMA(df * df, (real)a->gbuf_dr_factors[g], d2);                                   \
                            }                                                                                         \
                        }                                                                                             \
                        /* SD.cu:261: expf(dS2 * dSFactor + dr2)   [nvcc-fma] */                                      \
                        const real w = EXPFN(FMA((real)dS2, (real)a->ds_factor, d2));                                 \
                        for (int c = 0; c < VC; c++) num[c] = FMA(w, (real)v[c], num[c]); /* [nvcc-fma] */            \
                        den += w;                                                                                     \
                        acc++;                                                                                        \
                    }                                                                                                 \
                float *o = rowf(a->out, yC) + xC * VC;                                                                \
                for (int c = 0; c < VC; c++) o[c] = (float)(num[c] / den); /* SD.cu:341-344 */                        \
                if (a->accepted) rowi(a->accepted, yC)[xC] = acc;                                                     \
            }                                                                                                         \
        }                                                                                                             \
    }

SMO_DEFINE_FILTER(smo_filter_f32, float, expf, fmaf)
SMO_DEFINE_FILTER(smo_filter_f64, double, exp, fma)

/* Two-sided tap count of the window, for the algorithmic-work figures:
 * #{(dy,dx) in [-r, r)^2 : dy^2+dx^2 <= r^2}. */
int smo_taps_in_window(int r) {
    int c = 0;
    for (int dy = -r; dy < r; dy++)
        for (int dx = -r; dx < r; dx++)
            if (dy * dy + dx * dx <= r * r) c++;
    return c;
}

/* Chan/Pebay pairwise combination of two moment sets (n, mean, M2, M3), in double.
 * The reference has NO such merge (EST.cpp:341-352 just copies running totals); this is the
 * checker for the new smc_merge_moments capability: tests compare it and the sequential
 * accumulation of the concatenated sample stream within tolerance. */
void smo_merge_moments_f64(int64_t npix, int C, const int64_t *nA, const double *meanA, const double *m2A,
                           const double *m3A, const int64_t *nB, const double *meanB, const double *m2B,
                           const double *m3B, int64_t *nO, double *meanO, double *m2O, double *m3O) {
    for (int64_t p = 0; p < npix; p++) {
        const double na = (double)nA[p], nb = (double)nB[p], nn = na + nb;
        nO[p] = nA[p] + nB[p];
        for (int c = 0; c < C; c++) {
            const size_t i = (size_t)p * C + c;
            if (nn == 0) {
                meanO[i] = m2O[i] = m3O[i] = 0;
                continue;
            }
            const double d = meanB[i] - meanA[i];
            meanO[i] = meanA[i] + d * nb / nn;
            m2O[i] = m2A[i] + m2B[i] + d * d * na * nb / nn;
            m3O[i] = m3A[i] + m3B[i] + d * d * d * na * nb * (na - nb) / (nn * nn) +
                     3.0 * d * (na * m2B[i] - nb * m2A[i]) / nn;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * PROTOTYPE (DESIGN.md "what comes next" item 1): the filter with every unordered pair of positions evaluated ONCE.
 *
 * Membership (SD.cu:81-88) and the weight (SD.cu:90-112, :261) are symmetric in (C, I), so for a forward offset
 * (dy, dx) -- dy == 0 and dx in [1, r], or dy in [1, r] and dx in [-r, r], inside the disc -- the pair P, Q = P + (dy, dx)
 * is evaluated once and contributes  w * V(Q) to P  when (dy, dx) is a tap of P's half-open window (dy <= r-1, dx <= r-1),
 * and  w * V(P) to Q  when (-dy, -dx) is a tap of Q's (dx >= -(r-1)).  Replicated borders: P and Q range over the image
 * extended by r on every side, their records are the clamped ones (BrdReplicate), and only positions inside the image
 * receive anything.  Sums in double: the result differs from smo_filter_f64 only by the order of summation, which is what
 * tests/test_oracle_cpu.py::test_symmetric_pair_evaluation_prototype asserts.  Mode 0 (Welch) only: the Moon test is not
 * symmetric.  Single-threaded over rows of P with per-thread row accumulators merged at the end (simple, not fast).
 * ------------------------------------------------------------------------------------------ */
void smo_filter_sym_f64(const smo_filter_args *a) {
    const int W = a->W, H = a->H, C = a->C, VC = a->value_channels, rad = a->radius, rad2 = rad * rad;
    double *num = (double *)calloc((size_t)W * H * 3, sizeof(double));
    double *den = (double *)calloc((size_t)W * H, sizeof(double));
    int *acc = (int *)calloc((size_t)W * H, sizeof(int));
    for (int yP = -rad; yP < H; yP++)
        for (int xP = -rad; xP < W + rad; xP++) {
            const int ypc = clampi(yP, 0, H - 1), xpc = clampi(xP, 0, W - 1);
            const int p_real = yP >= 0 && xP >= 0 && xP < W; /* yP < H by the loop */
            const float *mP = rowf(a->mean_corr, ypc) + xpc * C, *dP = rowf(a->disc, ypc) + xpc * C;
            const float *vP = rowf(a->value, ypc) + xpc * VC;
            if (p_real) { /* centre tap, weight 1 (SD.cu:78, 250-254) */
                for (int c = 0; c < VC; c++) num[((size_t)yP * W + xP) * 3 + c] += (double)vP[c];
                den[(size_t)yP * W + xP] += 1.0;
                acc[(size_t)yP * W + xP]++;
            }
            for (int dy = 0; dy <= rad; dy++)
                for (int dx = (dy == 0 ? 1 : -rad); dx <= rad; dx++) {
                    if (dy * dy + dx * dx > rad2) continue;
                    const int yQ = yP + dy, xQ = xP + dx;
                    const int q_real = yQ >= 0 && yQ < H && xQ >= 0 && xQ < W;
                    const int to_p = p_real && dy <= rad - 1 && dx <= rad - 1; /* (dy, dx) in P's window [-r, r) */
                    const int to_q = q_real && dx >= -(rad - 1);               /* (-dy, -dx) in Q's window */
                    if (!to_p && !to_q) continue;
                    const int yqc = clampi(yQ, 0, H - 1), xqc = clampi(xQ, 0, W - 1);
                    if (!member_welch(C, mP, dP, rowf(a->mean_corr, yqc) + xqc * C, rowf(a->disc, yqc) + xqc * C)) continue;
                    double d2 = 0;
                    for (int g = 0; g < a->n_gbufs; g++) {
                        const int gc = a->gbuf_channels[g];
                        const float *gP = rowf(&a->gbufs[g], ypc) + xpc * gc, *gQ = rowf(&a->gbufs[g], yqc) + xqc * gc;
                        double t = 0;
                        for (int c = 0; c < gc; c++) {
                            const double df = (double)gP[c] - (double)gQ[c];
                            t = fma(df, df, t);
                        }
                        d2 = fma(t, (double)a->gbuf_dr_factors[g], d2);
                    }
                    const double w = exp(fma((double)(dy * dy + dx * dx), (double)a->ds_factor, d2));
                    const float *vQ = rowf(a->value, yqc) + xqc * VC;
                    if (to_p) {
                        for (int c = 0; c < VC; c++) num[((size_t)yP * W + xP) * 3 + c] += w * (double)vQ[c];
                        den[(size_t)yP * W + xP] += w;
                        acc[(size_t)yP * W + xP]++;
                    }
                    if (to_q) {
                        for (int c = 0; c < VC; c++) num[((size_t)yQ * W + xQ) * 3 + c] += w * (double)vP[c];
                        den[(size_t)yQ * W + xQ] += w;
                        acc[(size_t)yQ * W + xQ]++;
                    }
                }
        }
    for (int y = 0; y < H; y++) {
        float *o = rowf(a->out, y);
        for (int x = 0; x < W; x++) {
            for (int c = 0; c < VC; c++) o[x * VC + c] = (float)(num[((size_t)y * W + x) * 3 + c] / den[(size_t)y * W + x]);
            if (a->accepted) rowi(a->accepted, y)[x] = acc[(size_t)y * W + x];
        }
    }
    free(num);
    free(den);
    free(acc);
}
